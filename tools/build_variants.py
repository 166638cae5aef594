"""Development aid: prebuilds experimental variants of the library (extra -D switches on csrc/mlp_tc.cu only) next to the
default objects, so that a GPU call can time them without compiling:  NPCD_LIB_PATH=<variant .so> python bench.py ...

usage: python tools/build_variants.py tag1="-DNPCD_EXP_NOCVT=1" tag2="-DNPCD_EXP_NOF8MMA=1 -DNPCD_EXP_NOSTS8=1" ...
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import npcd_b200  # noqa: E402,F401
from npcd_b200 import build as B  # noqa: E402

B.build()
out_dir = os.path.join(B.OBJ_DIR, "variants")
os.makedirs(out_dir, exist_ok=True)
base_flags = [f for f in B.NVCC_FLAGS]
procs = []
for arg in sys.argv[1:]:
    tag, defs = arg.split("=", 1)
    obj = os.path.join(out_dir, f"mlp_tc_{tag}.o")
    cmd = [B._nvcc(), *base_flags, *defs.split(), "-c", os.path.join(B.CSRC, "mlp_tc.cu"), "-o", obj]
    procs.append((tag, obj, subprocess.Popen(cmd)))
for tag, obj, p in procs:
    if p.wait() != 0:
        raise SystemExit(f"variant {tag}: nvcc failed")
    objs = [os.path.join(B.OBJ_DIR, f) for f in sorted(os.listdir(B.OBJ_DIR)) if f.endswith(".o") and f != "mlp_tc.o"] + [obj]
    lib = os.path.join(out_dir, f"libnpcd_{tag}.so")
    subprocess.check_call([B._nvcc(), "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    print(lib)
