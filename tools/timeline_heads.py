"""Development aid: phase timeline (clock64) of CTA 0 of the inference HEADS kernel.  python tools/timeline_heads.py [precision]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import npcd_b200  # noqa: E402,F401
from npcd_b200 import _lib, synthetic as syn  # noqa: E402
from npcd_b200.pointnerf import PointNeRF  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "f16+e4m3x2"
dev = torch.device("cuda")
m = PointNeRF(1, 32, 512, False).eval().to(dev)
sd = m.state_dict()
with torch.no_grad():
    for k, v in syn.make_weights(0).items():
        sd[k].copy_(torch.from_numpy(v))
m.field.precision = prec
poses, intr = syn.load_cameras()
coords, feats = syn.make_clouds([0])
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
args = (t(coords), t(feats), t(poses[:64][None]), t(intr[:64][None]), 128, False)
with torch.no_grad():
    m.renderer(*args)
    buf = torch.zeros((64, 32), dtype=torch.int64, device=dev)
    _lib.call("npcd_debug_set_timeline_heads", buf.data_ptr())
    m.renderer(*args)
    torch.cuda.synchronize()
    _lib.call("npcd_debug_set_timeline_heads", None)
T = buf.cpu().numpy().astype(np.float64)
sel = slice(8, 60)
names = {10: "tile done"}
for l in range(5):
    names[2 * l] = f"epi: before wait L{l}"
    names[2 * l + 1] = f"epi: acc L{l} ready"
for l in range(4):
    names[16 + 3 * l] = f"mma: L{l} start"
    names[17 + 3 * l] = f"mma: L{l} operand there"
    names[18 + 3 * l] = f"mma: L{l} issued"
base = T[sel, 0:1]
rel = T[sel] - base
period = np.diff(T[sel, 0]).mean()
print(f"heads, precision {prec}: tile period {period:.0f} cycles (CTA 0, tiles 8..59)")
print(f"  issue loop waited per tile (median): weights {np.median(T[sel, 28]):.0f}, operands {np.median(T[sel, 29]):.0f}, "
      f"accumulator-free {np.median(T[sel, 30]):.0f} cycles")
for e in sorted(names, key=lambda e: np.median(rel[:, e])):
    print(f"  {np.median(rel[:, e]):9.0f}  {names[e]}")
