"""Builds ``csrc/*.cu`` into the in-tree shared library ``libnpcd_b200.so`` with nvcc for sm_100a (no torch dependency)."""
from __future__ import annotations

import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libnpcd_b200.so")
OBJ_DIR = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-I", INCLUDE, "-I", CSRC,
] + os.environ.get("NPCD_NVCC_DEFINES", "").split()  # development aid: extra -D switches (part of the build digest)


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isfile(c) or c == "nvcc"):
            return c
    raise RuntimeError("nvcc not found")


def have_nvcc() -> bool:
    try:
        n = _nvcc()
    except RuntimeError:
        return False
    if n == "nvcc":
        import shutil

        return shutil.which("nvcc") is not None
    return True


def _digest(paths) -> str:
    """Content digest of the sources + build switches.  File NAMES (not absolute paths) and the flags without the include
    directories go in, so the stamp of a library built in one checkout still matches after the tree is copied elsewhere (gpurun
    ships the built .so to a scratch path: hashing absolute paths made every fresh box rebuild at first import -- and two ranks of a
    torchrun job race on that rebuild)."""
    h = hashlib.sha256()
    for p in sorted(paths, key=os.path.basename):
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    flags = [f for f in NVCC_FLAGS if f not in (INCLUDE, CSRC)]
    h.update(" ".join(flags).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    headers = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(INCLUDE, "*.h")))
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(OBJ_DIR, "stamp.txt")
    digest = _digest(sources + headers)

    def up_to_date() -> bool:
        return os.path.isfile(LIB) and os.path.isfile(stamp) and open(stamp).read() == digest

    if not force and up_to_date():
        return LIB
    # one builder at a time (ranks of a multi-process job import the package simultaneously); whoever waited re-checks the stamp
    import fcntl

    lock = open(os.path.join(OBJ_DIR, "build.lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and up_to_date():
            return LIB
        return _build_locked(sources, stamp, digest, verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(sources, stamp, digest, verbose) -> str:
    nvcc = _nvcc()
    objs, procs = [], []
    for src in sources:
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj] + (["-Xptxas", "-v"] if verbose else [])
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {os.path.basename(src)} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    tmp = LIB + f".tmp{os.getpid()}"  # link next to the target, then rename: a concurrent reader never sees a partial file
    subprocess.check_call([nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    os.replace(tmp, LIB)
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
