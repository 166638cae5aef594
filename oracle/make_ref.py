"""Recipe for ``oracle/_ref/`` -- the UNMODIFIED reference Python package, staged so that it travels to the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY (the product never imports anything under ``oracle/``).

The reference (`/root/reference`) is pure Python; `/root/reference` exists in the build container only.  This script copies the
``npcd/`` package tree byte for byte into the git-ignored directory ``oracle/_ref/npcd/`` (outputs go nowhere else; nothing from
the reference enters the repository history -- ``oracle/_ref/`` is listed in ``.gitignore`` but NOT in ``.gpurunignore``, so
``gpurun`` ships it like our own built ``.so``).  ``oracle/ref_loader.py`` then imports it with the six stub modules of SURVEY.md
Appendix C.  Used as

  * the timed CPU baseline of ``bench.py`` (``--impl reference`` and the ``cpu_baseline`` leg, ``kind: "reference"``):
    `npcd/models/pointnerf/pointnerf.py:126` ``render`` with the reference's own pure-torch kNN branch
    (`fields/aggregators/aggregator.py:42-58`: ``voxel_grid=None``, ``r=0.08``), all host cores;
  * the checker of ``tests/test_reference_callers.py`` (the reference's own loss / trainer step / evaluation chunk running on
    top of the drop-in renderer).

``__graft_entry__.build()`` calls ``make()`` whenever `/root/reference` is present.

    python oracle/make_ref.py
"""
from __future__ import annotations

import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")


def make(verbose: bool = False) -> str | None:
    """Copies `/root/reference/npcd/**/*.py` to ``oracle/_ref/npcd/`` (unmodified).  Returns the directory, or None when the
    reference is not on this machine (GPU box: the prebuilt copy, if any, is used as is)."""
    src = os.path.join(REF_SRC, "npcd")
    if not os.path.isdir(src):
        return REF_DST if os.path.isdir(os.path.join(REF_DST, "npcd")) else None
    n = 0
    for d, _, files in os.walk(src):
        rel = os.path.relpath(d, REF_SRC)
        if "__pycache__" in rel:
            continue
        for f in files:
            if not f.endswith(".py"):
                continue
            dst_dir = os.path.join(REF_DST, rel)
            os.makedirs(dst_dir, exist_ok=True)
            s, t = os.path.join(d, f), os.path.join(dst_dir, f)
            if not (os.path.isfile(t) and filecmp.cmp(s, t, shallow=False)):
                shutil.copyfile(s, t)
                n += 1
    with open(os.path.join(REF_DST, "README.txt"), "w") as fh:
        fh.write("Unmodified copy of /root/reference/npcd (made by oracle/make_ref.py); git-ignored, never edited, never imported by "
                 "the product.\n")
    if verbose:
        print(f"oracle/_ref: {n} file(s) refreshed")
    return REF_DST


if __name__ == "__main__":
    print(make(verbose=True))
