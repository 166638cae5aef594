"""Checkpoint layout of the UNMODIFIED reference `PointNeRF`: state_dict keys, shapes and dtypes, plus the
layout of the embedding `_extra_state` entries (`npcd/utils/flex_embedding.py:9-25`, `embeddings/embedding.py:53-60`).
Run in the build container only (needs /root/reference):  python tests/golden/make_golden_state_dict.py
Output: tests/golden/state_dict_layout.json (committed)."""
import importlib.util
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)


def describe(v):
    if isinstance(v, torch.Tensor):
        return {"shape": list(v.shape), "dtype": str(v.dtype)}
    if isinstance(v, dict):
        return {k: describe(x) for k, x in v.items()}
    return {"type": type(v).__name__}


def main():
    mg.install_stubs()
    from npcd.models.pointnerf.pointnerf import PointNeRF

    m = PointNeRF(3, 32, 512, False)
    sd = m.state_dict()
    out = {"n_obj": 3, "entries": {k: describe(v) for k, v in sd.items()}}
    with open(os.path.join(HERE, "state_dict_layout.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(len(sd), "entries")


if __name__ == "__main__":
    main()
