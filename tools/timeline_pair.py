"""Development aid: phase timeline (clock64) of CTA 0 of the inference pair kernel.  python tools/timeline_pair.py [precision]"""
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
    _lib.call("npcd_debug_set_timeline", buf.data_ptr())
    m.renderer(*args)
    torch.cuda.synchronize()
    _lib.call("npcd_debug_set_timeline", None)
T = buf.cpu().numpy().astype(np.float64)
sel = slice(8, 60)
names = {0: "epi: before wait L0", 1: "epi: acc L0 ready", 2: "epi: L0 done / before wait L1", 3: "epi: acc L1 ready", 4: "epi: L1 done / before wait L2",
         5: "epi: acc L2 ready", 6: "epi: L2 done / before wait L3", 7: "epi: acc L3 ready", 8: "agg: staged pass 0", 9: "agg: after bar", 10: "agg: summed pass 0",
         11: "agg: after bar", 12: "agg: staged pass 1", 13: "agg: after bar", 14: "agg: summed pass 1", 15: "agg: after bar (tile done)",
         16: "mma: L0 start", 17: "mma: L0 operand there", 18: "mma: L0 issued", 19: "mma: L1 start", 20: "mma: L1 operand there", 21: "mma: L1 issued",
         22: "mma: L2 start", 23: "mma: L2 operand there", 24: "mma: L2 issued", 25: "mma: L3 start", 26: "mma: L3 operand there", 27: "mma: L3 issued"}
base = T[sel, 0:1]
rel = T[sel] - base
period = np.diff(T[sel, 0]).mean()
print(f"precision {prec}: tile period {period:.0f} cycles (CTA 0, tiles 8..59)")
print(f"  issue loop waited per tile (median): weights {np.median(T[sel, 28]):.0f}, operands {np.median(T[sel, 29]):.0f}, "
      f"accumulator-free {np.median(T[sel, 30]):.0f} cycles")
order = sorted(names, key=lambda e: np.median(rel[:, e]))
for e in order:
    print(f"  {np.median(rel[:, e]):9.0f}  {names[e]}")
