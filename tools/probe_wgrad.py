"""MN-major descriptor probe: which of the (LBO, SBO) assignments reproduces A^T B?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import npcd_b200  # noqa
from npcd_b200 import ops

torch.manual_seed(0)
for rows, m, n in [(64, 256, 256), (128, 256, 256), (1000, 256, 256), (129, 3, 256), (300, 256, 112)]:
    a = torch.randn(rows, m) * 1e-2
    b = torch.randn(rows, n)
    want = a.double().t() @ b.double()
    ia, ib = ops.tc_pack(a.cuda()), ops.tc_pack(b.cuda())
    for flags in (0, 1):
        got = ops.tc_wgrad(ia, ib, flags=flags, row_splits=1 if rows < 500 else 0).cpu().double()
        torch.cuda.synchronize()
        err = (got - want).abs().max().item() / want.abs().max().item()
        print(f"rows {rows} {m}x{n} flags {flags}: rel err {err:.3e}")
