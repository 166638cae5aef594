"""B200-native PointNeRF render path for NPCD (drop-in for ``npcd.models.pointnerf`` renderer call).

Import as ``npcd_b200`` (see ``/npcd_b200.py``).  Sub-modules are imported lazily: ``synthetic`` is numpy-only,
everything that touches the GPU goes through ``_lib`` (ctypes over the C-ABI ``libnpcd_b200.so``) and raises
if the CUDA library is missing -- there is no CPU fallback.
"""
__all__ = ["synthetic"]
