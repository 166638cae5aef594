"""Small host-side helpers mirroring the reference utilities the hot path uses."""
from __future__ import annotations

from typing import Optional

from torch import nn


class AttrDict(dict):
    """Attribute- and key-accessible dict (stands in for ``easydict.EasyDict`` returned at `renderers/renderer.py:268`;
    callers use ``pred.channels`` and ``pred.get("ray_idx")``: `npcd/losses/image_reconstruction_loss.py:33-34`)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def define_mlp(dims, d_in, d_out: Optional[int] = None, act: str = "ReLU", layer_norm: bool = True):
    """Same module tree (and therefore the same ``state_dict`` keys 0,2,4,...) as `npcd/utils/model.py:22-36`."""
    if layer_norm:
        raise NotImplementedError("layer_norm=True is never used by the reference configuration (pointnerf.py:164,179)")
    act_cls = getattr(nn, act)
    mods, cur = [], d_in
    for dim in dims:
        mods += [nn.Linear(cur, dim), act_cls(inplace=True)]
        cur = dim
    if d_out is not None:
        mods.append(nn.Linear(cur, d_out))
    return nn.Sequential(*mods)
