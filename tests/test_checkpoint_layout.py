"""checkpoints of the reference load into the drop-in.  The golden layout (keys, shapes, dtypes, structure of
the embedding `_extra_state` entries) comes from the UNMODIFIED reference `PointNeRF` (tests/golden/make_golden_state_dict.py); the
drop-in must produce exactly that layout and round-trip through `load_state_dict`.  CPU only: module construction needs no GPU."""
import json
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _describe(v):
    if isinstance(v, torch.Tensor):
        return {"shape": list(v.shape), "dtype": str(v.dtype)}
    if isinstance(v, dict):
        return {k: _describe(x) for k, x in v.items()}
    return {"type": type(v).__name__}


def test_state_dict_layout_matches_reference():
    import npcd_b200  # noqa: F401
    from npcd_b200.pointnerf import PointNeRF

    gold = json.load(open(os.path.join(HERE, "golden", "state_dict_layout.json")))
    m = PointNeRF(gold["n_obj"], 32, 512, False)
    ours = {k: _describe(v) for k, v in m.state_dict().items()}
    assert set(ours) == set(gold["entries"]), (set(ours) ^ set(gold["entries"]))
    for k, d in gold["entries"].items():
        assert ours[k] == d, (k, ours[k], d)


def test_reference_style_checkpoint_round_trip():
    """A checkpoint written with the reference's keys (random values) restores every parameter AND the embedding tables."""
    import npcd_b200  # noqa: F401
    from npcd_b200.pointnerf import PointNeRF

    torch.manual_seed(0)
    src = PointNeRF(2, 32, 512, False)
    with torch.no_grad():
        src.feats.get_emb().weight.normal_()
        src.coords.get_emb().weight.uniform_(-0.5, 0.5)
    sd = {k: (v.clone() if isinstance(v, torch.Tensor) else {"emb": {"weight": v["emb"]["weight"].clone()}}) for k, v in src.state_dict().items()}
    dst = PointNeRF(2, 32, 512, False)
    missing, unexpected = dst.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    for (ka, a), (kb, b) in zip(src.named_parameters(), dst.named_parameters()):
        assert ka == kb and torch.equal(a, b), ka
    assert torch.equal(dst.get_all_coords(), src.get_all_coords()) and torch.equal(dst.get_all_feats(), src.get_all_feats())
    # `field.*` and `renderer.field.*` are the same tensors (the reference registers the field twice, pointnerf.py:27-28)
    assert dst.field.channel_net[0].weight is dst.renderer.field.channel_net[0].weight
