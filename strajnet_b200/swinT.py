"""Module alias expected by the reference's inference driver (`from swinT import ...`, inference.py:7,141).

`SwinTransformerDecoder` exists nowhere in the reference; it is exported as a stub that raises.
"""
from .layers import STrajNet, SwinTransformerEncoder  # noqa: F401

CFGS = {
    "strajnet_256": dict(input_size=(256, 256), window_size=8, embed_dim=96, depths=[2, 2, 2], num_heads=[3, 6, 12]),
    "strajnet_512": dict(input_size=(512, 512), window_size=8, embed_dim=96, depths=[2, 2, 2], num_heads=[3, 6, 12]),
}


class SwinTransformerDecoder:  # pragma: no cover - mirrors a name the reference imports but never defines
    def __init__(self, *a, **k):
        raise NotImplementedError("SwinTransformerDecoder is imported by inference.py:7 but defined nowhere in the reference")
