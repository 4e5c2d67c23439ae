"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, struct
layouts agree, argument validation works without a GPU, weight preparation is consistent."""
import ctypes as C
import os
import re

import pytest
import torch

from strajnet_b200 import _lib as L
from strajnet_b200 import weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "strajnet_b200.h")).read()
    declared = set(re.findall(r"\b(sj_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = L.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    assert lib.sj_version() >= 100


def test_struct_layouts_match():
    lib = L.lib()
    structs = [L.SjLinear, L.SjNorm, L.SjSwinBlockW, L.SjPatchMergeW, L.SjPatchEmbedW, L.SjBasicLayerW, L.SjEncoderW,
               L.SjFgmsaW, L.SjTrajW, L.SjDecoderW, L.SjModelW]
    for i, s in enumerate(structs):
        assert lib.sj_sizeof(i) == C.sizeof(s)


def test_status_strings_and_errors():
    lib = L.lib()
    assert lib.sj_strerror(0) == b"ok"
    for s in (-1, -2, -3, -4):
        assert lib.sj_strerror(s) not in (b"ok", b"unknown status")
    with pytest.raises(ValueError):
        L.check(L.SJ_EINVAL, "x")
    with pytest.raises(L.SjError):
        L.check(L.SJ_EWORKSPACE, "x")
    # null pointers are rejected before anything touches the device
    assert lib.sj_strajnet_fwd(None, None, None, None, None, None, None, 1, 256, 0, None, 0, None) == L.SJ_EINVAL
    assert lib.sj_relative_position_index(8, None, None) == L.SJ_EINVAL


def test_workspace_queries_scale_with_batch():
    lib = L.lib()
    w1 = lib.sj_strajnet_workspace_bytes(1, 256, L.SJ_F32)
    w4 = lib.sj_strajnet_workspace_bytes(4, 256, L.SJ_F32)
    assert 0 < w1 < w4 <= 4 * w1 + (1 << 20)
    assert lib.sj_strajnet_workspace_bytes(4, 256, L.SJ_BF16) < w4
    assert lib.sj_swin_block_workspace_bytes(1, 64, 64, 32, L.SJ_F32) > 4096 * 32 * 4 * 9


def test_packer_builds_model_struct_on_cpu():
    from oracle import strajnet_oracle as O
    w = O.make_weights(O.CFG256, seed=0)
    assert set(W.model_shapes(O.CFG256, True, True)) == set(O.weight_shapes(O.CFG256, True, True))
    p = W.Packer(w, "cpu", tc=True)
    m = p.model(O.CFG256, True, True, False)
    assert m.encoder.num_layers == 3 and m.encoder.layers[2].dim == 384 and m.encoder.layers[2].heads == 12
    assert m.decoder.upconv[0].w and m.decoder.upconv[0].w_tc and m.traj.ca_q.w
    assert m.fg_msa == 1 and m.large_ogm == 0


def test_tfa_relayout():
    k = torch.arange(3 * 5 * 4, dtype=torch.float32).reshape(3, 5, 4)
    m = W.tfa_in_kernel(k, 16)
    assert m.shape == (5, 16) and torch.equal(m[:, 4:8], k[1]) and torch.all(m[:, 12:] == 0)
    x = torch.randn(7, 5)
    assert torch.allclose(x @ m[:, :12], torch.einsum("ni,hio->nho", x, k).reshape(7, 12))
    pk = torch.randn(3, 4, 6)
    o = torch.randn(7, 3, 4)
    assert torch.allclose(o.reshape(7, 12) @ W.tfa_out_kernel(pk), torch.einsum("nhi,hio->no", o, pk), atol=1e-6)


def test_layer_argument_validation():
    import strajnet_b200 as sj
    with pytest.raises(ValueError):
        sj.STrajNet(dict(input_size=(256, 256), window_size=8, embed_dim=96, depths=[2, 2, 2], num_heads=[3, 6, 12]),
                    large_ogm=True, device="cpu")
    with pytest.raises(ValueError):
        sj.STrajNet(dict(input_size=(256, 256), window_size=8, embed_dim=96, depths=[2, 2, 2], num_heads=[3, 6, 12]),
                    large_ogm=False, fg=True, fg_msa=False, device="cpu")
    m = sj.STrajNet(dict(input_size=(256, 256), window_size=8, embed_dim=96, depths=[2, 2, 2], num_heads=[3, 6, 12]),
                    large_ogm=False, fg_msa=True, fg=True, device="cpu")
    with pytest.raises(NotImplementedError):
        m(None, None, training=True)
    with pytest.raises(AssertionError):
        sj.SwinTransformerBlock(32, (64, 64), 2, window_size=8, shift_size=8, device="cpu")
    blk = sj.SwinTransformerBlock(32, (4, 4), 2, window_size=8, shift_size=4, device="cpu")
    assert blk.window_size == 4 and blk.shift_size == 0  # modules.py:173-175
