"""CPU tests of the oracle itself: pins against the reference's own NumPy lines (golden fixtures),
an independent cross-check against torchvision, algebraic identities the CUDA path relies on."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import strajnet_oracle as O
from tests.util import randn


def test_relative_position_index_matches_reference_lines(golden_dir):
    g = np.load(f"{golden_dir}/ref_index_maps.npz")
    rpi = O.relative_position_index(8)
    assert rpi.dtype == np.int64 and rpi.shape == (64, 64)
    assert np.array_equal(rpi, g["relative_position_index_ws8"])
    # SURVEY App. D known answers
    assert (rpi[0, 0], rpi[0, 1], rpi[0, 8], rpi[0, 63], rpi[63, 0]) == (112, 111, 97, 0, 224)
    assert rpi.sum() == 458752
    assert hashlib.sha256(np.ascontiguousarray(rpi).tobytes()).hexdigest().startswith("4a65ba32bf59f64d")


@pytest.mark.parametrize("H,count,sha", [(16, 7168, "b0e03e309bdc1d5d"), (32, 15360, "f6078878e7abed4c"),
                                         (64, 31744, "ab76e9ac4c9729a2"), (128, 64512, "420c8988c61c9971")])
def test_shift_mask_matches_reference_lines(golden_dir, H, count, sha):
    g = np.load(f"{golden_dir}/ref_index_maps.npz")
    m = O.shift_attn_mask(H, H, 8, 4)
    assert np.array_equal(m.astype(np.int8), g[f"shift_mask_{H}"])
    assert int((m != 0).sum()) == count
    assert hashlib.sha256(m.astype(np.float32).tobytes()).hexdigest().startswith(sha)


def test_partition_reverse_roundtrip():
    x = randn((2, 16, 24, 8))
    w = O.window_partition(x, 8)
    assert w.shape == (2 * 2 * 3, 8, 8, 8)
    assert torch.equal(O.window_reverse(w, 8, 16, 24, 8), x)
    # window id = b*nW + (i//8)*(W/8) + j//8 ; token = (i%8)*8 + j%8  
    idx = torch.arange(16 * 24, dtype=torch.float32).reshape(1, 16, 24, 1)
    wi = O.window_partition(idx, 8).reshape(-1, 64)
    for (i, j) in [(0, 0), (9, 17), (15, 23)]:
        assert wi[(i // 8) * 3 + j // 8, (i % 8) * 8 + j % 8] == i * 24 + j


def test_oracle_outputs_match_golden(golden_dir):
    g = np.load(f"{golden_dir}/oracle_outputs.npz")
    for shift in (0, 4):
        for heads in (1, 2):
            w = O.make_block_weights(32, heads, seed=0)
            x = randn((1, 4096, 32), seed=0)
            y = O.swin_block(x, w, "", 64, 64, heads, 8, shift)
            np.testing.assert_allclose(y[0, ::37].numpy(), g[f"block_c32_h{heads}_s{shift}"], atol=2e-5, rtol=0)


def test_full_forward_matches_golden(golden_dir):
    g = np.load(f"{golden_dir}/oracle_outputs.npz")
    w = O.make_weights(O.CFG256, seed=0)
    inp = O.make_inputs(1, 256, seed=0)
    y = O.forward_from_inputs(w, O.CFG256, inp)
    assert y.shape == (1, 256, 256, 32)
    np.testing.assert_allclose(y[0, ::16, ::16].numpy(), g["forward_cfg256_fg"], atol=2e-4, rtol=0)


@pytest.mark.parametrize("shift", [0, 4])
def test_window_attention_vs_torchvision(shift):
    """Independent cross-check of modules.py:103-134/220-258 against torchvision's implementation."""
    from torchvision.models.swin_transformer import shifted_window_attention
    C, heads, H = 96, 3, 32
    w = O.make_block_weights(C, heads, seed=3)
    x = randn((2, H * H, C), seed=4)
    # attention half only: feed identity norms
    w2 = dict(w)
    w2["norm1.gamma"], w2["norm1.beta"] = torch.ones(C), torch.zeros(C)
    y = layer = None
    # oracle attention half on pre-normalised input
    xn = O.layer_norm(x, w2["norm1.gamma"], w2["norm1.beta"], 1e-5)
    yo = xn.reshape(2, H, H, C)
    if shift:
        yo = torch.roll(yo, (-shift, -shift), (1, 2))
        mask = torch.from_numpy(O.shift_attn_mask(H, H, 8, shift))
    else:
        mask = None
    aw = O.window_attention(O.window_partition(yo, 8).reshape(-1, 64, C), w, "attn.", heads, 8, mask)
    yo = O.window_reverse(aw.reshape(-1, 8, 8, C), 8, H, H, C)
    if shift:
        yo = torch.roll(yo, (shift, shift), (1, 2))
    idx = torch.from_numpy(O.relative_position_index(8).reshape(-1))
    rpb = w["attn.relative_position_bias_table"][idx].reshape(64, 64, heads).permute(2, 0, 1)[None]
    yt = shifted_window_attention(xn.reshape(2, H, H, C), w["attn.qkv.kernel"].t(), w["attn.proj.kernel"].t(), rpb,
                                  [8, 8], heads, [shift, shift], qkv_bias=w["attn.qkv.bias"],
                                  proj_bias=w["attn.proj.bias"])
    assert (yo - yt).abs().max() < 2e-5


def test_fully_masked_rows_are_uniform():
    """tfa additive mask -1e10 in fp32: a fully masked row softmaxes to uniform weights (Q7)."""
    w = {"query_kernel": randn((2, 8, 4), 1), "key_kernel": randn((2, 8, 4), 2), "value_kernel": randn((2, 8, 4), 3),
         "projection_kernel": randn((2, 4, 6), 4), "projection_bias": torch.zeros(6)}
    q, k = randn((1, 3, 8), 5), randn((1, 5, 8), 6)
    mask = torch.zeros(1, 3, 5, dtype=torch.int32)
    out = O.tfa_mha(q, k, k, w, "", mask)
    V = torch.einsum("bmi,hio->bmho", k, w["value_kernel"]).mean(1, keepdim=True)  # uniform average
    ref = torch.einsum("bnhi,hio->bno", V.expand(-1, 3, -1, -1), w["projection_kernel"])
    assert (out - ref).abs().max() < 1e-5
    out64 = O.tfa_mha(q.double(), k.double(), k.double(), {a: b.double() for a, b in w.items()}, "", mask)
    assert (out64.float() - ref).abs().max() < 1e-5


def test_bilinear_sampler_properties():
    img = randn((1, 31, 31, 1), 7)
    pts = torch.tensor([[[3.0, 5.0], [3.5, 5.0], [-1.0, 4.0], [30.0, 30.0], [31.0, 2.0], [-5.0, -5.0], [30.5, 0.0]]])
    out = O.bilinear_sample_zero(img, pts)[0, :, 0]
    assert out[0] == img[0, 5, 3, 0]                                   # (x=3,y=5) -> image[row 5, col 3]
    assert torch.isclose(out[1], 0.5 * (img[0, 5, 3, 0] + img[0, 5, 4, 0]))
    assert out[2] == 0 and out[4] == 0 and out[5] == 0                 # zero border
    assert out[3] == img[0, 30, 30, 0]
    assert torch.isclose(out[6], 0.5 * img[0, 0, 30, 0])               # half way into the zero border


def test_decoder_collapses_are_exact():
    """The two re-associations the CUDA decoder uses against the literal oracle ops, fp64."""
    from strajnet_b200.weights import collapse_conv3d_811, fold_upconv_subpixel
    torch.manual_seed(0)
    # (8,1,1) Conv3D over an 8x-repeated tensor == per-waypoint 1x1 with summed taps
    k = torch.randn(8, 1, 1, 6, 5, dtype=torch.float64)
    b = torch.randn(5, dtype=torch.float64)
    r = torch.randn(2, 4, 4, 6, dtype=torch.float64)
    lit = O._conv3d_811(r[:, None].expand(-1, 8, -1, -1, -1), {"kernel": k, "bias": b}, "")
    weff = collapse_conv3d_811(k)
    fast = O.elu(torch.einsum("bhwc,tcd->bthwd", r, weff) + b)
    assert (lit - fast).abs().max() < 1e-12
    # nearest x2 + 3x3 SAME == four 2x2 sub-pixel convs on the zero-padded low-res input
    kk = torch.randn(3, 3, 4, 3, dtype=torch.float64)
    x = torch.randn(2, 5, 7, 4, dtype=torch.float64)
    lit = O.conv2d_nhwc(O._up2(x), kk, None, padding="same")
    f = fold_upconv_subpixel(kk)
    xp = torch.nn.functional.pad(x, (0, 0, 1, 1, 1, 1))
    out = torch.zeros(2, 10, 14, 3, dtype=torch.float64)
    for py in range(2):
        for px in range(2):
            acc = 0
            for a in range(2):
                for bb in range(2):
                    patch = xp[:, py + a: py + a + 5, px + bb: px + bb + 7]  # L[y-1+py+a, x-1+px+b]
                    acc = acc + patch @ f[py, px, a, bb]
            out[:, py::2, px::2] = acc
    assert (lit - out).abs().max() < 1e-12


def test_large_input_mode_shapes():
    w = O.make_weights(O.CFG512, seed=1)
    inp = O.make_inputs(1, 512, seed=1)
    res = O.encoder_forward(inp["ogm"], inp["map_img"], inp["flow"], w, O.CFG512, True)
    assert [tuple(r.shape) for r in res] == [(1, 4096, 96), (1, 4096, 96), (1, 1024, 192), (1, 256, 384)]


def test_bf16_storage_mode_is_scoped_and_bf16_sized():
    """O.bf16_storage(): layer-granularity model of the bf16 path's rounding points.  Outside the context the oracle is
    untouched (bit-identical results before / after); inside, a Swin block deviates by bf16 rounding noise, not more."""
    import torch
    from oracle import strajnet_oracle as O
    w = O.make_block_weights(96, 3, seed=3)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 16 * 16, 96, generator=g)
    y0 = O.swin_block(x, w, "", 16, 16, 3, 8, 4)
    with O.bf16_storage():
        y1 = O.swin_block(x, w, "", 16, 16, 3, 8, 4)
        assert O._EMULATE_BF16
    assert not O._EMULATE_BF16
    y2 = O.swin_block(x, w, "", 16, 16, 3, 8, 4)
    assert torch.equal(y0, y2)
    err = (y1 - y0).abs().max().item()
    assert 1e-4 < err < 5e-2 * y0.abs().max().item()
    assert torch.equal(y1, y1.to(torch.bfloat16).float())  # the block output is a stored tensor
