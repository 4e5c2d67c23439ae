"""GPU tests of the tcgen05 (tensor-core, bf16) kernels against a plain fp32 torch reference computed
from the same bf16-rounded operands, plus bf16-mode layer parity against the fp32 oracle.
Tolerance: bf16 output rounding (2^-9 relative) + fp32 accumulation-order noise."""
import pytest
import torch

from oracle import strajnet_oracle as O
from tests.util import max_abs, oracle_model, randn, sub

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sj():
    import strajnet_b200
    return strajnet_b200


def _tc_count():
    from strajnet_b200 import _lib
    return _lib.lib().sj_tc_launch_count(1)


def _bf(x):
    return x.to(torch.bfloat16).float()


@pytest.mark.parametrize("M,K,N,act", [
    (128, 64, 16, None), (128, 96, 96, None), (300, 384, 1152, None), (4096, 96, 384, "gelu"),
    (1000, 1536, 384, "elu"), (647, 128, 256, None), (16384, 192, 128, None), (64, 256, 320, None),
    (5, 384, 768, None), (2048, 512, 384, None), (40000, 96, 288, None),
])
def test_dense_bf16_tensor_core(sj, M, K, N, act):
    w = randn((K, N), 1, K ** -0.5)
    b = randn((N,), 2, 0.1)
    x = randn((M, K), 3)
    layer = sj.Dense(N, K, activation=act, dtype="bfloat16")
    layer.set_weights({"kernel": w, "bias": b})
    _tc_count()
    y = layer(x)
    assert _tc_count() == 1, "the tcgen05 kernel did not run (silent SIMT fallback)"
    ref = _bf(x) @ _bf(w) + b
    if act == "gelu":
        ref = O.gelu_tanh(ref)
    elif act == "elu":
        ref = O.elu(ref)
    err = max_abs(y, ref)
    tol = 2e-2 * max(1.0, ref.abs().max().item()) / 2
    assert err < tol, f"err {err} tol {tol}"


def test_mlp_bf16(sj):
    w = O.make_block_weights(96, 3, seed=5)
    m = sj.Mlp(96, 384, dtype="bfloat16")
    m.set_weights(sub(w, "mlp."))
    x = randn((2, 1000, 96), 6)
    ref = O.dense(O.gelu_tanh(O.dense(_bf(x), w["mlp.fc1.kernel"], w["mlp.fc1.bias"])), w["mlp.fc2.kernel"], w["mlp.fc2.bias"])
    _tc_count()
    y = m(x)
    assert _tc_count() == 2
    assert max_abs(y, ref) < 3e-2


@pytest.mark.parametrize("C,heads,H,B", [(96, 3, 64, 2), (192, 6, 32, 2), (384, 12, 16, 3)])
@pytest.mark.parametrize("shift", [0, 4])
def test_swin_block_bf16(sj, C, heads, H, B, shift):
    w = O.make_block_weights(C, heads, seed=11)
    blk = sj.SwinTransformerBlock(C, (H, H), heads, window_size=8, shift_size=shift, dtype="bfloat16")
    blk.set_weights(w)
    x = _bf(randn((B, H * H, C), 12))
    ref = O.swin_block(x, w, "", H, H, heads, 8, shift)
    _tc_count()
    y = blk(x)
    # C=96: fused window-MSA kernel (K1) + fused MLP kernel; C=192: K1 + fc1 + fc2; C=384: qkv, proj, fc1, fc2
    assert _tc_count() == {96: 2, 192: 3, 384: 4}[C]
    err = max_abs(y, ref)
    print(f"swin block bf16 C={C} shift={shift}: max abs err {err:.3e}")
    assert err < 6e-2


@pytest.mark.parametrize("C,H", [(96, 64), (192, 32)])
def test_patch_merging_bf16(sj, C, H):
    w = oracle_model()
    p = f"encoder.basic_layers.{0 if C == 96 else 1}.downsample."
    layer = sj.PatchMerging((H, H), C, dtype="bfloat16")
    layer.set_weights(sub(w, p))
    x = _bf(randn((2, H * H, C), 13))
    _tc_count()
    y = layer(x)
    assert _tc_count() == 1
    assert max_abs(y, O.patch_merging(x, w, p, H, H)) < 4e-2


@pytest.mark.parametrize("C,heads,masked", [(32, 2, False), (32, 1, True), (192, 6, True), (384, 12, False)])
def test_window_attention_bf16(sj, C, heads, masked):
    """Warp-MMA window attention core (attn_mma.cu): head dims 16 and 32, explicit [nW,64,64] mask."""
    w = O.make_block_weights(C, heads, seed=7)
    layer = sj.WindowAttention(C, (8, 8), heads, dtype="bfloat16")
    layer.set_weights(sub(w, "attn."))
    nW = 4
    x = _bf(randn((2 * nW, 64, C), 8))
    mask = torch.from_numpy(O.shift_attn_mask(16, 16, 8, 4)).float() if masked else None
    ref = O.window_attention(x, w, "attn.", heads, 8, mask)
    err = max_abs(layer(x, mask=mask), ref)
    print(f"window attention bf16 C={C} heads={heads}: max abs err {err:.3e}")
    assert err < 3e-2


@pytest.mark.parametrize("mode", ["normal", "no_occ", "all_padded"])
def test_traj_cross_attention_bf16_masks(sj, mode):
    """tfa mask semantics on the warp-MMA cores: partially and fully masked rows (uniform attention, Q7)."""
    from tests.test_gpu_parity import _traj_inputs
    w = oracle_model()
    layer = sj.TrajNetCrossAttention(dict(traj_heads=4, att_heads=6, out_dim=384, no_attn=False), pic_size=(16, 16),
                                     pic_dim=384, dtype="bfloat16")
    layer.set_weights(sub(w, "trajnet_attn."))
    pic = _bf(randn((2, 8, 16, 16, 384), 17))
    obs, occ = _traj_inputs(2, 3, mode)
    out = layer(pic, obs, occ, None, training=False)
    ref = O.trajnet_cross_attention(pic, obs, occ, w)
    err = max_abs(out, ref)
    rmax = ref.abs().max().item()
    print(f"traj cross attention bf16 ({mode}): max abs err {err:.3e}, max |ref| {rmax:.2f}")
    # bf16 chain through three LayerNorms (eps 1e-3): gated relative to the output range, like the whole model
    assert torch.isfinite(out).all() and err < 0.05 * rmax


def test_traj_cross_attention_bf16(sj):
    w = oracle_model()
    layer = sj.TrajNetCrossAttention(dict(traj_heads=4, att_heads=6, out_dim=384, no_attn=False), pic_size=(16, 16),
                                     pic_dim=384, dtype="bfloat16")
    layer.set_weights(sub(w, "trajnet_attn."))
    pic = _bf(randn((2, 8, 16, 16, 384), 17))
    inp = O.make_inputs(2, 256, seed=3)
    _tc_count()
    out = layer(pic, inp["obs"], inp["occ"], None, training=False)
    assert _tc_count() >= 10
    ref = O.trajnet_cross_attention(pic, inp["obs"], inp["occ"], w)
    err = max_abs(out, ref)
    print(f"traj cross attention bf16: max abs err {err:.3e}")
    assert err < 0.15


def test_fgmsa_bf16(sj):
    w = oracle_model()
    layer = sj.FGMSA((16, 16), (16, 16), 8, 48, n_groups=8, out_dim=384, fg=True, dtype="bfloat16")
    layer.set_weights(sub(w, "fg_msa_layer."))
    x = _bf(randn((2, 16, 16, 384), 16))
    y, pos, hid = layer(x, training=False)
    ry, rpos, rhid = O.fgmsa_forward(x, w)
    assert max_abs(pos, rpos) < 0.25 and max_abs(y, ry) < 0.1


@pytest.mark.parametrize("B", [1, 3])
def test_decoder_bf16(sj, B):
    """B = 3 gives the skip-add GEMMs several tiles per CTA (the residual prefetch crosses tile boundaries)."""
    w = oracle_model()
    dec = sj.Pyramid3DDecoder(None, (256, 256), use_pyramid=True, timestep_split=True, shallow_decode=1,
                              flow_sep_decode=True, conv_cnn=False, dtype="bfloat16")
    dec.set_weights(sub(w, "decoder."))
    x = _bf(randn((B, 8, 16, 16, 384), 18))
    res = [_bf(randn((B, 4096, 96), 19)), _bf(randn((B, 4096, 96), 20)), _bf(randn((B, 1024, 192), 21)),
           _bf(randn((B, 16, 16, 384), 22))]
    out = dec(x, training=False, res_list=res)
    ref = O.decoder_forward(x, res, w)
    err = max_abs(out, ref)
    print(f"decoder bf16 B={B}: max abs err {err:.3e}, max |ref| {ref.abs().max().item():.2f}")
    assert err < 0.05 * ref.abs().max().item()
