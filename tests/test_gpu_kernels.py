"""Tight per-kernel parity tests of the tcgen05 (bf16) decoder kernels.

Each kernel is called on its own through the C ABI and compared with a plain fp32 torch computation on the SAME
bf16-rounded operands (inputs, and the tensor-core copy of the weights exactly as the packer builds it), so the only
differences left are fp32 accumulation order and ONE rounding of the output:

    |y - ref| <= 2^-8 * |ref| + 1e-3        (bf16 outputs; fp32 outputs: 1e-3 abs)

asserted separately on the interior, the four image borders and the four corners (a wrong border tap of one sub-pixel
phase changes edge pixels by ~1/9 of their value: far outside this gate, invisible to a 5 %-of-range gate), at the real
decoder resolutions with >= 2 tiles per CTA.  Reference ops: modules.py:746-749 (up-sampling stage), :750-757 (skip
add), :767-770 (heads)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from oracle import strajnet_oracle as O
from tests.util import oracle_model, randn, sub

pytestmark = pytest.mark.gpu

REL, ABS = 2.0 ** -8, 1e-3


@pytest.fixture(scope="module")
def env():
    import strajnet_b200  # noqa: F401
    from strajnet_b200 import _lib, weights
    return _lib, weights, torch.device("cuda")


def _bf(x):
    return x.to(torch.bfloat16).float()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _regions(H, W):
    """name -> boolean [H,W] mask: interior, 4 borders (without corners), 4 corners."""
    m = {}
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    top, bot, lef, rig = yy == 0, yy == H - 1, xx == 0, xx == W - 1
    m["interior"] = ~(top | bot | lef | rig)
    m["top"], m["bottom"], m["left"], m["right"] = top & ~lef & ~rig, bot & ~lef & ~rig, lef & ~top & ~bot, rig & ~top & ~bot
    m["corner_tl"], m["corner_tr"], m["corner_bl"], m["corner_br"] = top & lef, top & rig, bot & lef, bot & rig
    return m


def _assert_tight(y, ref, what, rel=REL, abs_=ABS):
    """y, ref: [N,H,W,C] (cpu fp32).  Per-region check of |y-ref| <= rel*|ref| + abs_."""
    assert torch.isfinite(y).all(), f"{what}: non-finite output"
    excess = (y - ref).abs() - (rel * ref.abs() + abs_)
    N, H, W, _ = ref.shape
    for name, mask in _regions(H, W).items():
        e = excess[:, mask]
        worst = e.max().item()
        assert worst <= 0, (f"{what}: region '{name}' exceeds one output rounding by {worst:.3e} "
                            f"(max |err| there {(y - ref).abs()[:, mask].max().item():.3e})")


def subpixel_upconv_ref(x, w_tc, bias):
    """fp32 reference of nearest-x2 + 3x3 SAME conv + bias + ELU in the folded formulation the kernels execute, using the
    tensor-core weight copy itself: x [N,H,W,Ci] fp32 (bf16-valued), w_tc [4 phases][Co][4 taps * Ci] fp32 (bf16-valued).
    out[2y+py, 2x+px] = sum_ab Wf[py,px,a,b] . L[y-1+py+a, x-1+px+b]  (weights.fold_upconv_subpixel)."""
    N, H, W_, Ci = x.shape
    Co = w_tc.shape[1]
    xp = F.pad(x.permute(0, 3, 1, 2), (1, 1, 1, 1))                       # [N,Ci,H+2,W+2]
    out = torch.empty(N, Co, 2 * H, 2 * W_)
    for py in range(2):
        for px in range(2):
            k = w_tc[py * 2 + px].reshape(Co, 2, 2, Ci).permute(0, 3, 1, 2)  # [Co,Ci,a,b]
            c = F.conv2d(xp, k)                                           # [N,Co,H+1,W+1]
            out[:, :, py::2, px::2] = c[:, :, py:py + H, px:px + W_]
    return O.elu(out.permute(0, 2, 3, 1) + bias)


# (Cin, Cout, H, NB, kernel that must run): NB gives every persistent CTA >= 2 tiles
UPCONV_CASES = [
    (96, 48, 128, 3, "tc_upconv4"),     # dec.upconv3 / upconvf1: 384 tiles of 16x8 over 148 CTAs
    (128, 96, 64, 5, "tc_upconv1p"),    # dec.upconv2 / upconvf0
    (192, 128, 32, 8, "tc_upconv"),     # dec.upconv1
    (384, 192, 16, 16, "tc_upconv"),    # dec.upconv0
]


@pytest.mark.parametrize("Cin,Cout,H,NB,kernel", UPCONV_CASES)
def test_upconv_stage_tight(env, Cin, Cout, H, NB, kernel):
    _lib, weights, dev = env
    lib = _lib.lib()
    torch.backends.cudnn.allow_tf32 = False
    k = randn((3, 3, Cin, Cout), 31, (6.0 / (9 * Cin + 9 * Cout)) ** 0.5)
    b = randn((Cout,), 32, 0.1)
    x = _bf(randn((NB, H, H, Cin), 33))
    p = weights.Packer({"kernel": k, "bias": b}, dev, tc=True)
    lin = p._upconv("")
    w_tc = _bf(weights.fold_upconv_subpixel(k).reshape(4, 4 * Cin, Cout).transpose(1, 2).contiguous())
    xd = x.to(dev, torch.bfloat16)
    y = torch.empty(NB, 2 * H, 2 * H, Cout, dtype=torch.bfloat16, device=dev)
    lib.sj_tc_launch_count(1)
    _lib.check(lib.sj_upconv_fwd(xd.data_ptr(), y.data_ptr(), C.byref(lin), NB, H, Cin, Cout, _lib.SJ_BF16, _stream()),
               kernel)
    torch.cuda.synchronize()
    assert lib.sj_tc_launch_count(1) == 1, "the tcgen05 up-convolution did not run (silent fallback)"
    ref = subpixel_upconv_ref(x, w_tc, b)
    _assert_tight(y.float().cpu(), ref, f"{kernel} {Cin}->{Cout} @{H}")
    # and against the literal op with fp32 weights (reference formulation): only the bf16 rounding of the folded
    # weights is added (sqrt(4*Cin) terms of relative 2^-9)
    lit = O.elu(O.conv2d_nhwc(O._up2(x), k, b, padding="same"))
    assert (y.float().cpu() - lit).abs().max().item() < 1.5e-2 * max(1.0, lit.abs().max().item())


@pytest.mark.parametrize("mode", ["fp32", "quantised"])
def test_out_heads_tight(env, mode):
    """tc_outconv: both 3x3 48->2 heads + transpose (+ fused submission quantisation), B = 2 (512 tiles)."""
    _lib, weights, dev = env
    lib = _lib.lib()
    w = oracle_model()
    B = 2
    dw = sub(w, "decoder.")
    p = weights.Packer(dw, dev, tc=True)
    dec = p.decoder("")
    xo = _bf(randn((B * 8, 256, 256, 48), 41))
    xf = _bf(randn((B * 8, 256, 256, 48), 42))
    xod, xfd = xo.to(dev, torch.bfloat16), xf.to(dev, torch.bfloat16)
    layout = 2 if mode == "quantised" else 1
    out = torch.empty(B, 256, 256, 32, dtype=torch.uint8 if layout == 2 else torch.float32, device=dev)
    lib.sj_tc_launch_count(1)
    _lib.check(lib.sj_out_head_fwd(xod.data_ptr(), xfd.data_ptr(), out.data_ptr(), C.byref(dec), B, layout, _lib.SJ_BF16,
                                   _stream()), "tc_out_conv")
    torch.cuda.synchronize()
    assert lib.sj_tc_launch_count(1) >= 1
    # same operands: bf16-rounded head kernels (the tensor-core copy), fp32 biases
    ko, kf = _bf(dw["output_layer.kernel"]), _bf(dw["output_layer_f.kernel"])
    occ = O.conv2d_nhwc(xo, ko, dw["output_layer.bias"], padding="same")
    fl = O.conv2d_nhwc(xf, kf, dw["output_layer_f.bias"], padding="same")
    ref = torch.cat([occ, fl], -1).reshape(B, 8, 256, 256, 4).permute(0, 2, 3, 1, 4).reshape(B, 256, 256, 32)
    if layout == 1:
        _assert_tight(out.cpu(), ref, "tc_outconv fp32", rel=0.0, abs_=1e-3)
    else:
        # bytes may differ by one step only where the fp32 logit sits within 1e-3 of a rounding boundary
        q = out.cpu().to(torch.int16)
        qr = O.quantize_outputs(ref).to(torch.int16)
        lo = O.quantize_outputs(ref - 1e-3).to(torch.int16)
        hi = O.quantize_outputs(ref + 1e-3).to(torch.int16)
        ok = (q == qr) | (q == lo) | (q == hi)
        assert ok.all(), f"{(~ok).sum().item()} quantised bytes differ beyond a 1e-3 logit perturbation"
        assert (q != qr).float().mean().item() < 2e-3


@pytest.mark.parametrize("Cin,Cout,HW,B", [(192, 192, 1024, 3), (96, 128, 4096, 3)])
def test_res_add_tight(env, Cin, Cout, HW, B):
    """Grouped skip-add tc_gemm (collapsed (8,1,1) Conv3D + ELU + residual), several tiles per CTA."""
    _lib, weights, dev = env
    lib = _lib.lib()
    k = randn((8, 1, 1, Cin, Cout), 51, (6.0 / (8 * Cin + 8 * Cout)) ** 0.5)
    b = randn((Cout,), 52, 0.1)
    p = weights.Packer({"kernel": k, "bias": b}, dev, tc=True)
    lin = p._res("")
    skip = _bf(randn((B, HW, Cin), 53))
    src = _bf(randn((B, 8, HW, Cout), 54))
    sd, rd = skip.to(dev, torch.bfloat16), src.to(dev, torch.bfloat16)
    dst = torch.empty_like(rd)
    lib.sj_tc_launch_count(1)
    _lib.check(lib.sj_res_add_fwd(sd.data_ptr(), rd.data_ptr(), dst.data_ptr(), C.byref(lin), B, HW, Cin, Cout,
                                  _lib.SJ_BF16, _stream()), "res_add")
    torch.cuda.synchronize()
    assert lib.sj_tc_launch_count(1) == 1
    weff = _bf(weights.collapse_conv3d_811(k))                       # [8,Cin,Cout], bf16-valued like the tc copy
    ref = src + O.elu(torch.einsum("bnc,tcd->btnd", skip, weff) + b)
    y = dst.float().cpu()
    excess = (y - ref).abs() - (REL * ref.abs() + ABS)
    assert excess.max().item() <= 0, f"res_add {Cin}->{Cout}: exceeds one output rounding by {excess.max().item():.3e}"
    # in-place form used by the decoder (dst aliases src)
    _lib.check(lib.sj_res_add_fwd(sd.data_ptr(), rd.data_ptr(), rd.data_ptr(), C.byref(lin), B, HW, Cin, Cout,
                                  _lib.SJ_BF16, _stream()), "res_add in place")
    torch.cuda.synchronize()
    assert torch.equal(rd, dst)


def _tail_reference(x3, f3, dw, weights):
    """Fused tail with the kernel's own rounding points: bf16 x4 (tcgen05.st operand), fp16 projected columns Z,
    fp32 9-tap sum.  Returns ([B,256,256,32] logits, max |Z|)."""
    outs, zmax = [], 0.0
    for x, up, head in ((x3, "upconv_0s.3.", "output_layer."), (f3, "upconv_f.1.", "output_layer_f.")):
        k = dw[up + "kernel"]
        w_tc = _bf(weights.fold_upconv_subpixel(k).reshape(4, 4 * 96, 48).transpose(1, 2).contiguous())
        x4 = _bf(subpixel_upconv_ref(x, w_tc, dw[up + "bias"]))                     # [NB,256,256,48]
        kh = _bf(dw[head + "kernel"])                                               # [3,3,48,2]
        z = torch.einsum("nyxc,abco->nyxabo", x4, kh).to(torch.float16).float()     # [NB,256,256,3,3,2]
        zmax = max(zmax, z.abs().max().item())
        zp = F.pad(z, (0, 0, 0, 0, 0, 0, 1, 1, 1, 1))                               # pad x and y by 1
        acc = torch.zeros(x.shape[0], 256, 256, 2)
        for a in range(3):
            for b in range(3):
                acc = acc + zp[:, a:a + 256, b:b + 256, a, b]
        outs.append(acc + dw[head + "bias"])
    B = x3.shape[0] // 8
    y = torch.cat(outs, -1).reshape(B, 8, 256, 256, 4).permute(0, 2, 3, 1, 4).reshape(B, 256, 256, 32)
    return y, zmax


@pytest.mark.parametrize("mode", ["fp32", "quantised"])
def test_decoder_tail_fused_tight(env, mode):
    """tc_upconv4h (96->48 up-convolution + ELU + TS-form head projection) x2 + head_tapsum against an fp32 torch
    computation with the same three rounding points; 1024 tiles per launch (7 per CTA), borders checked separately."""
    _lib, weights, dev = env
    lib = _lib.lib()
    w = oracle_model()
    dw = sub(w, "decoder.")
    p = weights.Packer(dw, dev, tc=True)
    dec = p.decoder("")
    B = 1
    x3 = _bf(randn((B * 8, 128, 128, 96), 61))
    f3 = _bf(randn((B * 8, 128, 128, 96), 62))
    xd, fd = x3.to(dev, torch.bfloat16), f3.to(dev, torch.bfloat16)
    layout = 2 if mode == "quantised" else 1
    out = torch.empty(B, 256, 256, 32, dtype=torch.uint8 if layout == 2 else torch.float32, device=dev)
    n = lib.sj_decoder_tail_workspace_bytes(B, _lib.SJ_BF16)
    ws = torch.empty(n, dtype=torch.uint8, device=dev)
    lib.sj_tc_launch_count(1)
    _lib.check(lib.sj_decoder_tail_fwd(xd.data_ptr(), fd.data_ptr(), out.data_ptr(), C.byref(dec), B, layout, _lib.SJ_BF16,
                                       ws.data_ptr(), n, _stream()), "decoder tail")
    torch.cuda.synchronize()
    assert lib.sj_tc_launch_count(1) == 2, "the fused up-convolution + head kernel did not run"
    ref, zmax = _tail_reference(x3, f3, dw, weights)
    # one fp16 step of the largest projected column (rounding ties fall differently under another summation order)
    # + one bf16 step of x4 through one tap + fp32 noise
    tol = 2.0 ** -10 * zmax + 2e-3
    if layout == 1:
        _assert_tight(out.cpu(), ref, "fused decoder tail", rel=0.0, abs_=tol)
        print(f"fused decoder tail: max |err| {(out.cpu() - ref).abs().max().item():.3e} (tol {tol:.3e}, max |Z| {zmax:.2f})")
    else:
        q = out.cpu().to(torch.int16)
        qr = O.quantize_outputs(ref).to(torch.int16)
        lo, hi = O.quantize_outputs(ref - tol).to(torch.int16), O.quantize_outputs(ref + tol).to(torch.int16)
        ok = (q == qr) | (q == lo) | (q == hi)
        assert ok.all(), f"{(~ok).sum().item()} quantised bytes differ beyond a {tol:.1e} logit perturbation"


# ---- fused Swin kernels (K1 = tc_wmsa, fused MLP = tc_mlp96), isolated through SwinTransformerBlock -----------------
def _ln_stats(x):
    mu = x.mean(-1, keepdim=True)
    var = (x * x).mean(-1, keepdim=True) - mu * mu
    return mu, torch.rsqrt(var.clamp_min(0) + 1e-5)


def _wmsa_emulated(x, w, H, heads, shift):
    """x + proj(window_attention(norm1(x))) with the kernel's rounding points (tc_wmsa.cu): norm1 folded into the bf16
    qkv weights (rstd*(x.W' - mean*colsum) + bias'), q/k/v rounded to bf16 (q pre-scaled by d^-0.5 log2 e), exp2-domain
    softmax with bf16 un-normalised P and an fp32 row sum, O/sum rounded to bf16, bf16 proj weights."""
    B, L, C = x.shape
    d = C // heads
    g, bt = w["norm1.gamma"], w["norm1.beta"]
    wq = _bf(w["attn.qkv.kernel"] * g[:, None])                                 # [C,3C] folded, bf16-valued
    colsum = wq.sum(0)
    biasf = w["attn.qkv.bias"] + (bt[:, None].double() * w["attn.qkv.kernel"].double()).sum(0).float()
    mu, rstd = _ln_stats(x)
    qkv = rstd * (x @ wq - mu * colsum) + biasf                                  # [B,L,3C]
    y = qkv.reshape(B, H, H, 3 * C)
    if shift:
        y = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
    yw = O.window_partition(y, 8).reshape(-1, 64, 3, heads, d).permute(2, 0, 3, 1, 4)   # [3, nWin, h, 64, d]
    LOG2E = 1.4426950408889634
    q, k, v = _bf(yw[0] * (d ** -0.5 * LOG2E)), _bf(yw[1]), _bf(yw[2])
    s = q @ k.transpose(-1, -2)
    idx = torch.from_numpy(O.relative_position_index(8).reshape(-1))
    bias = w["attn.relative_position_bias_table"][idx].reshape(64, 64, heads).permute(2, 0, 1)
    s = s + bias.unsqueeze(0) * LOG2E
    if shift:
        mask = torch.from_numpy(O.shift_attn_mask(H, H, 8, shift)).float()
        nW = mask.shape[0]
        s = (s.reshape(-1, nW, heads, 64, 64) + (mask * LOG2E)[None, :, None]).reshape(-1, heads, 64, 64)
    p = torch.exp2(s - s.max(-1, keepdim=True).values)
    o = _bf((_bf(p) @ v) / p.sum(-1, keepdim=True))                              # [nWin,h,64,d]
    o = o.permute(0, 2, 1, 3).reshape(-1, 8, 8, C)
    o = O.window_reverse(o, 8, H, H, C)
    if shift:
        o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
    return x + o.reshape(B, L, C) @ _bf(w["attn.proj.kernel"]) + w["attn.proj.bias"]


@pytest.mark.parametrize("C,heads,B,H", [(96, 3, 20, 64), (192, 6, 18, 32)])
@pytest.mark.parametrize("shift", [0, 4])
def test_tc_wmsa_tight(env, shift, C, heads, B, H):
    """K1 alone: fc2 is zeroed so that the block returns x1 = x + attention exactly (the MLP half adds 0).  640 tiles at
    C = 96 (two CTAs per SM), 144 at C = 192 (layer 1 of the encoder: one CTA per SM, 6 heads streamed)."""
    import strajnet_b200 as sj
    _lib, _, _ = env
    w = O.make_block_weights(C, heads, seed=21)
    w["mlp.fc2.kernel"] = torch.zeros_like(w["mlp.fc2.kernel"])
    w["mlp.fc2.bias"] = torch.zeros_like(w["mlp.fc2.bias"])
    blk = sj.SwinTransformerBlock(C, (H, H), heads, window_size=8, shift_size=shift, dtype="bfloat16")
    blk.set_weights(w)
    x = _bf(randn((B, H * H, C), 22))
    _lib.lib().sj_tc_launch_count(1)
    y = blk(x).float().cpu()
    n_tc = _lib.lib().sj_tc_launch_count(1)
    assert n_tc == (2 if C == 96 else 3), "fused window-MSA + fused MLP (C = 96) or fc1 / fc2 GEMMs expected"
    ref = _wmsa_emulated(x, w, H, heads, shift)
    err = (y - ref).abs()
    tol = REL * ref.abs() + 4e-3   # one output rounding + rounding-tie flips of the bf16 q/k/v/P/O intermediates
    print(f"tc_wmsa C {C} shift {shift}: max |err| {err.max().item():.3e}")
    assert (err <= tol).all(), f"tc_wmsa C {C} shift {shift}: max excess {(err - tol).max().item():.3e}"
    # against the plain fp32 oracle block: bf16 operand rounding only
    lit = O.swin_block(x, w, "", H, H, heads, 8, shift)
    assert (y - lit).abs().max().item() < 3e-2


def test_tc_mlp96_tight(env):
    """Fused MLP alone: proj is zeroed so that x1 = x.  norm2 folded into bf16 fc1, tanh-GELU, bf16 hidden, bf16 fc2."""
    import strajnet_b200 as sj
    _lib, _, _ = env
    B, H, C, heads = 20, 64, 96, 3
    w = O.make_block_weights(C, heads, seed=23)
    w["attn.proj.kernel"] = torch.zeros_like(w["attn.proj.kernel"])
    w["attn.proj.bias"] = torch.zeros_like(w["attn.proj.bias"])
    blk = sj.SwinTransformerBlock(C, (H, H), heads, window_size=8, shift_size=0, dtype="bfloat16")
    blk.set_weights(w)
    x = _bf(randn((B, H * H, C), 24))
    y = blk(x).float().cpu()
    g, bt = w["norm2.gamma"], w["norm2.beta"]
    w1 = _bf(w["mlp.fc1.kernel"] * g[:, None])
    b1 = w["mlp.fc1.bias"] + (bt[:, None].double() * w["mlp.fc1.kernel"].double()).sum(0).float()
    mu, rstd = _ln_stats(x)
    h = _bf(O.gelu_tanh(rstd * (x @ w1 - mu * w1.sum(0)) + b1))
    ref = x + h @ _bf(w["mlp.fc2.kernel"]) + w["mlp.fc2.bias"]
    err = (y - ref).abs()
    tol = REL * ref.abs() + 4e-3   # + tanh.approx (2^-11) through fc2 and rounding-tie flips of the bf16 hidden tile
    print(f"tc_mlp96: max |err| {err.max().item():.3e}")
    assert (err <= tol).all(), f"tc_mlp96: max excess {(err - tol).max().item():.3e}"


@pytest.mark.parametrize("B,inplace", [(3, False), (5, True)])
def test_res_add2_fused_tight(env, B, inplace):
    """tc_resadd2: both 64x64 skip connections (res_layer[1] on res0, res_f on flow_res) in one kernel, against fp32 torch
    on the same bf16 operands with the same two rounding points (x stored as bf16, fx = stored x + ...).  B = 3 / 5 give
    the 18 CTAs of a waypoint uneven tile counts; the in-place form (dst_a aliasing src) is what the decoder uses."""
    _lib, weights, dev = env
    lib = _lib.lib()
    HW, Cin, Cout = 4096, 96, 128
    ka = randn((8, 1, 1, Cin, Cout), 71, (6.0 / (8 * Cin + 8 * Cout)) ** 0.5)
    kb = randn((8, 1, 1, Cin, Cout), 72, (6.0 / (8 * Cin + 8 * Cout)) ** 0.5)
    ba, bb = randn((Cout,), 73, 0.1), randn((Cout,), 74, 0.1)
    pa = weights.Packer({"kernel": ka, "bias": ba}, dev, tc=True)
    pb = weights.Packer({"kernel": kb, "bias": bb}, dev, tc=True)
    la, lb = pa._res(""), pb._res("")
    sa, sb = _bf(randn((B, HW, Cin), 75)), _bf(randn((B, HW, Cin), 76))
    src = _bf(randn((B, 8, HW, Cout), 77))
    sad, sbd, srcd = (t.to(dev, torch.bfloat16) for t in (sa, sb, src))
    dst_a = srcd if inplace else torch.empty_like(srcd)
    dst_b = torch.empty_like(srcd)
    lib.sj_tc_launch_count(1)
    _lib.check(lib.sj_res_add2_fwd(sad.data_ptr(), sbd.data_ptr(), srcd.data_ptr(), dst_a.data_ptr(), dst_b.data_ptr(),
                                   C.byref(la), C.byref(lb), B, HW, Cin, Cout, _lib.SJ_BF16, _stream()), "res_add2")
    torch.cuda.synchronize()
    assert lib.sj_tc_launch_count(1) == 1, "the fused skip-add kernel did not run"
    wa, wb = _bf(weights.collapse_conv3d_811(ka)), _bf(weights.collapse_conv3d_811(kb))
    ref_a = src + O.elu(torch.einsum("bnc,tcd->btnd", sa, wa) + ba)
    ya = dst_a.float().cpu()
    ex = (ya - ref_a).abs() - (REL * ref_a.abs() + ABS)
    assert ex.max().item() <= 0, f"res_add2 x: exceeds one output rounding by {ex.max().item():.3e}"
    # the flow branch adds onto the STORED x (bf16), exactly as two separate launches would
    ref_b = ya + O.elu(torch.einsum("bnc,tcd->btnd", sb, wb) + bb)
    yb = dst_b.float().cpu()
    ex = (yb - ref_b).abs() - (REL * ref_b.abs() + ABS)
    assert ex.max().item() <= 0, f"res_add2 fx: exceeds one output rounding by {ex.max().item():.3e}"


def _pe_weights(cin, seed):
    return {"proj.kernel": randn((4, 4, cin, 96), seed, (1.0 / (16 * cin)) ** 0.5), "proj.bias": randn((96,), seed + 1, 0.1),
            "norm.gamma": 1 + randn((96,), seed + 2, 0.1), "norm.beta": randn((96,), seed + 3, 0.1)}


def _pe_conv(x, w):
    """4x4 / stride 4 conv of bf16-rounded pixels with the bf16-rounded kernel (the tensor-core operands), fp32 LN."""
    y = O.conv2d_nhwc(_bf(x), _bf(w["proj.kernel"]), w["proj.bias"], stride=4)
    B, Hp, Wp, E = y.shape
    return O.layer_norm(y.reshape(B, Hp * Wp, E), w["norm.gamma"], w["norm.beta"], 1e-5).reshape(B, Hp, Wp, E)


@pytest.mark.parametrize("case", ["flow", "raster_f32", "raster_raw", "raster_plane", "raster_512"])
def test_patch_embed_fused_tight(env, case):
    """tc_patch_embed: im2col in shared memory + tcgen05 projections + the LayerNorms and the sum in the epilogue, against
    fp32 torch on the same bf16 operands with one output rounding (modules.py:572-587, :602; :576-577 for `flow`).
    raster_raw feeds the record's own bool / int8 bytes, raster_plane the [.., 11] vehicle plane alone, raster_512 the
    512-input mode where the map only covers the centre 64 x 64 tokens."""
    _lib, weights, dev = env
    lib = _lib.lib()
    B = 3
    S = 512 if case == "raster_512" else 256
    P = S // 4
    nf = {"gamma": 1 + randn((96,), 90, 0.1), "beta": randn((96,), 91, 0.1)}
    pk_n = weights.Packer({"n.gamma": nf["gamma"], "n.beta": nf["beta"]}, dev, tc=True)
    norm_f = pk_n.norm("n.")
    if case == "flow":
        w0 = _pe_weights(2, 80)
        img0 = randn((B, S, S, 2), 81)
        x0, t0, es0, cin0 = img0, _lib.SJ_IN_F32, 1, 2
        d0 = img0.to(dev)
        w1 = img1 = d1 = None
    else:
        w0, w1 = _pe_weights(11, 82), _pe_weights(3, 86)
        if case == "raster_raw":
            ogm = (randn((B, S, S, 11, 2), 83) > 0.5)
            d0, t0, es0 = ogm.view(torch.uint8).to(dev), _lib.SJ_IN_U8, 2
            x0 = ogm[..., 0].float()
            m8 = (randn((B, 256, 256, 3), 84) * 60).clamp(-128, 127).to(torch.int8)
            d1, t1, img1 = m8.to(dev), _lib.SJ_IN_I8_DIV256, m8.float() / 256.0
        else:
            ogm = randn((B, S, S, 11, 2), 83)
            if case == "raster_plane":
                d0, es0 = ogm[..., 0].contiguous().to(dev), 1
            else:
                d0, es0 = ogm.to(dev), 2
            t0, x0 = _lib.SJ_IN_F32, ogm[..., 0]
            img1 = randn((B, 256, 256, 3), 84)
            d1, t1 = img1.to(dev), _lib.SJ_IN_F32
        cin0 = 11
    pk0 = weights.Packer(w0, dev, tc=True)
    pe0 = pk0.patch_embed("")
    pe1 = None
    if w1 is not None:
        pk1 = weights.Packer(w1, dev, tc=True)
        pe1 = pk1.patch_embed("")
    y = torch.empty(B, P * P, 96, dtype=torch.bfloat16, device=dev)
    mean = torch.empty(B * P * P, dtype=torch.float32, device=dev)
    rstd = torch.empty_like(mean)
    pad1 = (P - 64) // 2
    lib.sj_tc_launch_count(1)
    _lib.check(lib.sj_patch_embed_sum_fwd(d0.data_ptr(), t0, S, cin0, es0, C.byref(pe0),
                                          d1.data_ptr() if d1 is not None else None, t1 if d1 is not None else 0, 256, 3,
                                          C.byref(pe1) if pe1 is not None else None, pad1, C.byref(norm_f), B, y.data_ptr(),
                                          mean.data_ptr(), rstd.data_ptr(), _stream()), "patch_embed_sum")
    torch.cuda.synchronize()
    assert lib.sj_tc_launch_count(1) == 1, "the fused patch-embedding kernel did not run"
    t = _pe_conv(x0, w0)
    if w1 is not None:
        m = _pe_conv(img1, w1)
        t = t.clone()
        t[:, pad1:pad1 + 64, pad1:pad1 + 64] += m
    ref = O.layer_norm(t, nf["gamma"], nf["beta"], 1e-5)
    yc = y.float().cpu().reshape(B, P, P, 96)
    _assert_tight(yc, ref, f"patch_embed {case}", abs_=2e-3)
    mu, rs = _ln_stats(yc.reshape(-1, 96))
    assert (mean.cpu() - mu.reshape(-1)).abs().max().item() < 1e-4
    assert ((rstd.cpu() - rs.reshape(-1)).abs() / rs.reshape(-1)).max().item() < 1e-3
