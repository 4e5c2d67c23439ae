"""CPU oracle for the STrajNet occupancy-flow forward path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``strajnet_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and there only as the
checker / the reported CPU baseline.

What it is: an op-for-op restatement (torch-CPU, channels-last, fp32 by default,
fp64 on request) of the *live* graph of ``STrajNet.call`` in the reference
(georgeliu233/STrajNet @ 21884df).  Every function cites the reference
file:line it follows (paths relative to /root/reference).

PARITY UNPINNED for the third-party arithmetic: TensorFlow/Keras and
``tensorflow_addons`` are not installable in this environment (no wheels, no
network), so the Keras layer defaults and ``tfa.layers.MultiHeadAttention``
are restated from their published upstream behaviour
and cannot be executed side by side.  What *is* pinned: the integer maps
(relative-position index, shift masks, window partition) are checked against
the reference's own pure-NumPy lines executed verbatim
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``), and window
attention is cross-checked against torchvision's independent implementation.

Layouts follow Keras: Dense.kernel [in,out]; Conv2D.kernel [kh,kw,in/groups,out];
Conv3D.kernel [kd,kh,kw,in,out]; Conv1D.kernel [k,in,out]; tfa-MHA kernels
[H,in,hs] / projection [H,hs,out].
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Weights = Dict[str, Tensor]

CFG256 = dict(input_size=(256, 256), window_size=8, embed_dim=96, depths=[2, 2, 2], num_heads=[3, 6, 12])
CFG512 = dict(input_size=(512, 512), window_size=8, embed_dim=96, depths=[2, 2, 2], num_heads=[3, 6, 12])
DECODER_CHANNELS = [48, 96, 128, 192, 384]  # modules.py:636
NUM_WAYPOINTS = 8


# --------------------------------------------------------------------------
# bf16 storage emulation (off by default: the oracle is the fp32 / fp64 restatement)
# --------------------------------------------------------------------------
# Inside `with bf16_storage():` the same graph is evaluated with the rounding points of the benchmarked CUDA path modelled
# at LAYER granularity: every Dense / Conv / tfa-MHA kernel is rounded to bf16 (the tensor-core copies), and every tensor
# that path stores in HBM or feeds to a tensor core as an operand (layer outputs after their activation, residual sums,
# q / k / v, attention outputs) is rounded to bf16; accumulation, LayerNorm, softmax and bias adds stay in the working
# precision.  It is NOT a bit-level model of the kernels (LayerNorm folding, sub-pixel tap folding, un-normalised bf16
# softmax weights are not modelled: tests/test_gpu_kernels.py does that per kernel); it removes the systematic part of the
# bf16-vs-fp32 gap so that the whole-forward gate can sit at rounding-noise level instead of at 8 % of the logit range.
_EMULATE_BF16 = False


class bf16_storage:
    def __enter__(self):
        global _EMULATE_BF16
        self.prev, _EMULATE_BF16 = _EMULATE_BF16, True
        return self

    def __exit__(self, *exc):
        global _EMULATE_BF16
        _EMULATE_BF16 = self.prev
        return False


def _st(x: Tensor) -> Tensor:
    """a stored activation / tensor-core operand of the bf16 path"""
    return x.to(torch.bfloat16).to(x.dtype) if _EMULATE_BF16 else x


# --------------------------------------------------------------------------
# elementwise / small helpers
# --------------------------------------------------------------------------
def gelu_tanh(x: Tensor) -> Tensor:
    """modules.py:18-29 (dup FG_MSA.py:7-18): x*0.5*(1+tanh(sqrt(2/pi)*(x+0.044715*x^3)))."""
    return x * (0.5 * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x * x * x))))


def layer_norm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float) -> Tensor:
    """Keras LayerNormalization(axis=-1): biased variance, gamma*(x-mu)/sqrt(var+eps)+beta."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * gamma + beta


def dense(x: Tensor, kernel: Tensor, bias: Optional[Tensor] = None, store: bool = True) -> Tensor:
    """store=False: the result is consumed in the producing epilogue (activation / residual) before it is stored"""
    y = x @ _st(kernel)
    y = y if bias is None else y + bias
    return _st(y) if store else y


def conv2d_nhwc(x: Tensor, kernel: Tensor, bias: Optional[Tensor], stride: int = 1,
                padding: str = "valid", groups: int = 1, store: bool = True) -> Tensor:
    """Keras Conv2D on [N,H,W,C] with kernel [kh,kw,cin/groups,cout]; 'same' = TF SAME (odd k, stride 1)."""
    w = _st(kernel).permute(3, 2, 0, 1)
    pad = 0
    if padding == "same":
        assert stride == 1 and kernel.shape[0] % 2 == 1
        pad = kernel.shape[0] // 2
    y = F.conv2d(x.permute(0, 3, 1, 2), w, bias, stride=stride, padding=pad, groups=groups)
    y = y.permute(0, 2, 3, 1)
    return _st(y) if store else y


# --------------------------------------------------------------------------
# Swin pieces (modules.py)
# --------------------------------------------------------------------------
def window_partition(x: Tensor, ws: int) -> Tensor:
    """modules.py:49-55. [B,H,W,C] -> [B*nW, ws, ws, C]."""
    B, H, W, C = x.shape
    x = x.reshape(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, ws, ws, C)


def window_reverse(windows: Tensor, ws: int, H: int, W: int, C: int) -> Tensor:
    """modules.py:58-63."""
    x = windows.reshape(-1, H // ws, W // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, H, W, C)


def relative_position_index(ws: int) -> np.ndarray:
    """modules.py:88-98, closed form: ((n//ws - m//ws + ws-1)*(2ws-1) + (n%ws - m%ws + ws-1))."""
    n = np.arange(ws * ws)
    dr = n[:, None] // ws - n[None, :] // ws + ws - 1
    dc = n[:, None] % ws - n[None, :] % ws + ws - 1
    return (dr * (2 * ws - 1) + dc).astype(np.int64)


def shift_attn_mask(H: int, W: int, ws: int, shift: int) -> np.ndarray:
    """modules.py:189-214. Returns float64 [nW, ws*ws, ws*ws] in {0,-100}."""
    img = np.zeros((1, H, W, 1))
    cnt = 0
    for h in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for w in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, h, w, :] = cnt
            cnt += 1
    mw = img.reshape(1, H // ws, ws, W // ws, ws, 1).transpose(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws)
    diff = mw[:, None, :] - mw[:, :, None]
    return np.where(diff != 0, -100.0, 0.0)


def window_attention(xw: Tensor, w: Weights, p: str, num_heads: int, ws: int,
                     mask: Optional[Tensor]) -> Tensor:
    """modules.py:103-134. xw [B_,N,C]; mask [nW,N,N] or None."""
    B_, N, C = xw.shape
    d = C // num_heads
    qkv = dense(xw, w[p + "qkv.kernel"], w[p + "qkv.bias"]).reshape(B_, N, 3, num_heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * (d ** -0.5)
    attn = q @ k.transpose(-1, -2)
    idx = torch.from_numpy(relative_position_index(ws).reshape(-1))
    bias = w[p + "relative_position_bias_table"][idx].reshape(N, N, num_heads).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.reshape(-1, nW, num_heads, N, N) + mask.to(attn.dtype)[None, :, None]
        attn = attn.reshape(-1, num_heads, N, N)
    attn = torch.softmax(attn, dim=-1)
    x = _st((attn @ v).permute(0, 2, 1, 3).reshape(B_, N, C))
    return dense(x, w[p + "proj.kernel"], w[p + "proj.bias"], store=False)


def swin_block(x: Tensor, w: Weights, p: str, H: int, W: int, num_heads: int, ws: int, shift: int) -> Tensor:
    """modules.py:220-262 (training=False; DropPath inert)."""
    B, L, C = x.shape
    assert L == H * W, "input feature has wrong size"
    if min(H, W) <= ws:  # modules.py:173-175
        shift, ws = 0, min(H, W)
    shortcut = x
    y = layer_norm(x, w[p + "norm1.gamma"], w[p + "norm1.beta"], 1e-5).reshape(B, H, W, C)
    if shift > 0:
        y = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
        mask = torch.from_numpy(shift_attn_mask(H, W, ws, shift))
    else:
        mask = None
    xw = window_partition(y, ws).reshape(-1, ws * ws, C)
    aw = window_attention(xw, w, p + "attn.", num_heads, ws, mask)
    y = window_reverse(aw.reshape(-1, ws, ws, C), ws, H, W, C)
    if shift > 0:
        y = torch.roll(y, shifts=(shift, shift), dims=(1, 2))
    x = _st(shortcut + y.reshape(B, L, C))
    h = layer_norm(x, w[p + "norm2.gamma"], w[p + "norm2.beta"], 1e-5)
    h = _st(gelu_tanh(dense(h, w[p + "mlp.fc1.kernel"], w[p + "mlp.fc1.bias"], store=False)))
    return _st(x + dense(h, w[p + "mlp.fc2.kernel"], w[p + "mlp.fc2.bias"], store=False))


def patch_merging(x: Tensor, w: Weights, p: str, H: int, W: int) -> Tensor:
    """modules.py:274-292."""
    B, L, C = x.shape
    assert L == H * W and H % 2 == 0 and W % 2 == 0
    x = x.reshape(B, H, W, C)
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], dim=-1)
    x = x.reshape(B, (H // 2) * (W // 2), 4 * C)
    x = layer_norm(x, w[p + "norm.gamma"], w[p + "norm.beta"], 1e-5)
    return dense(x, w[p + "reduction.kernel"])


def patch_embed(x: Tensor, w: Weights, p: str) -> Tensor:
    """modules.py:437-446: Conv2D k=4 s=4 VALID + bias -> flatten -> LN(1e-5)."""
    y = conv2d_nhwc(_st(x), w[p + "proj.kernel"], w[p + "proj.bias"], stride=4, store=False)
    B, Hp, Wp, E = y.shape
    return layer_norm(y.reshape(B, Hp * Wp, E), w[p + "norm.gamma"], w[p + "norm.beta"], 1e-5)


def basic_layer(x: Tensor, w: Weights, p: str, H: int, W: int, depth: int, heads: int, ws: int,
                downsample: bool) -> Tuple[Tensor, Tensor]:
    """modules.py:351-364."""
    for i in range(depth):
        x = swin_block(x, w, f"{p}blocks.{i}.", H, W, heads, ws, 0 if i % 2 == 0 else ws // 2)
    res = x
    if downsample:
        return patch_merging(x, w, p + "downsample.", H, W), res
    return x, x


def encoder_forward(ogm: Tensor, map_img: Tensor, flow: Tensor, w: Weights, cfg: dict,
                    large_input: bool) -> List[Tensor]:
    """modules.py:570-624 with sep_encode=flow_sep=use_flow=True (modules.py:782-785)."""
    p = "encoder."
    ws, E, depths, heads = cfg["window_size"], cfg["embed_dim"], cfg["depths"], cfg["num_heads"]
    S = cfg["input_size"][0]
    P = S // 4
    nl = len(depths)
    vec = ogm[..., 0]  # :572 (ogm[...,1] is never used, Q4)
    f = patch_embed(flow, w, p + "patch_embed_flow.")
    f = _st(layer_norm(f, w[p + "flow_norm.gamma"], w[p + "flow_norm.beta"], 1e-5))
    flow_x, flow_res = basic_layer(f, w, p + "flow_layer.", P, P, depths[0], heads[0], ws, nl > 1)
    if not large_input:
        x = patch_embed(vec, w, p + "patch_embed_vecicle.") + patch_embed(map_img, w, p + "patch_embed_map.")
    else:  # :582-587 (hard-coded 64/128, Q13)
        maps = patch_embed(map_img, w, p + "patch_embed_map.").reshape(-1, 64, 64, E)
        maps = F.pad(maps, (0, 0, 32, 32, 32, 32)).reshape(-1, 128 * 128, E)
        x = patch_embed(vec, w, p + "patch_embed_vecicle.") + maps
    x = _st(layer_norm(x, w[p + "all_patch_norm.gamma"], w[p + "all_patch_norm.beta"], 1e-5))
    res_list = []
    for i in range(nl):
        Hi = P // (2 ** i)
        x, res = basic_layer(x, w, f"{p}basic_layers.{i}.", Hi, Hi, depths[i], heads[i], ws, i < nl - 1)
        if i == nl - 1:
            res = res.reshape(-1, Hi, Hi, res.shape[-1])
        if i == 0:
            x = _st(x + flow_x)
            if large_input:
                flow_res = flow_res.reshape(-1, 128, 128, E)[:, 32:96, 32:96, :].reshape(-1, 64 * 64, 96)
            res_list.append(flow_res)
        if large_input:
            init_res = 128 // (2 ** i)
            dim = E * (2 ** i)
            crop = init_res // 2
            cb, ce = int(init_res * 0.25), int(init_res * 0.75)
            res = res.reshape(-1, init_res, init_res, dim)[:, cb:ce, cb:ce, :].reshape(-1, crop * crop, dim)
        res_list.append(res)
    return res_list


# --------------------------------------------------------------------------
# zero-border bilinear sampler (occu_metric.py:345-409 + tfa_image.py:87-173)
# --------------------------------------------------------------------------
def bilinear_sample_zero(image: Tensor, warp: Tensor) -> Tensor:
    """sample(image, warp, pixel_type=0): image [B,H,W,C], warp [B,...,2] (x=width, y=height) -> [B,...,C].

    pixel_type=0 is an int, not PixelType.HALF_INTEGER, so no -0.5 shift (occu_metric.py:394).
    Border ZERO: pad image by 1, warp+1 (:400-402). interpolate_bilinear(indexing='xy'):
    floor clamped to [0,size-2], alpha clamped to [0,1] (tfa_image.py:116-139).
    """
    B, H, W, C = image.shape
    img = F.pad(image, (0, 0, 1, 1, 1, 1))
    Hp, Wp = H + 2, W + 2
    wshape = warp.shape
    q = (warp + 1).reshape(B, -1, 2)
    qy, qx = q[..., 1], q[..., 0]
    fy = torch.clamp(torch.floor(qy), 0, Hp - 2)
    fx = torch.clamp(torch.floor(qx), 0, Wp - 2)
    ay = torch.clamp(qy - fy, 0, 1).unsqueeze(-1)
    ax = torch.clamp(qx - fx, 0, 1).unsqueeze(-1)
    iy, ix = fy.long(), fx.long()
    flat = img.reshape(B, Hp * Wp, C)

    def gather(yy, xx):
        lin = (yy * Wp + xx).unsqueeze(-1).expand(-1, -1, C)
        return torch.gather(flat, 1, lin)

    tl, tr = gather(iy, ix), gather(iy, ix + 1)
    bl, br = gather(iy + 1, ix), gather(iy + 1, ix + 1)
    top = ax * (tr - tl) + tl
    bot = ax * (br - bl) + bl
    out = ay * (bot - top) + top
    return out.reshape(*wshape[:-1], C)


# --------------------------------------------------------------------------
# FG-MSA (FG_MSA.py:106-183), n_heads = n_groups = 8, 48 channels each
# --------------------------------------------------------------------------
def fgmsa_forward(x: Tensor, w: Weights, p: str = "fg_msa_layer.", n_heads: int = 8, n_groups: int = 8,
                  fg: bool = True) -> Tuple[Tensor, Tensor, Optional[Tensor]]:
    B, H, W, C = x.shape
    hc = C // n_heads
    gc = C // n_groups
    gh = n_heads // n_groups
    q = conv2d_nhwc(x, w[p + "proj_q.kernel"], w[p + "proj_q.bias"])  # :109
    # _get_offset :84-92
    u = conv2d_nhwc(q, w[p + "conv_offset_0.kernel"], w[p + "conv_offset_0.bias"], padding="same", groups=n_groups)
    u = layer_norm(u.reshape(B, H * W, C), w[p + "conv_norm.gamma"], w[p + "conv_norm.beta"], 1e-3)
    u = gelu_tanh(u.reshape(B, H, W, C))
    u = u.reshape(B, H, W, n_groups, gc).permute(0, 3, 1, 2, 4).reshape(B * n_groups, H, W, gc)
    offset = conv2d_nhwc(u, w[p + "conv_offset_proj.kernel"], None)  # [B*G,H,W,2]
    Hk, Wk = H, W
    ns = Hk * Wk
    rng = torch.tensor([Hk / 2, Wk / 2], dtype=x.dtype).reshape(1, 1, 1, 2)  # :115
    offset = torch.tanh(offset) * rng
    flow_hidden = None
    if fg:  # :120-123, Conv2D on rank-5 input (Q11)
        to = offset.reshape(B, n_groups, Hk, Wk, 2)
        flow_hidden = dense(to, w[p + "conv_offset_proj2.kernel"][0, 0], w[p + "conv_offset_proj2.bias"])
    # _get_ref_points :95-104: tf.meshgrid default 'xy' => ref[i,j] = (j, i)
    ii, jj = torch.meshgrid(torch.arange(Hk), torch.arange(Wk), indexing="ij")
    ref = torch.stack((jj, ii), -1).to(x.dtype)  # [H,W,2]
    pos = offset + ref[None]  # :134
    # Q1: sampled features are discarded (:141-142); k, v come from x itself
    xs = x.reshape(B, ns, C)
    k = dense(xs, w[p + "proj_k.kernel"][0, 0], w[p + "proj_k.bias"])
    v = dense(xs, w[p + "proj_v.kernel"][0, 0], w[p + "proj_v.bias"])
    qh = q.reshape(B, H * W, n_heads, hc).permute(0, 2, 1, 3).reshape(B * n_heads, H * W, hc)
    kh = k.reshape(B, ns, n_heads, hc).permute(0, 2, 1, 3).reshape(B * n_heads, ns, hc)
    vh = v.reshape(B, ns, n_heads, hc).permute(0, 2, 1, 3).reshape(B * n_heads, ns, hc)
    attn = torch.einsum("bqc,bkc->bqk", qh, kh) * (hc ** -0.5)
    # rpe bias :150-172
    rpe = w[p + "rpe_table"]  # [2H-1, 2W-1, heads]
    rpe_b = rpe[None].expand(B, -1, -1, -1).reshape(B, 2 * H - 1, 2 * W - 1, n_groups, gh).permute(0, 3, 1, 2, 4)
    q_grid = ref[None].expand(B * n_groups, -1, -1, -1).reshape(B * n_groups, H * W, 2)
    disp = q_grid[:, :, None, :] - pos.reshape(B * n_groups, ns, 2)[:, None, :, :]
    disp = torch.stack((disp[..., 1], disp[..., 0]), -1)  # :160
    bias = bilinear_sample_zero(rpe_b.reshape(B * n_groups, 2 * H - 1, 2 * W - 1, gh), disp)
    bias = bias.reshape(B * n_groups, H * W, ns, gh).permute(0, 3, 1, 2).reshape(B * n_heads, H * W, ns)
    attn = torch.softmax(attn + bias, dim=2)
    out = torch.einsum("bkv,bvc->bck", attn, vh)  # [B*heads, hc, HW]
    out = out.reshape(B, C, H, W).permute(0, 2, 3, 1)
    y = conv2d_nhwc(out, w[p + "proj_out.kernel"], w[p + "proj_out.bias"])
    return y, pos.reshape(B, n_groups, Hk, Wk, 2), flow_hidden


# --------------------------------------------------------------------------
# tfa.layers.MultiHeadAttention restated [unpinned]
# --------------------------------------------------------------------------
def tfa_mha(query: Tensor, key: Tensor, value: Tensor, w: Weights, p: str, mask: Optional[Tensor]) -> Tensor:
    Wq, Wk, Wv = w[p + "query_kernel"], w[p + "key_kernel"], w[p + "value_kernel"]
    Wp, bp = w[p + "projection_kernel"], w[p + "projection_bias"]
    hs = Wq.shape[-1]
    Q = _st(torch.einsum("...ni,hio->...nho", query, _st(Wq)))
    K = _st(torch.einsum("...mi,hio->...mho", key, _st(Wk)))
    V = _st(torch.einsum("...mi,hio->...mho", value, _st(Wv)))
    Q = Q / math.sqrt(float(hs))
    logits = torch.einsum("...nho,...mho->...hnm", Q, K)
    if mask is not None:
        m = mask.to(torch.float32)
        if m.dim() != logits.dim():
            m = m.unsqueeze(-3)
        # the reference graph is fp32: l + (-1e10) rounds to -1e10, so fully masked rows become
        # uniform (Q7).  Do the add in fp32 even when the oracle runs in fp64.
        masked = (logits.to(torch.float32) + (-10e9)).to(query.dtype)
        logits = torch.where(m.bool().expand_as(logits), logits, masked)
    attn = torch.softmax(logits, dim=-1)
    out = _st(torch.einsum("...hnm,...mhi->...nhi", attn, V))
    return _st(torch.einsum("...nhi,hio->...no", out, _st(Wp)) + bp)


def elu(x: Tensor) -> Tensor:
    return F.elu(x)


def traj_encoder(inputs: Tensor, mask: Tensor, w: Weights, p: str) -> Tensor:
    """trajNet.py:38-48. inputs [B,11,8], mask [B,11] bool -> [B,384]."""
    m = mask.to(torch.int32)
    m2 = m[:, :, None] * m[:, None, :]
    nodes = elu(dense(inputs[:, :, :5], w[p + "node_feature.kernel"][0], w[p + "node_feature.bias"]))
    nodes = tfa_mha(nodes, nodes, nodes, w, p + "node_attention.", m2)
    nodes = nodes.max(dim=1).values  # GlobalMaxPooling1D, mask-unaware (Q6)
    vector = dense(inputs[:, 0, 5:], w[p + "vector_feature.kernel"])
    out = torch.cat([nodes, vector], dim=1)
    return elu(dense(out, w[p + "sublayer.kernel"], w[p + "sublayer.bias"]))


def cross_attention_block(query: Tensor, key: Tensor, mask: Tensor, w: Weights, p: str) -> Tensor:
    """trajNet.py:79-87 and :224-234 (same structure; no residual inside)."""
    v = tfa_mha(query, key, key, w, p + "mha.", mask)
    v = layer_norm(v, w[p + "norm1.gamma"], w[p + "norm1.beta"], 1e-3)
    v = _st(elu(dense(v, w[p + "FFN1.kernel"], w[p + "FFN1.bias"], store=False)))
    v = dense(v, w[p + "FFN2.kernel"], w[p + "FFN2.bias"])
    return layer_norm(v, w[p + "norm2.gamma"], w[p + "norm2.beta"], 1e-3)


def trajnet_forward(obs_traj: Tensor, occ_traj: Tensor, w: Weights, p: str) -> Tuple[Tensor, Tensor, Tensor]:
    """trajNet.py:125-187 live branch (no_attn=False, double_net=False)."""
    n_obs, n_occ = obs_traj.shape[1], occ_traj.shape[1]
    obs_mask = (obs_traj != 0)[:, :, :, 0]
    occ_mask = (occ_traj != 0)[:, :, :, 0]
    obs = torch.stack([traj_encoder(obs_traj[:, i], obs_mask[:, i], w, p + "traj_encoder.") for i in range(n_obs)], 1)
    occ = torch.stack([traj_encoder(occ_traj[:, i], occ_mask[:, i], w, p + "traj_encoder.") for i in range(n_occ)], 1)
    bi = torch.zeros(n_obs + n_occ, 2, dtype=obs.dtype)
    bi[:n_obs, 0] = 1
    bi[n_obs:, 1] = 1
    embed = dense(bi, w[p + "seg_embed.kernel"])[None].expand(obs.shape[0], -1, -1)
    c = (torch.cat([obs_mask, occ_mask], 1).to(torch.int32).sum(-1) != 0).to(torch.int32)
    actors = torch.cat([obs, occ], 1) * c[:, :, None].to(obs.dtype)
    query = actors + embed
    amask = c[:, :, None] * c[:, None, :]
    value = cross_attention_block(query, actors, amask, w, p + "cross_attention.")
    obs = obs + value[:, :n_obs]  # un-masked encoder output (Q8)
    occ = occ + value[:, n_obs:]
    obs = layer_norm(obs + embed[:, :n_obs], w[p + "obs_norm.gamma"], w[p + "obs_norm.beta"], 1e-3)
    occ = layer_norm(occ + embed[:, n_obs:], w[p + "occ_norm.gamma"], w[p + "occ_norm.beta"], 1e-3)
    return obs, occ, c


def trajnet_cross_attention(pic_encode: Tensor, obs_traj: Tensor, occ_traj: Tensor, w: Weights,
                            p: str = "trajnet_attn.") -> Tensor:
    """trajNet.py:284-319 live branch (actor_only=True, sep_actors=False)."""
    B, T, H, W, D = pic_encode.shape
    obs, occ, traj_mask = trajnet_forward(obs_traj, occ_traj, w, p + "traj_net.")
    flat = pic_encode.reshape(B, T, H * W, D)
    pic_mask = torch.ones(B, H * W, dtype=torch.int32)
    amask = pic_mask[:, :, None] * traj_mask[:, None, :]
    key = torch.cat([obs, occ], 1)
    res = []
    for t in range(T):
        o = cross_attention_block(flat[:, t], key, amask, w, f"{p}cross_attn_obs.{t}.")
        res.append(_st(o + flat[:, t]))
    return torch.stack(res, 1).reshape(B, T, H, W, D)


# --------------------------------------------------------------------------
# decoder (modules.py:739-772), literal formulation
# --------------------------------------------------------------------------
def _up2(x: Tensor) -> Tensor:
    """UpSampling3D(size=(1,2,2)) nearest on [N,H,W,C]."""
    return x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)


def _conv_elu(x5: Tensor, w: Weights, p: str, act: bool = True) -> Tensor:
    """Conv2D 3x3 SAME on rank-5 [B,T,H,W,C] (weights shared over B,T; Q11)."""
    B, T, H, W, C = x5.shape
    y = conv2d_nhwc(x5.reshape(B * T, H, W, C), w[p + "kernel"], w[p + "bias"], padding="same", store=False)
    if act:
        y = _st(elu(y))  # the heads (act=False) leave as fp32 logits
    return y.reshape(B, T, H, W, -1)


def _conv3d_811(r5: Tensor, w: Weights, p: str) -> Tensor:
    """Conv3D(kernel (8,1,1), SAME, ELU) on [B,8,H,W,Ci]; TF SAME for k=8: pad 3 before, 4 after."""
    k = w[p + "kernel"]  # [8,1,1,Ci,Co]
    x = r5.permute(0, 4, 1, 2, 3)  # [B,Ci,T,H,W]
    x = F.pad(x, (0, 0, 0, 0, 3, 4))
    y = F.conv3d(x, _st(k).permute(4, 3, 0, 1, 2), w[p + "bias"])
    return elu(y.permute(0, 2, 3, 4, 1))


def decoder_forward(x: Tensor, res_list: List[Tensor], w: Weights, p: str = "decoder.") -> Tensor:
    """x [B,8,16,16,384]; res_list = [flow_res, res0, res1, res2]. Flags from modules.py:800-801."""
    flow_res, res = res_list[0], res_list[1:]
    ind_list, reshape_dim = [1, 0], [32, 64]  # shallow_decode=1 (:718-719)
    flow_x = None
    for i in range(4):
        x5 = x
        B, T, H, W, C = x5.shape
        x = _conv_elu(_up2(x5.reshape(B * T, H, W, C)).reshape(B, T, 2 * H, 2 * W, C), w, f"{p}upconv_0s.{i}.")
        if i <= len(ind_list) - 1:
            r = res[ind_list[i]]
            r = r[:, None].expand(-1, 8, *r.shape[1:])  # tf.repeat 8x (:752)
            r = r.reshape(-1, 8, reshape_dim[i], reshape_dim[i], r.shape[-1])
            x = _st(x + _conv3d_811(r, w, f"{p}res_layer.{i}."))
        if i == len(ind_list) - 1:
            fr = flow_res.reshape(-1, 64, 64, 96)
            fr = fr[:, None].expand(-1, 8, -1, -1, -1)
            flow_x = _st(x + _conv3d_811(fr, w, p + "res_f."))
    occ = _conv_elu(x, w, p + "output_layer.", act=False)
    for j in range(2):
        B, T, H, W, C = flow_x.shape
        flow_x = _conv_elu(_up2(flow_x.reshape(B * T, H, W, C)).reshape(B, T, 2 * H, 2 * W, C), w, f"{p}upconv_f.{j}.")
    fl = _conv_elu(flow_x, w, p + "output_layer_f.", act=False)
    return torch.cat([occ, fl], dim=-1)


# --------------------------------------------------------------------------
# assembly (modules.py:815-839)
# --------------------------------------------------------------------------
def strajnet_forward(w: Weights, cfg: dict, ogm: Tensor, map_img: Tensor, obs: Tensor, occ: Tensor,
                     flow: Tensor, fg_msa: bool = True, fg: bool = True, large_ogm: bool = False,
                     return_intermediates: bool = False):
    res_list = encoder_forward(ogm, map_img, flow, w, cfg, large_ogm)
    q = res_list[-1]
    B = q.shape[0]
    ref = None
    inter = {"res_list": res_list}
    if fg_msa:
        q4 = q.reshape(-1, 16, 16, 384)
        y, pos, ref = fgmsa_forward(q4, w, fg=fg)
        inter.update(fg_y=y, fg_pos=pos, fg_hidden=ref)
        q = (y + q4).reshape(-1, 256, 384)
    else:
        q = q.reshape(-1, 256, 384)
    query = q[:, None].expand(-1, 8, -1, -1)
    if fg:
        query = ref.reshape(-1, 8, 256, 384) + query
    inter["query"] = query
    obs_value = trajnet_cross_attention(query.reshape(B, 8, 16, 16, 384), obs, occ, w)
    inter["obs_value"] = obs_value
    y = decoder_forward(obs_value, res_list, w)
    y = y.permute(0, 2, 3, 1, 4).reshape(-1, 256, 256, 32)
    if return_intermediates:
        return y, inter
    return y


# --------------------------------------------------------------------------
# seeded synthetic weights / inputs
# --------------------------------------------------------------------------
def _glorot(rng: np.random.Generator, shape, fan_in: int, fan_out: int) -> np.ndarray:
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def swin_block_weight_shapes(C: int, heads: int, ws: int = 8, mlp_ratio: float = 4.0):
    Hd = int(C * mlp_ratio)
    return {
        "norm1.gamma": (C,), "norm1.beta": (C,),
        "attn.qkv.kernel": (C, 3 * C), "attn.qkv.bias": (3 * C,),
        "attn.relative_position_bias_table": ((2 * ws - 1) ** 2, heads),
        "attn.proj.kernel": (C, C), "attn.proj.bias": (C,),
        "norm2.gamma": (C,), "norm2.beta": (C,),
        "mlp.fc1.kernel": (C, Hd), "mlp.fc1.bias": (Hd,),
        "mlp.fc2.kernel": (Hd, C), "mlp.fc2.bias": (C,),
    }


def weight_shapes(cfg: dict = CFG256, fg_msa: bool = True, fg: bool = True) -> Dict[str, tuple]:
    """Attribute path -> shape."""
    E, depths, heads, ws = cfg["embed_dim"], cfg["depths"], cfg["num_heads"], cfg["window_size"]
    s: Dict[str, tuple] = {}
    for name, cin in (("vecicle", 11), ("map", 3), ("flow", 2)):
        p = f"encoder.patch_embed_{name}."
        s[p + "proj.kernel"] = (4, 4, cin, E)
        s[p + "proj.bias"] = (E,)
        s[p + "norm.gamma"] = (E,)
        s[p + "norm.beta"] = (E,)
    for n in ("flow_norm", "all_patch_norm"):
        s[f"encoder.{n}.gamma"] = (E,)
        s[f"encoder.{n}.beta"] = (E,)

    def layer(p, C, h, depth, down):
        for i in range(depth):
            for k, v in swin_block_weight_shapes(C, h, ws).items():
                s[f"{p}blocks.{i}.{k}"] = v
        if down:
            s[p + "downsample.norm.gamma"] = (4 * C,)
            s[p + "downsample.norm.beta"] = (4 * C,)
            s[p + "downsample.reduction.kernel"] = (4 * C, 2 * C)

    layer("encoder.flow_layer.", E, heads[0], depths[0], len(depths) > 1)
    for i in range(len(depths)):
        layer(f"encoder.basic_layers.{i}.", E * 2 ** i, heads[i], depths[i], i < len(depths) - 1)
    if fg_msa:
        p = "fg_msa_layer."
        for n in ("q", "k", "v", "out"):
            s[f"{p}proj_{n}.kernel"] = (1, 1, 384, 384)
            s[f"{p}proj_{n}.bias"] = (384,)
        s[p + "conv_offset_0.kernel"] = (3, 3, 48, 384)
        s[p + "conv_offset_0.bias"] = (384,)
        s[p + "conv_norm.gamma"] = (384,)
        s[p + "conv_norm.beta"] = (384,)
        s[p + "conv_offset_proj.kernel"] = (1, 1, 48, 2)
        if fg:
            s[p + "conv_offset_proj2.kernel"] = (1, 1, 2, 384)
            s[p + "conv_offset_proj2.bias"] = (384,)
        s[p + "rpe_table"] = (31, 31, 8)
    p = "trajnet_attn.traj_net.traj_encoder."
    s[p + "node_feature.kernel"] = (1, 5, 64)
    s[p + "node_feature.bias"] = (64,)
    for n in ("query", "key", "value"):
        s[f"{p}node_attention.{n}_kernel"] = (4, 64, 64)
    s[p + "node_attention.projection_kernel"] = (4, 64, 320)
    s[p + "node_attention.projection_bias"] = (320,)
    s[p + "vector_feature.kernel"] = (3, 64)
    s[p + "sublayer.kernel"] = (384, 384)
    s[p + "sublayer.bias"] = (384,)
    p = "trajnet_attn.traj_net.cross_attention."
    for n in ("query", "key", "value"):
        s[f"{p}mha.{n}_kernel"] = (6, 384, 64)
    s[p + "mha.projection_kernel"] = (6, 64, 384)
    s[p + "mha.projection_bias"] = (384,)
    for n in ("norm1", "norm2"):
        s[f"{p}{n}.gamma"] = (384,)
        s[f"{p}{n}.beta"] = (384,)
    s[p + "FFN1.kernel"] = (384, 1536)
    s[p + "FFN1.bias"] = (1536,)
    s[p + "FFN2.kernel"] = (1536, 384)
    s[p + "FFN2.bias"] = (384,)
    for n in ("obs_norm", "occ_norm"):
        s[f"trajnet_attn.traj_net.{n}.gamma"] = (384,)
        s[f"trajnet_attn.traj_net.{n}.beta"] = (384,)
    s["trajnet_attn.traj_net.seg_embed.kernel"] = (2, 384)
    for t in range(8):
        p = f"trajnet_attn.cross_attn_obs.{t}."
        for n in ("query", "key", "value"):
            s[f"{p}mha.{n}_kernel"] = (3, 384, 42)
        s[p + "mha.projection_kernel"] = (3, 42, 128)
        s[p + "mha.projection_bias"] = (128,)
        s[p + "norm1.gamma"] = (128,)
        s[p + "norm1.beta"] = (128,)
        s[p + "FFN1.kernel"] = (128, 512)
        s[p + "FFN1.bias"] = (512,)
        s[p + "FFN2.kernel"] = (512, 384)
        s[p + "FFN2.bias"] = (384,)
        s[p + "norm2.gamma"] = (384,)
        s[p + "norm2.beta"] = (384,)
    dc = DECODER_CHANNELS
    chain = [(384, dc[3]), (dc[3], dc[2]), (dc[2], dc[1]), (dc[1], dc[0])]
    for i, (ci, co) in enumerate(chain):
        s[f"decoder.upconv_0s.{i}.kernel"] = (3, 3, ci, co)
        s[f"decoder.upconv_0s.{i}.bias"] = (co,)
    s["decoder.res_layer.0.kernel"] = (8, 1, 1, 192, 192)
    s["decoder.res_layer.0.bias"] = (192,)
    s["decoder.res_layer.1.kernel"] = (8, 1, 1, 96, 128)
    s["decoder.res_layer.1.bias"] = (128,)
    s["decoder.res_f.kernel"] = (8, 1, 1, 96, 128)
    s["decoder.res_f.bias"] = (128,)
    s["decoder.upconv_f.0.kernel"] = (3, 3, 128, 96)
    s["decoder.upconv_f.0.bias"] = (96,)
    s["decoder.upconv_f.1.kernel"] = (3, 3, 96, 48)
    s["decoder.upconv_f.1.bias"] = (48,)
    for n in ("output_layer", "output_layer_f"):
        s[f"decoder.{n}.kernel"] = (3, 3, 48, 2)
        s[f"decoder.{n}.bias"] = (2,)
    return s


def init_weight(rng: np.random.Generator, name: str, shape: tuple) -> np.ndarray:
    """Deliberately non-trivial init so indexing bugs are visible."""
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "gamma":
        return (1.0 + 0.1 * rng.standard_normal(shape)).astype(np.float32)
    if leaf == "beta":
        return (0.1 * rng.standard_normal(shape)).astype(np.float32)
    if leaf in ("bias", "projection_bias"):
        return rng.uniform(-0.1, 0.1, size=shape).astype(np.float32)
    if leaf == "relative_position_bias_table":
        return (0.02 * rng.standard_normal(shape)).astype(np.float32)
    if leaf == "rpe_table":
        return (0.5 * rng.standard_normal(shape)).astype(np.float32)
    if len(shape) == 2:
        return _glorot(rng, shape, shape[0], shape[1])
    if leaf.endswith("_kernel"):  # tfa MHA [H,in,out]
        return _glorot(rng, shape, shape[1], shape[2])
    rf = int(np.prod(shape[:-2]))
    return _glorot(rng, shape, rf * shape[-2], rf * shape[-1])


def make_weights(cfg: dict = CFG256, seed: int = 0, fg_msa: bool = True, fg: bool = True,
                 dtype=torch.float32) -> Weights:
    rng = np.random.Generator(np.random.PCG64(seed))
    return {k: torch.from_numpy(init_weight(rng, k, shp)).to(dtype) for k, shp in weight_shapes(cfg, fg_msa, fg).items()}


def make_block_weights(C: int, heads: int, seed: int = 0, ws: int = 8, prefix: str = "", dtype=torch.float32) -> Weights:
    rng = np.random.Generator(np.random.PCG64(seed))
    return {prefix + k: torch.from_numpy(init_weight(rng, k, shp)).to(dtype)
            for k, shp in swin_block_weight_shapes(C, heads, ws).items()}


def make_inputs(B: int, S: int = 256, seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Synthetic inputs with the value ranges of inference.py:84-96."""
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    ogm = (rng.random((B, S, S, 11, 2)) < 0.03).astype(np.float32)
    map_img = rng.integers(-128, 128, size=(B, 256, 256, 3)).astype(np.float32) / 256.0
    flow = ((rng.random((B, S, S, 2)) < 0.03) * rng.uniform(-20, 20, size=(B, S, S, 2))).astype(np.float32)

    def actors(n_max, lo):
        a = np.zeros((B, n_max, 11, 8), np.float32)
        for b in range(B):
            nv = int(rng.integers(lo, n_max + 1))
            for i in range(nv):
                steps = int(rng.integers(1, 12))
                xy = rng.uniform(-80, 80, size=(steps, 2))
                xy[xy == 0] = 1.0
                a[b, i, 11 - steps:, 0:2] = xy
                a[b, i, 11 - steps:, 2:4] = rng.normal(0, 5, size=(steps, 2))
                a[b, i, 11 - steps:, 4] = rng.uniform(-math.pi, math.pi, size=steps)
                a[b, i, :, 5 + int(rng.integers(0, 3))] = 1.0
        return a

    return {
        "ogm": torch.from_numpy(ogm).to(dtype), "map_img": torch.from_numpy(map_img).to(dtype),
        "flow": torch.from_numpy(flow).to(dtype),
        "obs": torch.from_numpy(actors(48, 4)).to(dtype), "occ": torch.from_numpy(actors(16, 0)).to(dtype),
        "mapt": torch.zeros(B, 256, 10, 7, dtype=dtype),
    }


def forward_from_inputs(w: Weights, cfg: dict, inp: Dict[str, Tensor], **kw):
    return strajnet_forward(w, cfg, inp["ogm"], inp["map_img"], inp["obs"], inp["occ"], inp["flow"], **kw)


# --------------------------------------------------------------------------
# serving-loop I/O around the model (inference.py:84-96, :124-136, :160-182)
# --------------------------------------------------------------------------
def decode_raw_inputs(ogm_bool: Tensor, map_int8: Tensor) -> Tuple[Tensor, Tensor]:
    """inference.py:91,93: ogm bool bytes -> float32; map int8 -> float32 / 256."""
    return (ogm_bool != 0).to(torch.float32), map_int8.to(torch.float32) / 256


def quantize_outputs(y: Tensor) -> Tensor:
    """inference.py:124-136 + :160-182 on the packed [B,256,256,32] logits -> uint8 [B,256,256,32]:
    channels k*4+{0,1}: round(sigmoid(x)*255) as uint8; k*4+{2,3}: clip(round(x),-128,127) as int8 (two's complement)."""
    y = y.to(torch.float32).reshape(*y.shape[:-1], 8, 4)
    occ = torch.round(torch.sigmoid(y[..., :2]) * 255).to(torch.uint8)
    flw = torch.clamp(torch.round(y[..., 2:]), -128, 127).to(torch.int8).view(torch.uint8)
    return torch.cat([occ, flw], dim=-1).reshape(*y.shape[:-2], 32)
