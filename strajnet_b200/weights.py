"""Host-side weight preparation: Keras-layout parameters (attribute paths of SURVEY App. B)
-> the device layouts the C ABI expects (include/strajnet_b200.h), plus default initialisers.

All algebra here is exact re-association of the reference graph, done once at load time:
  * tfa MultiHeadAttention kernels [H,in,hs] -> [in, H*hs]            (trajNet.py:33,71,195)
  * (8,1,1) Conv3D over an 8x-repeated tensor -> 8 per-waypoint 1x1   (modules.py:709-717,752)
  * nearest-x2 upsample + 3x3 SAME conv -> four 2x2 sub-pixel convs   (modules.py:746-749)
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib as L

Tensor = torch.Tensor


# ------------------------------------------------------------------------------------------------
# pure algebra (device agnostic; unit-tested on CPU against the oracle)
# ------------------------------------------------------------------------------------------------
def collapse_conv3d_811(kernel: Tensor) -> Tensor:
    """[8,1,1,Ci,Co] -> [8,Ci,Co] with W_eff[t] = sum_{k: 0<=t+k-3<=7} W[k] (TF SAME pads 3 before, 4 after)."""
    k = kernel[:, 0, 0].to(torch.float64)
    out = torch.zeros_like(k)
    for t in range(8):
        for kk in range(8):
            if 0 <= t + kk - 3 <= 7:
                out[t] += k[kk]
    return out.to(kernel.dtype)


# taps of the 3-wide kernel that land on low-res offset a (0: previous/current, 1: current/next) for phase p
_FOLD = {(0, 0): (0,), (0, 1): (1, 2), (1, 0): (0, 1), (1, 1): (2,)}


def fold_upconv_subpixel(kernel: Tensor) -> Tensor:
    """[3,3,Ci,Co] -> [2,2,2,2,Ci,Co] indexed [py,px,a,b]: out[2y+py,2x+px] = sum_ab Wf[py,px,a,b] . L[y-1+py+a, x-1+px+b]."""
    k = kernel.to(torch.float64)
    ci, co = k.shape[2], k.shape[3]
    out = torch.zeros(2, 2, 2, 2, ci, co, dtype=torch.float64, device=k.device)
    for py in range(2):
        for px in range(2):
            for a in range(2):
                for b in range(2):
                    for dy in _FOLD[(py, a)]:
                        for dx in _FOLD[(px, b)]:
                            out[py, px, a, b] += k[dy, dx]
    return out.to(kernel.dtype)


def tfa_in_kernel(k: Tensor, pad_to: Optional[int] = None) -> Tensor:
    """tfa query/key/value kernel [H,in,hs] -> [in, H*hs] (optionally zero-padded columns)."""
    H, I, O = k.shape
    m = k.permute(1, 0, 2).reshape(I, H * O)
    if pad_to and pad_to > H * O:
        m = torch.cat([m, m.new_zeros(I, pad_to - H * O)], 1)
    return m.contiguous()


def tfa_out_kernel(k: Tensor, pad_to: Optional[int] = None) -> Tensor:
    """tfa projection kernel [H,hs,out] -> [H*hs, out] (optionally zero-padded rows)."""
    H, S, O = k.shape
    m = k.reshape(H * S, O)
    if pad_to and pad_to > H * S:
        m = torch.cat([m, m.new_zeros(pad_to - H * S, O)], 0)
    return m.contiguous()


# ------------------------------------------------------------------------------------------------
# default initialisation (Keras defaults: glorot_uniform kernels, zero biases, LN gamma=1/beta=0;
# relative_position_bias_table zeros (modules.py:86); rpe_table trunc-normal 0.01 (FG_MSA.py:72))
# ------------------------------------------------------------------------------------------------
def default_init(shapes: Dict[str, tuple], seed: int = 0) -> Dict[str, Tensor]:
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    for name, shp in shapes.items():
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "gamma":
            a = np.ones(shp, np.float32)
        elif leaf in ("beta", "bias", "projection_bias", "relative_position_bias_table"):
            a = np.zeros(shp, np.float32)
        elif leaf == "rpe_table":
            a = np.clip(rng.standard_normal(shp), -2, 2).astype(np.float32) * 0.01
        else:
            if len(shp) == 2:
                fi, fo = shp
            elif leaf.endswith("_kernel"):
                fi, fo = shp[1], shp[2]
            else:
                rf = int(np.prod(shp[:-2]))
                fi, fo = rf * shp[-2], rf * shp[-1]
            lim = math.sqrt(6.0 / (fi + fo))
            a = rng.uniform(-lim, lim, size=shp).astype(np.float32)
        out[name] = torch.from_numpy(a)
    return out


def swin_block_shapes(C: int, heads: int, ws: int = 8, mlp_ratio: float = 4.0) -> Dict[str, tuple]:
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


def patch_merging_shapes(C: int) -> Dict[str, tuple]:
    return {"norm.gamma": (4 * C,), "norm.beta": (4 * C,), "reduction.kernel": (4 * C, 2 * C)}


def patch_embed_shapes(cin: int, E: int) -> Dict[str, tuple]:
    return {"proj.kernel": (4, 4, cin, E), "proj.bias": (E,), "norm.gamma": (E,), "norm.beta": (E,)}


def basic_layer_shapes(C: int, heads: int, depth: int, down: bool, ws: int = 8) -> Dict[str, tuple]:
    s = {}
    for i in range(depth):
        for k, v in swin_block_shapes(C, heads, ws).items():
            s[f"blocks.{i}.{k}"] = v
    if down:
        for k, v in patch_merging_shapes(C).items():
            s[f"downsample.{k}"] = v
    return s


def encoder_shapes(E: int, depths: List[int], heads: List[int], ws: int = 8) -> Dict[str, tuple]:
    s = {}
    for name, cin in (("vecicle", 11), ("map", 3), ("flow", 2)):
        for k, v in patch_embed_shapes(cin, E).items():
            s[f"patch_embed_{name}.{k}"] = v
    for n in ("flow_norm", "all_patch_norm"):
        s[f"{n}.gamma"] = (E,)
        s[f"{n}.beta"] = (E,)
    nl = len(depths)
    for k, v in basic_layer_shapes(E, heads[0], depths[0], nl > 1, ws).items():
        s[f"flow_layer.{k}"] = v
    for i in range(nl):
        for k, v in basic_layer_shapes(E * 2 ** i, heads[i], depths[i], i < nl - 1, ws).items():
            s[f"basic_layers.{i}.{k}"] = v
    return s


def fgmsa_shapes(fg: bool = True) -> Dict[str, tuple]:
    s = {}
    for n in ("q", "k", "v", "out"):
        s[f"proj_{n}.kernel"] = (1, 1, 384, 384)
        s[f"proj_{n}.bias"] = (384,)
    s["conv_offset_0.kernel"] = (3, 3, 48, 384)
    s["conv_offset_0.bias"] = (384,)
    s["conv_norm.gamma"] = (384,)
    s["conv_norm.beta"] = (384,)
    s["conv_offset_proj.kernel"] = (1, 1, 48, 2)
    if fg:
        s["conv_offset_proj2.kernel"] = (1, 1, 2, 384)
        s["conv_offset_proj2.bias"] = (384,)
    s["rpe_table"] = (31, 31, 8)
    return s


def fg_conv_fragments(k: Tensor) -> Tensor:
    """conv_offset_0 kernel [3,3,48,384] (grouped 3x3, 8 groups of 48 -> 48) in the order the mma.sync kernel reads it
    (fg_offset_mma.cu): [group][tap][k-step][n-pair][lane][8], lane = 4*g + t, n-tile nt = 2*pair + (e >> 2), element
    e & 3 = k[tap][16*ks + 2*t + (0, 1, 8, 9)[e & 3]][48*group + 8*nt + g] -- one 16-byte load per lane and n-pair."""
    w = k.reshape(9, 48, 384)
    grp, tap, ks, pair, g, t, e = torch.meshgrid(torch.arange(8), torch.arange(9), torch.arange(3), torch.arange(3),
                                                 torch.arange(8), torch.arange(4), torch.arange(8), indexing="ij")
    kk = 16 * ks + 2 * t + torch.tensor([0, 1, 8, 9])[e & 3]
    n = 48 * grp + 8 * (2 * pair + (e >> 2)) + g
    return w[tap, kk, n].contiguous()          # [8,9,3,3,8,4,8]: (g, t) flatten to lane = 4*g + t


def traj_shapes() -> Dict[str, tuple]:
    s = {}
    p = "traj_net.traj_encoder."
    s[p + "node_feature.kernel"] = (1, 5, 64)
    s[p + "node_feature.bias"] = (64,)
    for n in ("query", "key", "value"):
        s[f"{p}node_attention.{n}_kernel"] = (4, 64, 64)
    s[p + "node_attention.projection_kernel"] = (4, 64, 320)
    s[p + "node_attention.projection_bias"] = (320,)
    s[p + "vector_feature.kernel"] = (3, 64)
    s[p + "sublayer.kernel"] = (384, 384)
    s[p + "sublayer.bias"] = (384,)
    p = "traj_net.cross_attention."
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
        s[f"traj_net.{n}.gamma"] = (384,)
        s[f"traj_net.{n}.beta"] = (384,)
    s["traj_net.seg_embed.kernel"] = (2, 384)
    for t in range(8):
        p = f"cross_attn_obs.{t}."
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
    return s


def decoder_shapes() -> Dict[str, tuple]:
    s = {}
    for i, (ci, co) in enumerate([(384, 192), (192, 128), (128, 96), (96, 48)]):
        s[f"upconv_0s.{i}.kernel"] = (3, 3, ci, co)
        s[f"upconv_0s.{i}.bias"] = (co,)
    s["res_layer.0.kernel"] = (8, 1, 1, 192, 192)
    s["res_layer.0.bias"] = (192,)
    s["res_layer.1.kernel"] = (8, 1, 1, 96, 128)
    s["res_layer.1.bias"] = (128,)
    s["res_f.kernel"] = (8, 1, 1, 96, 128)
    s["res_f.bias"] = (128,)
    s["upconv_f.0.kernel"] = (3, 3, 128, 96)
    s["upconv_f.0.bias"] = (96,)
    s["upconv_f.1.kernel"] = (3, 3, 96, 48)
    s["upconv_f.1.bias"] = (48,)
    for n in ("output_layer", "output_layer_f"):
        s[f"{n}.kernel"] = (3, 3, 48, 2)
        s[f"{n}.bias"] = (2,)
    return s


def model_shapes(cfg: dict, fg_msa: bool, fg: bool) -> Dict[str, tuple]:
    s = {}
    for k, v in encoder_shapes(cfg["embed_dim"], cfg["depths"], cfg["num_heads"], cfg["window_size"]).items():
        s["encoder." + k] = v
    if fg_msa:
        for k, v in fgmsa_shapes(fg).items():
            s["fg_msa_layer." + k] = v
    for k, v in traj_shapes().items():
        s["trajnet_attn." + k] = v
    for k, v in decoder_shapes().items():
        s["decoder." + k] = v
    return s


# ------------------------------------------------------------------------------------------------
# struct builders: every tensor handed to the C ABI is fp32/bf16, contiguous, on `device`
# ------------------------------------------------------------------------------------------------
class Packer:
    """Moves tensors to the device in the layout the kernels read and keeps them alive."""

    def __init__(self, weights: Dict[str, Tensor], device, tc: bool):
        self.w = weights
        self.device = torch.device(device)
        self.tc = tc  # also build the bf16 tensor-core copies
        self.keep: List[object] = []

    def dev(self, t: Tensor, dtype=torch.float32) -> Tensor:
        t = t.detach().to(device=self.device, dtype=dtype).contiguous()
        self.keep.append(t)
        return t

    def ptr(self, t: Optional[Tensor], dtype=torch.float32) -> Optional[int]:
        return None if t is None else self.dev(t, dtype).data_ptr()

    def get(self, name: str) -> Tensor:
        if name not in self.w:
            raise KeyError(f"missing weight '{name}'")
        return self.w[name].to(torch.float32)

    def linear(self, kernel: Tensor, bias: Optional[Tensor], tc_kernel: Optional[Tensor] = None,
               ln: Optional[tuple] = None) -> L.SjLinear:
        """kernel: [K,N] (or stacked [G,K,N]); tc copy is [N,K] (or [G,N,K]) bf16 unless given explicitly.
        ln = (gamma [.., K], beta [.., K]): fold the preceding LayerNorm into the tensor-core copy."""
        s = L.SjLinear()
        s.w = self.ptr(kernel)
        s.b = self.ptr(bias)
        if self.tc:
            if ln is not None:
                g, b = ln[0].to(torch.float32), ln[1].to(torch.float32)
                folded = (kernel * g.unsqueeze(-1)).transpose(-1, -2).to(torch.bfloat16)  # [.., N, K]
                s.w_tc = self.ptr(folded, torch.bfloat16)
                s.tc_colsum = self.ptr(folded.to(torch.float32).sum(-1))
                tb = (b.unsqueeze(-1).to(torch.float64) * kernel.to(torch.float64)).sum(-2).to(torch.float32)
                s.tc_bias = self.ptr(tb + bias if bias is not None else tb)
            else:
                t = tc_kernel if tc_kernel is not None else kernel.transpose(-1, -2)
                s.w_tc = self.ptr(t, torch.bfloat16)
        return s

    def norm(self, prefix: str) -> L.SjNorm:
        s = L.SjNorm()
        s.g = self.ptr(self.get(prefix + "gamma"))
        s.b = self.ptr(self.get(prefix + "beta"))
        return s

    # ---- layers -----------------------------------------------------------------------------
    def swin_block(self, p: str) -> L.SjSwinBlockW:
        s = L.SjSwinBlockW()
        s.norm1 = self.norm(p + "norm1.")
        s.qkv = self.linear(self.get(p + "attn.qkv.kernel"), self.get(p + "attn.qkv.bias"))
        s.rpb_table = self.ptr(self.get(p + "attn.relative_position_bias_table"))
        s.proj = self.linear(self.get(p + "attn.proj.kernel"), self.get(p + "attn.proj.bias"))
        s.norm2 = self.norm(p + "norm2.")
        s.fc1 = self.linear(self.get(p + "mlp.fc1.kernel"), self.get(p + "mlp.fc1.bias"),
                            ln=(self.get(p + "norm2.gamma"), self.get(p + "norm2.beta")))
        s.fc2 = self.linear(self.get(p + "mlp.fc2.kernel"), self.get(p + "mlp.fc2.bias"))
        if self.tc:
            s.qkv_ln = self.linear(self.get(p + "attn.qkv.kernel"), self.get(p + "attn.qkv.bias"),
                                   ln=(self.get(p + "norm1.gamma"), self.get(p + "norm1.beta")))
        return s

    def patch_merge(self, p: str) -> L.SjPatchMergeW:
        s = L.SjPatchMergeW()
        s.norm = self.norm(p + "norm.")
        s.reduction = self.linear(self.get(p + "reduction.kernel"), None,
                                  ln=(self.get(p + "norm.gamma"), self.get(p + "norm.beta")))
        return s

    def patch_embed(self, p: str) -> L.SjPatchEmbedW:
        s = L.SjPatchEmbedW()
        k = self.get(p + "proj.kernel")  # [4,4,Cin,E] -> [16*Cin, E]
        k2 = k.reshape(-1, k.shape[-1])
        kpad = (k2.shape[0] + 63) // 64 * 64  # tensor-core copy: [E, Kpad], K zero-padded to a multiple of 64
        tck = torch.cat([k2, k2.new_zeros(kpad - k2.shape[0], k2.shape[1])], 0).t().contiguous()
        s.proj = self.linear(k2, self.get(p + "proj.bias"), tc_kernel=tck)
        s.norm = self.norm(p + "norm.")
        return s

    def basic_layer(self, p: str, dim: int, heads: int, depth: int, down: bool) -> L.SjBasicLayerW:
        s = L.SjBasicLayerW()
        arr = (L.SjSwinBlockW * depth)(*[self.swin_block(f"{p}blocks.{i}.") for i in range(depth)])
        self.keep.append(arr)
        s.blocks_host = C.cast(arr, C.POINTER(L.SjSwinBlockW))
        s.depth, s.dim, s.heads, s.has_down = depth, dim, heads, int(down)
        if down:
            s.down = self.patch_merge(p + "downsample.")
        return s

    def encoder(self, p: str, cfg: dict) -> L.SjEncoderW:
        E, depths, heads = cfg["embed_dim"], cfg["depths"], cfg["num_heads"]
        nl = len(depths)
        if nl > 4:
            raise ValueError("at most 4 encoder stages")
        s = L.SjEncoderW()
        s.pe_vec = self.patch_embed(p + "patch_embed_vecicle.")
        s.pe_map = self.patch_embed(p + "patch_embed_map.")
        s.pe_flow = self.patch_embed(p + "patch_embed_flow.")
        s.flow_norm = self.norm(p + "flow_norm.")
        s.all_patch_norm = self.norm(p + "all_patch_norm.")
        s.flow_layer = self.basic_layer(p + "flow_layer.", E, heads[0], depths[0], nl > 1)
        for i in range(nl):
            s.layers[i] = self.basic_layer(f"{p}basic_layers.{i}.", E * 2 ** i, heads[i], depths[i], i < nl - 1)
        s.num_layers, s.window_size, s.embed_dim = nl, cfg["window_size"], E
        return s

    def fgmsa(self, p: str, fg: bool) -> L.SjFgmsaW:
        s = L.SjFgmsaW()
        kq, kk, kv = (self.get(f"{p}proj_{n}.kernel")[0, 0] for n in ("q", "k", "v"))
        bq, bk, bv = (self.get(f"{p}proj_{n}.bias") for n in ("q", "k", "v"))
        s.qkv = self.linear(torch.cat([kq, kk, kv], 1), torch.cat([bq, bk, bv]))
        s.conv0_w = self.ptr(self.get(p + "conv_offset_0.kernel"))
        s.conv0_b = self.ptr(self.get(p + "conv_offset_0.bias"))
        s.conv_norm = self.norm(p + "conv_norm.")
        s.offproj_w = self.ptr(self.get(p + "conv_offset_proj.kernel")[0, 0])
        if fg:
            s.offproj2_w = self.ptr(self.get(p + "conv_offset_proj2.kernel")[0, 0])
            s.offproj2_b = self.ptr(self.get(p + "conv_offset_proj2.bias"))
        s.rpe_table = self.ptr(self.get(p + "rpe_table"))
        s.out = self.linear(self.get(p + "proj_out.kernel")[0, 0], self.get(p + "proj_out.bias"))
        if self.tc:
            s.conv0_w_tc = self.ptr(fg_conv_fragments(self.get(p + "conv_offset_0.kernel")), torch.bfloat16)
        return s

    def traj(self, p: str) -> L.SjTrajW:
        s = L.SjTrajW()
        e = p + "traj_net.traj_encoder."
        s.node_w = self.ptr(self.get(e + "node_feature.kernel")[0])
        s.node_b = self.ptr(self.get(e + "node_feature.bias"))
        s.node_qkv = self.linear(torch.cat([tfa_in_kernel(self.get(f"{e}node_attention.{n}_kernel"))
                                            for n in ("query", "key", "value")], 1), None)
        s.node_proj = self.linear(tfa_out_kernel(self.get(e + "node_attention.projection_kernel")),
                                  self.get(e + "node_attention.projection_bias"))
        s.vec_w = self.ptr(self.get(e + "vector_feature.kernel"))
        s.sublayer = self.linear(self.get(e + "sublayer.kernel"), self.get(e + "sublayer.bias"))
        a = p + "traj_net.cross_attention."
        s.ia_q = self.linear(tfa_in_kernel(self.get(a + "mha.query_kernel")), None)
        s.ia_kv = self.linear(torch.cat([tfa_in_kernel(self.get(a + "mha.key_kernel")),
                                         tfa_in_kernel(self.get(a + "mha.value_kernel"))], 1), None)
        s.ia_proj = self.linear(tfa_out_kernel(self.get(a + "mha.projection_kernel")), self.get(a + "mha.projection_bias"))
        s.ia_norm1 = self.norm(a + "norm1.")
        s.ia_ffn1 = self.linear(self.get(a + "FFN1.kernel"), self.get(a + "FFN1.bias"),
                                ln=(self.get(a + "norm1.gamma"), self.get(a + "norm1.beta")))
        s.ia_ffn2 = self.linear(self.get(a + "FFN2.kernel"), self.get(a + "FFN2.bias"))
        s.ia_norm2 = self.norm(a + "norm2.")
        s.obs_norm = self.norm(p + "traj_net.obs_norm.")
        s.occ_norm = self.norm(p + "traj_net.occ_norm.")
        s.seg_w = self.ptr(self.get(p + "traj_net.seg_embed.kernel"))

        def stack(fn):
            return torch.stack([fn(f"{p}cross_attn_obs.{t}.") for t in range(8)], 0)

        s.ca_q = self.linear(stack(lambda q: tfa_in_kernel(self.get(q + "mha.query_kernel"), 128)), None)
        s.ca_kv = self.linear(stack(lambda q: torch.cat([tfa_in_kernel(self.get(q + "mha.key_kernel"), 128),
                                                         tfa_in_kernel(self.get(q + "mha.value_kernel"), 128)], 1)), None)
        s.ca_proj = self.linear(stack(lambda q: tfa_out_kernel(self.get(q + "mha.projection_kernel"), 128)),
                                stack(lambda q: self.get(q + "mha.projection_bias")))
        s.ca_norm1.g = self.ptr(stack(lambda q: self.get(q + "norm1.gamma")))
        s.ca_norm1.b = self.ptr(stack(lambda q: self.get(q + "norm1.beta")))
        s.ca_ffn1 = self.linear(stack(lambda q: self.get(q + "FFN1.kernel")), stack(lambda q: self.get(q + "FFN1.bias")),
                                ln=(stack(lambda q: self.get(q + "norm1.gamma")), stack(lambda q: self.get(q + "norm1.beta"))))
        s.ca_ffn2 = self.linear(stack(lambda q: self.get(q + "FFN2.kernel")), stack(lambda q: self.get(q + "FFN2.bias")))
        s.ca_norm2.g = self.ptr(stack(lambda q: self.get(q + "norm2.gamma")))
        s.ca_norm2.b = self.ptr(stack(lambda q: self.get(q + "norm2.beta")))
        return s

    def _upconv(self, p: str) -> L.SjLinear:
        k = self.get(p + "kernel")  # [3,3,Ci,Co]
        ci, co = k.shape[2], k.shape[3]
        tc = None
        if self.tc:  # [py,px,a,b,Ci,Co] -> [4 phases][Co][4*Ci]
            f = fold_upconv_subpixel(k).reshape(4, 4 * ci, co)
            tc = f.transpose(1, 2)
        return self.linear(k.reshape(9 * ci, co), self.get(p + "bias"), tc)

    def _res(self, p: str) -> L.SjLinear:
        return self.linear(collapse_conv3d_811(self.get(p + "kernel")), self.get(p + "bias"))

    def decoder(self, p: str) -> L.SjDecoderW:
        s = L.SjDecoderW()
        for i in range(4):
            s.upconv[i] = self._upconv(f"{p}upconv_0s.{i}.")
        for i in range(2):
            s.res[i] = self._res(f"{p}res_layer.{i}.")
            s.upconv_f[i] = self._upconv(f"{p}upconv_f.{i}.")
        s.res_f = self._res(p + "res_f.")
        ko, kf = self.get(p + "output_layer.kernel"), self.get(p + "output_layer_f.kernel")
        s.out_w = self.ptr(torch.stack([ko.reshape(432, 2), kf.reshape(432, 2)], 0))
        s.out_b = self.ptr(torch.stack([self.get(p + "output_layer.bias"), self.get(p + "output_layer_f.bias")], 0))
        if self.tc:  # pointwise projection [2 heads][32 rows = tap*2+o (18 real)][64 ch (48 real)], zero padded
            wt = torch.zeros(2, 32, 64)
            for h, k in enumerate((ko, kf)):
                wt[h, :18, :48] = k.reshape(9, 48, 2).permute(0, 2, 1).reshape(18, 48)
            s.out_w_tc = self.ptr(wt.reshape(64, 64), torch.bfloat16)
        return s

    def model(self, cfg: dict, fg_msa: bool, fg: bool, large_ogm: bool) -> L.SjModelW:
        s = L.SjModelW()
        s.encoder = self.encoder("encoder.", cfg)
        if fg_msa:
            s.fgmsa = self.fgmsa("fg_msa_layer.", fg)
        s.traj = self.traj("trajnet_attn.")
        s.decoder = self.decoder("decoder.")
        s.fg_msa, s.fg, s.large_ogm = int(fg_msa), int(fg), int(large_ogm)
        return s
