"""ctypes binding of libstrajnet_b200.so (the C ABI declared in include/strajnet_b200.h).

The product path has no CPU fallback: importing the library fails loudly when the CUDA
extension has not been built (run ``python -c "import __graft_entry__ as g; g.build()"``
or ``python -m strajnet_b200.build``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libstrajnet_b200.so")

SJ_F32, SJ_BF16 = 0, 1
SJ_IN_F32, SJ_IN_U8, SJ_IN_I8_DIV256 = 0, 1, 2
SJ_OK, SJ_EINVAL, SJ_EUNSUPPORTED, SJ_ECUDA, SJ_EWORKSPACE = 0, -1, -2, -3, -4

c_fp = C.c_void_p  # device pointers travel as plain addresses


class SjLinear(C.Structure):
    _fields_ = [("w", c_fp), ("b", c_fp), ("w_tc", c_fp), ("tc_colsum", c_fp), ("tc_bias", c_fp)]


class SjNorm(C.Structure):
    _fields_ = [("g", c_fp), ("b", c_fp)]


class SjSwinBlockW(C.Structure):
    _fields_ = [("norm1", SjNorm), ("qkv", SjLinear), ("rpb_table", c_fp), ("proj", SjLinear),
                ("norm2", SjNorm), ("fc1", SjLinear), ("fc2", SjLinear), ("qkv_ln", SjLinear)]


class SjPatchMergeW(C.Structure):
    _fields_ = [("norm", SjNorm), ("reduction", SjLinear)]


class SjPatchEmbedW(C.Structure):
    _fields_ = [("proj", SjLinear), ("norm", SjNorm)]


class SjBasicLayerW(C.Structure):
    _fields_ = [("blocks_host", C.POINTER(SjSwinBlockW)), ("depth", C.c_int), ("dim", C.c_int),
                ("heads", C.c_int), ("has_down", C.c_int), ("down", SjPatchMergeW)]


class SjEncoderW(C.Structure):
    _fields_ = [("pe_vec", SjPatchEmbedW), ("pe_map", SjPatchEmbedW), ("pe_flow", SjPatchEmbedW),
                ("flow_norm", SjNorm), ("all_patch_norm", SjNorm), ("flow_layer", SjBasicLayerW),
                ("layers", SjBasicLayerW * 4), ("num_layers", C.c_int), ("window_size", C.c_int),
                ("embed_dim", C.c_int)]


class SjFgmsaW(C.Structure):
    _fields_ = [("qkv", SjLinear), ("conv0_w", c_fp), ("conv0_b", c_fp), ("conv_norm", SjNorm),
                ("offproj_w", c_fp), ("offproj2_w", c_fp), ("offproj2_b", c_fp), ("rpe_table", c_fp),
                ("out", SjLinear), ("conv0_w_tc", C.c_void_p)]


class SjTrajW(C.Structure):
    _fields_ = [("node_w", c_fp), ("node_b", c_fp), ("node_qkv", SjLinear), ("node_proj", SjLinear),
                ("vec_w", c_fp), ("sublayer", SjLinear), ("ia_q", SjLinear), ("ia_kv", SjLinear),
                ("ia_proj", SjLinear), ("ia_norm1", SjNorm), ("ia_ffn1", SjLinear), ("ia_ffn2", SjLinear),
                ("ia_norm2", SjNorm), ("obs_norm", SjNorm), ("occ_norm", SjNorm), ("seg_w", c_fp),
                ("ca_q", SjLinear), ("ca_kv", SjLinear), ("ca_proj", SjLinear), ("ca_norm1", SjNorm),
                ("ca_ffn1", SjLinear), ("ca_ffn2", SjLinear), ("ca_norm2", SjNorm)]


class SjDecoderW(C.Structure):
    _fields_ = [("upconv", SjLinear * 4), ("res", SjLinear * 2), ("res_f", SjLinear),
                ("upconv_f", SjLinear * 2), ("out_w", c_fp), ("out_b", c_fp), ("out_w_tc", c_fp)]


class SjIoSpec(C.Structure):
    _fields_ = [("ogm_type", C.c_int), ("map_type", C.c_int), ("out_mode", C.c_int), ("ogm_planes", C.c_int)]


class SjEvalParams(C.Structure):
    _fields_ = [("flags", C.c_int), ("ogm_weight", C.c_float), ("occ_weight", C.c_float),
                ("flow_origin_weight", C.c_float), ("replica", C.c_float)]


class SjModelW(C.Structure):
    _fields_ = [("encoder", SjEncoderW), ("fgmsa", SjFgmsaW), ("traj", SjTrajW), ("decoder", SjDecoderW),
                ("fg_msa", C.c_int), ("fg", C.c_int), ("large_ogm", C.c_int)]


# name -> (restype, argtypes); mirrors include/strajnet_b200.h one to one
_i, _sz, _p, _ll = C.c_int, C.c_size_t, C.c_void_p, C.c_longlong
SIGNATURES = {
    "sj_version": (_i, []),
    "sj_sizeof": (_sz, [_i]),
    "sj_strerror": (C.c_char_p, [_i]),
    "sj_last_cuda_error": (C.c_char_p, []),
    "sj_launch_count": (_ll, [_i]),
    "sj_tc_launch_count": (_ll, [_i]),
    "sj_set_pdl": (_i, [_i]),
    "sj_crc32c": (C.c_uint32, [_p, _sz, C.c_uint32]),
    "sj_probe_start": (_i, [C.c_char_p]),
    "sj_probe_stop": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "sj_relative_position_index": (_i, [_i, _p, _p]),
    "sj_shift_attn_mask": (_i, [_i, _i, _i, _i, _p, _p]),
    "sj_window_token_map": (_i, [_i, _i, _i, _i, _p, _p]),
    "sj_window_partition_fwd": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "sj_window_reverse_fwd": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "sj_dense_fwd": (_i, [_p, _p, C.POINTER(SjLinear), _i, _i, _i, _i, _i, _p]),
    "sj_mlp_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "sj_mlp_fwd": (_i, [_p, _p, C.POINTER(SjLinear), C.POINTER(SjLinear), _i, _i, _i, _i, _p, _sz, _p]),
    "sj_window_attention_workspace_bytes": (_sz, [_i, _i, _i]),
    "sj_window_attention_fwd": (_i, [_p, _p, C.POINTER(SjSwinBlockW), _i, _i, _i, _i, _p, _i, _i, _p, _sz, _p]),
    "sj_swin_block_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "sj_swin_block_fwd": (_i, [_p, _p, C.POINTER(SjSwinBlockW), _i, _i, _i, _i, _i, _i, _i, _i, _p, _sz, _p]),
    "sj_patch_merging_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "sj_patch_merging_fwd": (_i, [_p, _p, C.POINTER(SjPatchMergeW), _p, _i, _i, _i, _i, _i, _p, _sz, _p]),
    "sj_patch_embed_fwd": (_i, [_p, _p, C.POINTER(SjPatchEmbedW), _i, _i, _i, _i, _i, _i, _p]),
    "sj_basic_layer_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "sj_basic_layer_fwd": (_i, [_p, _p, _p, C.POINTER(SjBasicLayerW), _i, _i, _i, _i, _i, _p, _sz, _p]),
    "sj_encoder_workspace_bytes": (_sz, [_i, _i, _i]),
    "sj_encoder_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, C.POINTER(SjEncoderW), _i, _i, _i, _i, _p, _sz, _p]),
    "sj_fgmsa_workspace_bytes": (_sz, [_i, _i]),
    "sj_fgmsa_fwd": (_i, [_p, _p, _p, _p, C.POINTER(SjFgmsaW), _i, _i, _p, _sz, _p]),
    "sj_traj_cross_attention_workspace_bytes": (_sz, [_i, _i]),
    "sj_traj_cross_attention_fwd": (_i, [_p, _p, _p, _p, C.POINTER(SjTrajW), _i, _i, _p, _sz, _p]),
    "sj_decoder_workspace_bytes": (_sz, [_i, _i]),
    "sj_decoder_fwd": (_i, [_p, _p, _p, _p, _p, C.POINTER(SjDecoderW), _i, _i, _i, _p, _sz, _p]),
    "sj_upconv_fwd": (_i, [_p, _p, C.POINTER(SjLinear), _i, _i, _i, _i, _i, _p]),
    "sj_res_add_fwd": (_i, [_p, _p, _p, C.POINTER(SjLinear), _i, _i, _i, _i, _i, _p]),
    "sj_patch_embed_sum_fwd": (_i, [_p, _i, _i, _i, _i, C.POINTER(SjPatchEmbedW), _p, _i, _i, _i, C.POINTER(SjPatchEmbedW), _i,
                                    C.POINTER(SjNorm), _i, _p, _p, _p, _p]),
    "sj_res_add2_fwd": (_i, [_p, _p, _p, _p, _p, C.POINTER(SjLinear), C.POINTER(SjLinear), _i, _i, _i, _i, _i, _p]),
    "sj_out_head_fwd": (_i, [_p, _p, _p, C.POINTER(SjDecoderW), _i, _i, _i, _p]),
    "sj_decoder_tail_workspace_bytes": (_sz, [_i, _i]),
    "sj_decoder_tail_fwd": (_i, [_p, _p, _p, C.POINTER(SjDecoderW), _i, _i, _i, _p, _sz, _p]),
    "sj_strajnet_workspace_bytes": (_sz, [_i, _i, _i]),
    "sj_strajnet_fwd_io": (_i, [_p, _p, _p, _p, _p, _p, C.POINTER(SjModelW), C.POINTER(SjIoSpec), _i, _i, _i, _p, _sz, _p]),
    "sj_strajnet_fwd": (_i, [_p, _p, _p, _p, _p, _p, C.POINTER(SjModelW), _i, _i, _i, _p, _sz, _p]),
    "sj_ogm_flow_eval_workspace_bytes": (_sz, []),
    "sj_ogm_flow_eval_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, C.POINTER(SjEvalParams), _p, _p, _sz, _p]),
}

_lib = None


class SjError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the sm_100a CUDA extension has not been built and there is no "
                "CPU fallback.  Build it with `python -m strajnet_b200.build`.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        structs = [SjLinear, SjNorm, SjSwinBlockW, SjPatchMergeW, SjPatchEmbedW, SjBasicLayerW, SjEncoderW,
                   SjFgmsaW, SjTrajW, SjDecoderW, SjModelW]
        for i, st in enumerate(structs):
            if L.sj_sizeof(i) != C.sizeof(st):
                raise ImportError(f"struct layout mismatch for {st.__name__}: library {L.sj_sizeof(i)} vs ctypes {C.sizeof(st)}")
        _lib = L
    return _lib


def check(status: int, what: str = "") -> None:
    """Convert a negative SjStatus into the exception the reference would raise."""
    if status == SJ_OK:
        return
    L = lib()
    msg = L.sj_strerror(status).decode()
    if status == SJ_ECUDA:
        msg += ": " + L.sj_last_cuda_error().decode()
    if status in (SJ_EINVAL, SJ_EUNSUPPORTED):
        raise ValueError(f"{what}: {msg}")  # Keras raises ValueError on shape problems
    raise SjError(f"{what}: {msg}")
