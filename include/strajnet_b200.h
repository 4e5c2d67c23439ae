/*
 * strajnet_b200 -- C ABI of the B200-native occupancy-flow forward path.
 *
 * This is the drop-in boundary.  The reference (georgeliu233/STrajNet) has no
 * FFI of its own: its hot path sits behind the Keras Layer/Model call protocol
 * (`__init__ / build() / call()`), so every entry point below replaces one Keras `call()`
 * and cites it as  <file>:<line>  relative to the reference tree.  The Python mirror of
 * those classes (strajnet_b200/layers.py, swinT.py) binds these symbols through ctypes;
 * INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary;
 *   - every data pointer is a DEVICE pointer unless the name ends in `_host`;
 *     weight STRUCTS live in host memory and hold device pointers;
 *   - tensors are channels-last, row-major, contiguous, 16-byte aligned;
 *   - `dtype` selects the activation type of `x`/`y`/intermediates (SJ_F32 or SJ_BF16);
 *     model inputs (rasters, actors) are always fp32; weights are fp32 ([in,out] Keras layout)
 *     with an optional bf16 tensor-core copy (`w_tc`, [out,in] K-major) used when dtype==SJ_BF16;
 *   - stream-ordered and asynchronous: no host sync, no allocation, no global mutable state that results depend on;
 *     the caller owns all buffers including `workspace`; safe to capture in a CUDA graph.  The whole-model forward forks
 *     the trajectory actor branch onto ONE helper stream per host thread and device (created lazily together with two
 *     events, joined before the branch's results are used: the caller sees plain stream order);
 *   - return 0 (SJ_OK) or a negative SjStatus; never throws.
 *
 * Process-level knobs (none of them changes a result beyond the documented bf16 tolerance; all are opt-in):
 *   - sj_set_pdl / SJ_PDL_MASK: programmatic dependent launch, OFF by default;
 *   - sj_probe_start / sj_probe_stop: per-thread timing probe for bench.py;
 *   - environment switches read once at first use, kept so that an earlier kernel generation can be re-measured against
 *     its replacement (each selects between two implementations of the same op): SJ_DISABLE_FUSED_WMSA,
 *     SJ_WMSA_MAX_C=96 (fused window-MSA for the 96-channel stages only), SJ_DISABLE_FUSED_PE (patch embedding as im2col -> GEMM -> combine),
 *     SJ_DISABLE_FUSED_STATS, SJ_DISABLE_FUSED_MLP, SJ_DISABLE_UPCONV4, SJ_DISABLE_UPCONV1P, SJ_DISABLE_HEAD_FUSION, SJ_DISABLE_RESADD2, SJ_DISABLE_LOCKSTEP,
 *     SJ_DISABLE_ATTN_MMA, SJ_DISABLE_IM2COL_STAGED, SJ_DISABLE_NORM_FAST, SJ_DISABLE_FG_OFFSET_MMA (fall back to the
 *     previous kernel), SJ_NO_SIDE_STREAM (keep the trajectory actor branch on the caller's stream), SJ_TCG_EW=16 / SJ_TCG_RPF (tc_gemm epilogue
 *     variants), SJ_TCG_BN=<n> (forces the tc_gemm n-tile width; read per call: tools/gemm_bn_sweep.py), SJ_NO_PDL (forces the PDL mask to 0).
 */
#ifndef STRAJNET_B200_H_
#define STRAJNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* sj_stream_t; /* cudaStream_t */

typedef enum { SJ_F32 = 0, SJ_BF16 = 1 } SjDType;

typedef enum {
  SJ_OK = 0,
  SJ_EINVAL = -1,       /* bad shape / null pointer / misalignment              */
  SJ_EUNSUPPORTED = -2, /* configuration outside what the kernels implement     */
  SJ_ECUDA = -3,        /* a CUDA call failed (see sj_last_cuda_error)          */
  SJ_EWORKSPACE = -4    /* workspace too small (see the *_workspace_bytes call) */
} SjStatus;

/* Dense / conv weights.  w: fp32 [K,N] (Keras [in,out], conv taps flattened into K);
 * b: fp32 [N] or NULL; w_tc: bf16 [N,K] (K-major) tensor-core copy or NULL.
 * When the layer consumes a LayerNorm output, the tensor-core copy has the norm folded in:
 * w_tc = bf16(gamma[k]*w[k,n]), tc_colsum[n] = sum_k w_tc[n,k], tc_bias[n] = b[n] + sum_k beta[k]*w[k,n];
 * the kernel then applies  rstd*(x.w_tc - mean*tc_colsum) + tc_bias  in its epilogue (else both NULL). */
typedef struct { const float* w; const float* b; const void* w_tc; const float* tc_colsum; const float* tc_bias; } SjLinear;
typedef struct { const float* g; const float* b; } SjNorm; /* LayerNormalization gamma/beta */

/* SwinTransformerBlock (modules.py:163-262) incl. WindowAttention (:66-134) and Mlp (:31-46). */
typedef struct {
  SjNorm norm1;
  SjLinear qkv;           /* [C,3C] + [3C]            modules.py:76  */
  const float* rpb_table; /* [(2ws-1)^2, heads]       modules.py:83  */
  SjLinear proj;          /* [C,C] + [C]              modules.py:79  */
  SjNorm norm2;
  SjLinear fc1;           /* [C,4C] + [4C]            modules.py:36  */
  SjLinear fc2;           /* [4C,C] + [C]             modules.py:37  */
  SjLinear qkv_ln;        /* qkv again, tensor-core copy with norm1 folded in (fused window-MSA kernel); w/b unused */
} SjSwinBlockW;

typedef struct { SjNorm norm; SjLinear reduction; } SjPatchMergeW; /* modules.py:265-292: LN(4C), [4C,2C] no bias */
typedef struct { SjLinear proj; SjNorm norm; } SjPatchEmbedW;      /* modules.py:417-446: [16*Cin,E]+[E], LN(E) */

/* BasicLayer (modules.py:317-364) */
typedef struct {
  const SjSwinBlockW* blocks_host; /* host array of `depth` blocks */
  int depth, dim, heads, has_down;
  SjPatchMergeW down;
} SjBasicLayerW;

/* SwinTransformerEncoder as configured by STrajNet (modules.py:782-785; forward :570-624) */
typedef struct {
  SjPatchEmbedW pe_vec, pe_map, pe_flow;
  SjNorm flow_norm, all_patch_norm;
  SjBasicLayerW flow_layer;
  SjBasicLayerW layers[4];
  int num_layers, window_size, embed_dim;
} SjEncoderW;

/* FGMSA (FG_MSA.py:20-183), n_heads = n_groups = 8, 48 channels per head, 16x16 tokens */
typedef struct {
  SjLinear qkv;            /* proj_q|proj_k|proj_v concatenated: [384,1152] + [1152]  FG_MSA.py:58-62 */
  const float* conv0_w;    /* [3,3,48,384]  grouped 3x3, groups=8                      FG_MSA.py:51    */
  const float* conv0_b;    /* [384] */
  SjNorm conv_norm;        /* eps 1e-3                                                FG_MSA.py:52    */
  const float* offproj_w;  /* [48,2] no bias                                          FG_MSA.py:54    */
  const float* offproj2_w; /* [2,384] (fg only, else NULL)                            FG_MSA.py:56    */
  const float* offproj2_b; /* [384] */
  const float* rpe_table;  /* [31,31,8]                                               FG_MSA.py:70    */
  SjLinear out;            /* proj_out [384,384] + [384]                              FG_MSA.py:64    */
  const void* conv0_w_tc;  /* bf16 copy of conv0_w in mma.m16n8k16 B-fragment order, or NULL (bf16 path then converts conv0_w
                            * on the fly): [8 groups][9 taps][3 k-steps][3 n-pairs][32 lanes][8] with, for lane = 4*g + t and
                            * n-tile nt = 2*pair + (e>>2): element e&3 = conv0_w[tap][16*ks + 2*t + {0,1,8,9}[e&3]][48*group + 8*nt + g] */
} SjFgmsaW;

/* TrajNetCrossAttention (trajNet.py:236-319) = TrajNet (:91-187) + 8x Cross_AttentionT (:189-234).
 * tfa.layers.MultiHeadAttention kernels [H,in,hs] are re-laid-out as [in, H*hs] (heads concatenated
 * along columns); head_size 42 is kept at 42 with the 126 columns zero-padded to 128.              */
typedef struct {
  const float* node_w;  /* [5,64]   Conv1D(64,1) kernel    trajNet.py:32 */
  const float* node_b;  /* [64] */
  SjLinear node_qkv;    /* [64,768] = q(4x64)|k|v, no bias trajNet.py:33 */
  SjLinear node_proj;   /* [256,320] + [320] */
  const float* vec_w;   /* [3,64] no bias                  trajNet.py:35 */
  SjLinear sublayer;    /* [384,384] + [384], ELU          trajNet.py:36 */
  SjLinear ia_q;        /* interaction Cross_Attention (trajNet.py:65-87): [384,384] */
  SjLinear ia_kv;       /* [384,768] = k|v */
  SjLinear ia_proj;     /* [384,384] + [384] */
  SjNorm ia_norm1;
  SjLinear ia_ffn1;     /* [384,1536] + b, ELU */
  SjLinear ia_ffn2;     /* [1536,384] + b */
  SjNorm ia_norm2;
  SjNorm obs_norm, occ_norm;
  const float* seg_w;   /* [2,384] */
  /* the 8 per-waypoint Cross_AttentionT, stacked on a leading dim of 8 */
  SjLinear ca_q;        /* [8][384,128]  (3 heads x 42 = 126 cols, padded) */
  SjLinear ca_kv;       /* [8][384,256]  k at cols 0..125, v at cols 128..253 */
  SjLinear ca_proj;     /* [8][128,128] + [8][128] (rows 126,127 zero) */
  SjNorm ca_norm1;      /* [8][128] */
  SjLinear ca_ffn1;     /* [8][128,512] + [8][512], ELU */
  SjLinear ca_ffn2;     /* [8][512,384] + [8][384] */
  SjNorm ca_norm2;      /* [8][384] */
} SjTrajW;

/* Pyramid3DDecoder as configured by STrajNet (modules.py:800-801; forward :739-772).
 * 3x3 kernels are flattened to [9*Cin, Cout] (tap-major).  The (8,1,1) Conv3D over the
 * 8x-repeated skip tensor is pre-collapsed on the host to 8 per-waypoint 1x1 kernels
 * W_eff[t] = sum_{k: 0<=t+k-3<=7} W[k]  (TF SAME: 3 before / 4 after), stored [8][Cin,Cout].
 * For dtype==SJ_BF16, upconv*.w_tc holds the sub-pixel-folded kernels: nearest-x2 followed by a
 * 3x3 SAME conv equals four 2x2 convs on the low-res input; layout [4 phases][Cout][4*Cin] bf16. */
typedef struct {
  SjLinear upconv[4];   /* 384->192, 192->128, 128->96, 96->48, ELU */
  SjLinear res[2];      /* collapsed res_layer: [8][192,192], [8][96,128], ELU */
  SjLinear res_f;       /* collapsed res_f: [8][96,128], ELU */
  SjLinear upconv_f[2]; /* 128->96, 96->48, ELU */
  const float* out_w;   /* [2][9*48,2]: output_layer then output_layer_f */
  const float* out_b;   /* [2][2] */
  const void* out_w_tc; /* bf16 [2 heads][32 rows][64]: row tap*2+o (<18), channel c (<48) = out_w[head][tap*48+c][o], rest 0; or NULL */
} SjDecoderW;

typedef struct {
  SjEncoderW encoder;
  SjFgmsaW fgmsa;
  SjTrajW traj;
  SjDecoderW decoder;
  int fg_msa, fg, large_ogm; /* STrajNet ctor flags, modules.py:778-779 */
} SjModelW;

/* ---- library ---------------------------------------------------------------------------- */
int sj_version(void);
/* sizeof() of the weight structs, in declaration order (SjLinear=0 ... SjModelW=10): lets a foreign-language
 * binding verify its struct layout against the library it loaded */
size_t sj_sizeof(int which);
const char* sj_strerror(int status);
const char* sj_last_cuda_error(void); /* text of the last CUDA failure seen by this thread */
/* number of kernels launched by this thread through the library since the last reset */
long long sj_launch_count(int reset);
/* of which tcgen05 (tensor-core) kernels */
long long sj_tc_launch_count(int reset);
/* Programmatic dependent launch (CUDA launch attribute "programmatic stream serialization"): OFF by default (inside a
 * replayed CUDA graph it measured slower); when on, each kernel of a forward is launched so that its prologue overlaps
 * the tail of its predecessor on the stream; every kernel of the library orders itself behind its predecessor with
 * griddepcontrol.wait before touching activations, so results are identical either way.  `mask`: bit 0 = the
 * tcgen05 kernels, bit 1 = all other kernels (3 = every launch, 0 = off; SJ_PDL_MASK in the environment sets the
 * initial value).  Process-wide; returns the previous mask.  mask < 0 only queries. */
int sj_set_pdl(int mask);

/* Host-side helper of the record / checkpoint readers (scope rows f2, f3): CRC-32C (Castagnoli) of `n` bytes,
 * continuing from `crc` (0 to start).  This is the checksum of the TFRecord framing the reference reads through
 * tf.data.TFRecordDataset (inference.py:256-258, train.py:380-388) and of TF tensor-bundle checkpoints
 * (model.load_weights, inference.py:283).  Pure host code: no GPU needed. */
uint32_t sj_crc32c(const void* data, size_t n, uint32_t crc);

/* Opt-in timing probe for bench.py: CUDA events (created lazily by the library; the one exception to
 * "allocates nothing") are recorded on the launching stream around every kernel launched by this thread
 * whose forward step name starts with `role_prefix` ("enc", "fgmsa", "traj", "dec.upconv3", ...).
 * sj_probe_stop synchronises on those events and returns their summed duration and count. */
int sj_probe_start(const char* role_prefix);
int sj_probe_stop(double* total_ms, int* n_launches);

/* ---- integer maps (bit-exact rows of SURVEY §8 a3/a4/a5) --------------------------------- */
/* WindowAttention.build, modules.py:88-100: int64 [ws*ws, ws*ws] */
int sj_relative_position_index(int ws, int64_t* out, sj_stream_t stream);
/* SwinTransformerBlock.build, modules.py:189-214: float [nW, ws*ws, ws*ws] in {0,-100} */
int sj_shift_attn_mask(int H, int W, int ws, int shift, float* out, sj_stream_t stream);
/* roll(-shift) + window_partition as a gather map: out[w*ws*ws+n] = source token index in [0,H*W);
 * shift=0 gives window_partition alone (modules.py:49-55, :230-239) */
int sj_window_token_map(int H, int W, int ws, int shift, int32_t* out, sj_stream_t stream);
/* window_partition / window_reverse on data (modules.py:49-63) */
int sj_window_partition_fwd(const void* x, void* windows, int B, int H, int W, int C, int ws, int dtype, sj_stream_t stream);
int sj_window_reverse_fwd(const void* windows, void* x, int B, int H, int W, int C, int ws, int dtype, sj_stream_t stream);

/* ---- layers ------------------------------------------------------------------------------ */
/* keras.layers.Dense(units=N, activation=act) as used throughout modules.py / trajNet.py: y = act(x[M,K] . w + b);
 * act 0 none, 1 tanh-GELU (modules.py:18), 2 ELU */
int sj_dense_fwd(const void* x, void* y, const SjLinear* w, int M, int N, int K, int act, int dtype, sj_stream_t stream);

/* Mlp.call, modules.py:40-46: y = fc2(Gelu(fc1(x))), x [M,C] */
size_t sj_mlp_workspace_bytes(int M, int C, int hidden, int dtype);
int sj_mlp_fwd(const void* x, void* y, const SjLinear* fc1, const SjLinear* fc2, int M, int C, int hidden,
               int dtype, void* workspace, size_t workspace_bytes, sj_stream_t stream);

/* WindowAttention.call, modules.py:103-134: xw [B_,N=ws*ws,C]; mask [nW,N,N] fp32 or NULL */
size_t sj_window_attention_workspace_bytes(int B_, int C, int dtype);
int sj_window_attention_fwd(const void* xw, void* y, const SjSwinBlockW* w, int B_, int C, int heads, int ws,
                            const float* mask, int nW, int dtype, void* workspace, size_t workspace_bytes,
                            sj_stream_t stream);

/* SwinTransformerBlock.call, modules.py:220-262: x,y [B,H*W,C] */
size_t sj_swin_block_workspace_bytes(int B, int H, int W, int C, int dtype);
int sj_swin_block_fwd(const void* x, void* y, const SjSwinBlockW* w, int B, int H, int W, int C, int heads,
                      int ws, int shift, int dtype, void* workspace, size_t workspace_bytes, sj_stream_t stream);

/* PatchMerging.call, modules.py:274-292: x [B,H*W,C] -> y [B,H*W/4,2C]; add (same shape as y) or NULL
 * is added to the result (the `x + flow_x` of modules.py:613) */
size_t sj_patch_merging_workspace_bytes(int B, int H, int W, int C, int dtype);
int sj_patch_merging_fwd(const void* x, void* y, const SjPatchMergeW* w, const void* add, int B, int H, int W,
                         int C, int dtype, void* workspace, size_t workspace_bytes, sj_stream_t stream);

/* PatchEmbed.call, modules.py:437-446: img fp32 [B,S,S,Cin] with element stride `elem_stride`
 * between channels (2 selects plane 0 of ogm[...,11,2]) -> y [B,(S/4)^2,E] */
int sj_patch_embed_fwd(const float* img, void* y, const SjPatchEmbedW* w, int B, int S, int Cin, int elem_stride,
                       int E, int dtype, sj_stream_t stream);

/* The encoder's embedding stage as one tcgen05 kernel (bf16 only; modules.py:572-587 + :602, and :576-577 for the flow
 * branch):  y = LN_final( PatchEmbed_0(img0) [+ PatchEmbed_1(img1)] ), bf16 [B,(S0/4)^2,96].
 * img0 [B,S0,S0,Cin0] of input type type0 (SjInputType below: 0 fp32, 1 bool bytes, 2 int8/256) with element stride es0
 * (2 selects plane 0 of ogm[...,11,2]); img1 [B,S1,S1,Cin1] or NULL: its S1/4 x S1/4 tokens are added at token offset
 * (pad1, pad1) of the grid (32 in 512-input mode, modules.py:582-585), tokens outside get no term.
 * st_mean / st_rstd: fp32 [B*(S0/4)^2] LayerNorm (eps 1e-5) statistics of y, or NULL.  Needs w*->proj.w_tc and 16*Cin0 <= 192,
 * 16*Cin1 <= 64, S0/4 in {64,128}; otherwise SJ_EUNSUPPORTED. */
int sj_patch_embed_sum_fwd(const void* img0, int type0, int S0, int Cin0, int es0, const SjPatchEmbedW* w0, const void* img1,
                           int type1, int S1, int Cin1, const SjPatchEmbedW* w1, int pad1, const SjNorm* final_norm, int B,
                           void* y, float* st_mean, float* st_rstd, sj_stream_t stream);

/* BasicLayer.call, modules.py:351-364: returns (x_down or x, res).  y_down may be NULL when !has_down. */
size_t sj_basic_layer_workspace_bytes(int B, int H, int W, int C, int dtype);
int sj_basic_layer_fwd(const void* x, void* y_down, void* res, const SjBasicLayerW* w, int B, int H, int W,
                       int ws, int dtype, void* workspace, size_t workspace_bytes, sj_stream_t stream);

/* SwinTransformerEncoder.call, modules.py:626 / :570-624.  ogm fp32 [B,S,S,11,2], map fp32 [B,256,256,3],
 * flow fp32 [B,S,S,2]; outputs flow_res [B,4096,96], res0 [B,4096,96], res1 [B,1024,192], res2 [B,256,384] */
size_t sj_encoder_workspace_bytes(int B, int S, int dtype);
int sj_encoder_fwd(const float* ogm, const float* map_img, const float* flow, void* flow_res, void* res0,
                   void* res1, void* res2, const SjEncoderW* w, int B, int S, int large_input, int dtype,
                   void* workspace, size_t workspace_bytes, sj_stream_t stream);

/* FGMSA.call, FG_MSA.py:106-183: x [B,16,16,384] -> y [B,16,16,384], pos fp32 [B,8,16,16,2],
 * flow_hidden [B,8,16,16,384] (NULL to skip; requires offproj2) */
size_t sj_fgmsa_workspace_bytes(int B, int dtype);
int sj_fgmsa_fwd(const void* x, void* y, float* pos, void* flow_hidden, const SjFgmsaW* w, int B, int dtype,
                 void* workspace, size_t workspace_bytes, sj_stream_t stream);

/* TrajNetCrossAttention.call, trajNet.py:284-319: pic [B,8,256,384], obs fp32 [B,48,11,8],
 * occ fp32 [B,16,11,8] -> out [B,8,256,384] */
size_t sj_traj_cross_attention_workspace_bytes(int B, int dtype);
int sj_traj_cross_attention_fwd(const void* pic, const float* obs, const float* occ, void* out, const SjTrajW* w,
                                int B, int dtype, void* workspace, size_t workspace_bytes, sj_stream_t stream);

/* Pyramid3DDecoder.call, modules.py:739-772: x [B,8,16,16,384] + skips -> fp32 out.
 * out_layout 0: [B,8,256,256,4] (the Keras layer's own result); 1: [B,256,256,32] (STrajNet, modules.py:838) */
size_t sj_decoder_workspace_bytes(int B, int dtype);
int sj_decoder_fwd(const void* x, const void* flow_res, const void* res0, const void* res1, float* out,
                   const SjDecoderW* w, int B, int out_layout, int dtype, void* workspace, size_t workspace_bytes,
                   sj_stream_t stream);

/* Single decoder stages, exported so that each tensor-core kernel has its own parity test.
 * One up-sampling stage, modules.py:746-749 (`UpSampling3D((1,2,2))` then `upconv_0s[i]` = Conv2D 3x3 SAME + ELU, shared
 * over the leading dims): x [NB,H,H,Cin] -> y [NB,2H,2H,Cout]. */
int sj_upconv_fwd(const void* x, void* y, const SjLinear* w, int NB, int H, int Cin, int Cout, int dtype,
                  sj_stream_t stream);
/* One skip connection, modules.py:750-757 (`res_layer[i]` = Conv3D (8,1,1) + ELU over the 8x-repeated skip tensor, then
 * add), with the collapsed per-waypoint kernels of SjDecoderW.res: dst[b,t] = src[b,t] + ELU(skip[b] . W_eff[t] + bias);
 * skip [B,HW,Cin], src/dst [B,8,HW,Cout] (dst may alias src). */
int sj_res_add_fwd(const void* skip, const void* src, void* dst, const SjLinear* w, int B, int HW, int Cin, int Cout,
                   int dtype, sj_stream_t stream);
/* Both skip connections that meet at 64x64, modules.py:750-757 (res_layer[1] on res0) and :762-765 (res_f on flow_res, added
 * to x AFTER the res0 add): dst_a = src + ELU(skip_a . Wa_eff[t] + ba), dst_b = dst_a + ELU(skip_b . Wb_eff[t] + bb);
 * skips [B,HW,Cin], src / dst [B,8,HW,Cout]; dst_a may alias src.  One fused kernel for dtype == SJ_BF16, Cin = 96,
 * Cout = 128; two sj_res_add_fwd steps otherwise. */
int sj_res_add2_fwd(const void* skip_a, const void* skip_b, const void* src, void* dst_a, void* dst_b, const SjLinear* wa,
                    const SjLinear* wb, int B, int HW, int Cin, int Cout, int dtype, sj_stream_t stream);
/* The two heads, modules.py:767-770 (`output_layer` on x, `output_layer_f` on the flow branch, Conv2D 3x3 SAME 48->2,
 * concatenated) + the transpose of :838: x_occ, x_flow [B*8,256,256,48] -> out; out_layout 0/1 as sj_decoder_fwd,
 * 2 = quantised submission bytes uint8 [B,256,256,32] (inference.py:124-136,160-182). */
int sj_out_head_fwd(const void* x_occ, const void* x_flow, void* out, const SjDecoderW* w, int B, int out_layout,
                    int dtype, sj_stream_t stream);

/* The decoder's tail: last up-sampling stage of both branches + both heads + final layout (modules.py:746-749 with
 * i = 3, :732-737 second iteration, :767-770, :838): x3, f3 [B*8,128,128,96] -> out (out_layout as sj_out_head_fwd).
 * With dtype == SJ_BF16 this is the fused path: the 48-channel full-resolution tensors stay on chip. */
size_t sj_decoder_tail_workspace_bytes(int B, int dtype);
int sj_decoder_tail_fwd(const void* x3, const void* f3, void* out, const SjDecoderW* w, int B, int out_layout, int dtype,
                        void* workspace, size_t workspace_bytes, sj_stream_t stream);

/* Raw I/O of the serving loop.  Inputs as the reference's record decode holds them before the
 * float casts (inference.py:91-93): ogm bool bytes [B,S,S,11,2] (SJ_IN_U8: nonzero -> 1.0), map int8 [B,256,256,3]
 * (SJ_IN_I8_DIV256: value/256).  out_mode 1 fuses the submission quantisation (inference.py:124-136,160-182) into the
 * decoder head: out is uint8 [B,256,256,32], channel k*4+{0,1} = round(sigmoid(logit)*255) as uint8, k*4+{2,3} =
 * clip(round(flow),-128,127) as int8. */
typedef enum { SJ_IN_F32 = 0, SJ_IN_U8 = 1, SJ_IN_I8_DIV256 = 2 } SjInputType;
typedef struct {
  int ogm_type;   /* SJ_IN_F32 or SJ_IN_U8 */
  int map_type;   /* SJ_IN_F32 or SJ_IN_I8_DIV256 */
  int out_mode;   /* 0: fp32 logits, 1: quantised submission bytes */
  int ogm_planes; /* 2: the record's [B,S,S,11,2] raster (the model reads plane 0 only, modules.py:572);
                     1: the caller hands over that plane alone, [B,S,S,11] (half the host->device bytes of the largest input) */
} SjIoSpec;
int sj_strajnet_fwd_io(const void* ogm, const void* map_img, const float* flow, const float* obs, const float* occ,
                       void* out, const SjModelW* w, const SjIoSpec* io, int B, int S, int dtype, void* workspace,
                       size_t workspace_bytes, sj_stream_t stream);

/* STrajNet.call, modules.py:815-839: -> out fp32 [B,256,256,32] */
size_t sj_strajnet_workspace_bytes(int B, int S, int dtype);
int sj_strajnet_fwd(const float* ogm, const float* map_img, const float* flow, const float* obs, const float* occ,
                    float* out, const SjModelW* w, int B, int S, int dtype, void* workspace, size_t workspace_bytes,
                    sj_stream_t stream);

/* ---- validation-side ops, forward values only ------------------------------------------
 * One pass over the grids computes both OGMFlow_loss.__call__ (loss.py:50-170; helpers :172-292) and
 * compute_occupancy_flow_metrics (occu_metric.py:26-140; helpers :152-343, the zero-border `sample` :345-409), as
 * val_step uses them (train.py:252-283).
 *   pred    fp32 [B,H,W,32]   channel 4k+0 / 4k+1 observed / occluded occupancy, 4k+2, 4k+3 flow (train.py:103-121);
 *                             logits, or probabilities when SJ_EVAL_PRED_IS_PROB (metrics only: the output of
 *                             _apply_sigmoid_to_occupancy_logits, train.py:142-154)
 *   gt_obs, gt_occ, origin   fp32 [B,8,H,W]; gt_flow fp32 [B,8,H,W,2]   (train.py:126-140)
 *   out     fp32 [SJ_EVAL_OUT_FLOATS] (device): [0..3] observed_xe, occluded_xe, flow, flow_warp_xe;
 *           [4..10] vehicles_observed_auc, vehicles_occluded_auc, vehicles_observed_iou, vehicles_occluded_iou,
 *           vehicles_flow_epe, vehicles_flow_warped_occupancy_auc, vehicles_flow_warped_occupancy_iou;
 *           [11..18] the per-waypoint gate `res` of the use_gt branch (loss.py:125-137).
 * flags: the OGMFlow_loss constructor switches (loss.py:24-25), which of the two results to compute, and
 * compute_occupancy_flow_metrics' no_warp. */
enum {
  SJ_EVAL_USE_FOCAL = 1, SJ_EVAL_NO_USE_WARP = 2, SJ_EVAL_USE_PRED = 4, SJ_EVAL_USE_GT = 8,
  SJ_EVAL_PRED_IS_PROB = 16, SJ_EVAL_LOSS = 32, SJ_EVAL_METRICS = 64, SJ_EVAL_METRICS_NO_WARP = 128,
  /* tf.keras.metrics.AUC label semantics.  Default (flag clear): y_true is cast to bool -- tf.keras <= 2.5 and >= 2.8.
   * Flag set: tf.keras 2.6 / 2.7, whose evenly-spaced-threshold path (_update_confusion_matrix_variables_optimized) keeps
   * the label as a float: a sample adds y_true to the true-positive mass and 1 - y_true to the false-positive mass.
   * Only vehicles_flow_warped_occupancy_auc can differ (its label is the fractional flow-grounded prediction,
   * occu_metric.py:121-123); the reference pins no TensorFlow version, so the caller chooses. */
  SJ_EVAL_AUC_FLOAT_LABELS = 256
};
#define SJ_EVAL_OUT_FLOATS 19
typedef struct { int flags; float ogm_weight; float occ_weight; float flow_origin_weight; float replica; } SjEvalParams;
size_t sj_ogm_flow_eval_workspace_bytes(void);
int sj_ogm_flow_eval_fwd(const float* pred, const float* gt_obs, const float* gt_occ, const float* gt_flow,
                         const float* origin, int B, int H, int W, const SjEvalParams* params, float* out,
                         void* workspace, size_t workspace_bytes, sj_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* STRAJNET_B200_H_ */
