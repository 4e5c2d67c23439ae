// Internal launcher interface between the C-ABI sequencing code (api.cu) and the kernels.
#pragma once
#include "common.cuh"

namespace sj {

// ---- generic GEMM / implicit-conv (gemm_simt.cu) ----------------------------------------------
enum AMode { A_PLAIN = 0, A_CONV3 = 1, A_MERGE = 2 };

// row = (m / inner) * outer + (m % inner) + group * gstride, then optionally through a gather map
// applied within blocks of map_len rows.  inner == 0 means identity.
struct RowMap {
  int inner = 0, outer = 0, gstride = 0;
  const int* map = nullptr;
  int map_len = 0;
};

struct GemmP {
  int amode = A_PLAIN;
  const void* A = nullptr;
  int lda = 0;
  const float* W = nullptr;  // [K,N] fp32
  int ldw = 0;
  long long w_gstride = 0;
  const float* bias = nullptr;
  int bias_gstride = 0;
  void* C = nullptr;
  int ldc = 0;
  const void* R = nullptr;  // residual added after the activation, indexed like C
  int ldr = 0;
  int M = 0, N = 0, K = 0, groups = 1;
  RowMap am, cm;
  // LayerNorm applied to A on load: (a - mean[row]) * rstd[row] * g[k] + b[k]
  const float* ln_mean = nullptr;
  const float* ln_rstd = nullptr;
  const float* ln_g = nullptr;
  const float* ln_b = nullptr;
  int ln_gstride = 0;
  int act = ACT_NONE;
  // optional (tensor-core path, N <= 256 in one tile): LayerNorm statistics of the output rows, written at the C row
  float* st_mean = nullptr;
  float* st_rstd = nullptr;
  float st_eps = 0.f;
  // A_CONV3: output H x Wd, input (H>>up) x (Wd>>up) x Cin, M = images*H*Wd, K = 9*Cin
  // A_MERGE: input H x Wd x Cin, M = B*(H/2)*(Wd/2), K = 4*Cin
  int H = 0, Wd = 0, Cin = 0, up = 0;
  // optional tensor-core copy of W (SjLinear.w_tc / tc_colsum / tc_bias), used when dtype == SJ_BF16
  const void* W_tc = nullptr;
  const float* tc_colsum = nullptr;
  const float* tc_bias = nullptr;
  void set_weights(const SjLinear& w) {
    W = w.w; bias = w.b; W_tc = w.w_tc; tc_colsum = w.tc_colsum; tc_bias = w.tc_bias;
  }
};
// dispatcher: tcgen05 kernel when dtype == SJ_BF16 and the shape/weights allow it, else the SIMT kernel
void gemm(Ctx& c, const GemmP& p);
void gemm_simt(Ctx& c, const GemmP& p);

// ---- tcgen05 dense GEMM (tc_gemm.cu) -------------------------------------------------------------
struct TcGemmP {
  const void* A = nullptr;  // bf16
  int a_mode = 0;           // 0: rows [M,K] (lda); 1: rows ordered [outer][G][inner]; 2: PatchMerging gather of [B,H,W,C]
  int lda = 0, a_inner = 0;
  int mH = 0, mW = 0, mC = 0;
  const void* Bw = nullptr;  // bf16 [G][N][K]
  int M = 0, N = 0, K = 0, groups = 1;
  const float* bias = nullptr;
  int bias_gstride = 0;
  int act = ACT_NONE;
  const void* R = nullptr;
  int ldr = 0;
  void* C = nullptr;
  int ldc = 0;
  RowMap cm;
  const float* ln_mean = nullptr;
  const float* ln_rstd = nullptr;
  const float* ln_s = nullptr;  // column sums of the folded weights, [G][N]
  int ln_gstride = 0;
  float* st_mean = nullptr;  // LayerNorm statistics of the output rows (needs pick_bn(N) == N)
  float* st_rstd = nullptr;
  float st_eps = 0.f;
  int dbg_shift = 0, dbg_bo = 0;  // hardware-semantics probe (sj_debug_gemm_shift)
};
bool tc_gemm_supported(const TcGemmP& p);
// true when tc_gemm can emit output-row statistics for an N-wide output (one n-tile)
bool tc_gemm_stats_ok(int N);
void tc_gemm(Ctx& c, const TcGemmP& p);
int num_sms();

// ---- tcgen05 decoder up-convolution (tc_conv.cu) ---------------------------------------------------
bool tc_upconv_supported(int H, int W, int Cin, int Cout);
// x bf16 [NB,H,W,Cin] -> y bf16 [NB,2H,2W,Cout] = ELU(conv3x3_SAME(nearest_up2(x)) + bias); w_tc = sub-pixel folded
// kernels [4 phases][Cout][4*Cin] bf16
void tc_upconv(Ctx& c, const void* x, void* y, const void* w_tc, const float* bias, int NB, int H, int W, int Cin,
               int Cout);

// small-channel variant (tc_upconv4.cu): Cin = 96, Cout = 48, all four sub-pixel phases per CTA with stacked-N MMAs
bool tc_upconv4_supported(int H, int W, int Cin, int Cout);
void tc_upconv4(Ctx& c, const void* x, void* y, const void* w_tc, const float* bias, int NB, int H, int W);

// mid-channel variant (tc_upconv1p.cu): Cin = 128, Cout = 96, one sub-pixel phase per CTA with resident weights
bool tc_upconv1p_supported(int H, int W, int Cin, int Cout);
void tc_upconv1p(Ctx& c, const void* x, void* y, const void* w_tc, const float* bias, int NB, int H, int W);

// ---- LayerNorm family (norm.cu) ---------------------------------------------------------------
// per-row mean / rstd (biased variance) of x[rows, C] (row stride ld)
void ln_stats(Ctx& c, const void* x, int rows, int C, int ld, float eps, float* mean, float* rstd);
// stats of the PatchMerging gather (modules.py:282-287) of x [B,H,W,C]: rows = B*H/2*W/2, width 4C
void ln_stats_merge(Ctx& c, const void* x, int B, int H, int W, int C, float eps, float* mean, float* rstd);
// y = LN(x) * g + b (+ res); gamma/beta of group ((row / g_div) % g_mod)
void layernorm(Ctx& c, const void* x, void* y, int rows, int C, const float* g, const float* b, float eps,
               const void* res, int g_div, int g_mod);
// y[r] = LN(x[gather(r)]) with the gather map applied within blocks of map_len rows
void layernorm_gather(Ctx& c, const void* x, void* y, int rows, int C, const float* g, const float* b, float eps,
                      const int* map, int map_len);

// bf16 fast paths for C in {96, 192, 384} (norm_fast.cu); each returns false when the shape is not covered
bool ln_fast(Ctx& c, bool stats_only, const void* x, void* y, int rows, int C, int ld, const float* g, const float* b,
             float eps, const void* res, int g_div, int g_mod, const int* map, int map_len, float* mean, float* rstd);
bool ln_stats_merge_fast(Ctx& c, const void* x, int B, int H, int W, int C, float eps, float* mean, float* rstd);
bool pe_combine_fast(Ctx& c, const void* c0, const void* c1, int B, int P, int pad1, const SjNorm& n0, const SjNorm& n1,
                     const SjNorm& nf, void* y, float* st_mean, float* st_rstd);

// ---- attention cores (attention.cu) -----------------------------------------------------------
// Window attention core, modules.py:109-131.  qkv [nWinTotal*64, 3C] (q|k|v, head-major inside),
// out [nWinTotal*64, C].  mask_mode 0: none; 1: shifted-window mask computed from (H,W,ws,shift);
// 2: explicit mask tensor [nW,64,64].
void window_attn_core(Ctx& c, const void* qkv, void* out, const float* rpb_table, int n_windows_total, int C,
                      int heads, int mask_mode, int H, int W, int shift, const float* mask, int nW);
// Generic multi-head attention core with tfa semantics: logits = (q/sqrt(D)) . k,
// + (-1e10) where (qmask*kmask)==0, softmax, . v.  q rows [batch*Nq] with stride ldq (head h at column
// h*D), same for k, v (batch*Nk rows).  out [batch*Nq, ldo], columns >= heads*D zero-filled.
// kmask index = (batch / mask_div) * Nk + key;  qmask index = (batch / mask_div) * Nq + query.
struct MhaP {
  const void *q = nullptr, *k = nullptr, *v = nullptr;
  void* out = nullptr;
  int ldq = 0, ldk = 0, ldv = 0, ldo = 0;
  int batch = 0, heads = 0, D = 0, Nq = 0, Nk = 0;
  const int* qmask = nullptr;
  const int* kmask = nullptr;
  int mask_div = 1;
  // FG-MSA bias (FG_MSA.py:150-172): pos fp32 [batch, heads, Nk, 2], rpe_table [31,31,heads]; 16x16 grid
  const float* fg_pos = nullptr;
  const float* fg_table = nullptr;
};
void mha_core(Ctx& c, const MhaP& p);
// bf16 warp-MMA variants (attn_mma.cu); return false when the shape is not covered
bool attn_mma_mha(Ctx& c, const MhaP& p);
bool attn_mma_window(Ctx& c, const void* qkv, void* out, const float* rpb_table, int n_windows_total, int C, int heads,
                     int mask_mode, int H, int W, int shift, const float* mask, int nW);

// ---- everything else (misc.cu) ----------------------------------------------------------------
void relative_position_index(Ctx& c, int ws, int64_t* out);
void shift_attn_mask(Ctx& c, int H, int W, int ws, int shift, float* out);
void window_token_map(Ctx& c, int H, int W, int ws, int shift, int32_t* out);
// window_partition (scatter 0) / window_reverse (scatter 1) on data (modules.py:49-63)
void window_permute(Ctx& c, const void* x, void* y, int B, int H, int W, int C, int ws, int scatter);

// Fused patch embedding (modules.py:437-446, :576-587, :602): for each token
//   y = LN_final( sum_i LN_i(conv4x4s4_i(img_i) + b_i) ), second input optional.
// Input i: fp32 [B,S_i,S_i,Cin_i] with channel element stride es_i.  The token grid is P x P with
// P = S_0/4; input 1 may cover only the centre (pad1 > 0: tokens outside [pad1, P-pad1) get 0 from it).
struct PatchEmbedP {
  const void* img[2] = {nullptr, nullptr};
  int itype[2] = {0, 0};  // InType of each input
  int Cin[2] = {0, 0}, es[2] = {1, 1}, S[2] = {0, 0};
  const float* w[2] = {nullptr, nullptr};   // [16*Cin, E]
  const float* bias[2] = {nullptr, nullptr};
  const float* g[2] = {nullptr, nullptr};
  const float* b[2] = {nullptr, nullptr};
  int n_in = 1, pad1 = 0;
  const float* gf = nullptr;  // final LN (NULL: skip)
  const float* bf = nullptr;
  void* y = nullptr;
  int B = 0, E = 0;
};
void patch_embed(Ctx& c, const PatchEmbedP& p);
// bf16 tensor-core variant: im2col to a bf16 matrix (then tc_gemm), and the LN / sum / LN combine
void im2col4(Ctx& c, const void* img, int itype, int B, int S, int Cin, int es, int Kpad, void* A);
// st_mean/st_rstd (optional): eps-1e-5 LayerNorm statistics of the output rows (norm1 of the first Swin block)
void pe_combine(Ctx& c, const void* c0, const void* c1, int B, int P, int pad1, const SjNorm& n0, const SjNorm& n1,
                const SjNorm& nf, void* y, float* st_mean = nullptr, float* st_rstd = nullptr);

// FG-MSA offset network (FG_MSA.py:84-92, :114-117, :134): q rows [B*256, ldq] (first 384 columns)
// -> off fp32 [B,8,256,2] (8*tanh), pos = off + (j,i)
void fg_offset(Ctx& c, const void* q, int ldq, const SjFgmsaW* w, int B, float* off, float* pos);
// bf16 warp-MMA variant (fg_offset_mma.cu); returns false when not applicable
bool fg_offset_mma(Ctx& c, const void* q, int ldq, const SjFgmsaW* w, int B, float* off, float* pos);
// flow_hidden [B,8,256,384] = off . Wp2 + bp2 (FG_MSA.py:120-123)
void fg_flow_hidden(Ctx& c, const float* off, const SjFgmsaW* w, int B, void* out);
// query [B,8,256,384] = q2[b,l,:] (+ off . Wp2 + bp2 if fg)   (modules.py:827-831)
void build_query(Ctx& c, const void* q2, const float* off, const SjFgmsaW* w, int B, int fg, void* query);

// trajectory glue (trajNet.py:38-48, :127-155, :179-185)
void traj_node(Ctx& c, const float* obs, const float* occ, const SjTrajW* w, int B, void* node, int* stepmask,
               int* cmask, float* vec);
void traj_pool_concat(Ctx& c, const void* proj, const float* vec, int n_actors, void* cat);
void traj_prep(Ctx& c, const void* E, const int* cmask, const float* seg_w, int n_actors, void* A, void* Q);
void traj_final(Ctx& c, const void* E, const void* F2, const SjTrajW* w, int n_actors, void* key);

// decoder head: two 3x3 48->2 convs (modules.py:767-770) written straight into the final layout
// out_layout 0: fp32 [B,8,256,256,4]; 1: fp32 [B,256,256,32]; 2: quantised bytes [B,256,256,32] (quantize_waypoint)
void out_conv(Ctx& c, const void* x_occ, const void* x_flow, const float* w, const float* b, int B, int out_layout,
              void* out);
// crop the centre of x [B,P,P,C] -> y [B,P/2,P/2,C] (modules.py:614-622)
void center_crop(Ctx& c, const void* x, void* y, int B, int P, int C);

// ---- K1: fused window-MSA on tcgen05 (tc_wmsa.cu), bf16, C = 96 / 3 heads or C = 192 / 6 heads, window 8 ----
bool tc_wmsa_supported(int B, int H, int W, int C, int heads, int ws, int shift);
// out[token] = x[token] + proj(window_attention(norm1(x)))  for x, out bf16 [B, H*W, C]; mean/rstd = norm1 stats of x
// mean2/rstd2 (optional): LayerNorm statistics (eps 1e-5) of the output rows, for norm2
void tc_wmsa(Ctx& c, const void* x, void* out, const float* mean, const float* rstd, const SjSwinBlockW& w, int B,
             int H, int W, int C, int shift, float* mean2, float* rstd2);

// two independent problems of the same geometry in one launch (the encoder's flow and raster branches in lock step)
void tc_wmsa_pair(Ctx& c, const void* const x[2], void* const out[2], const float* const mean[2], const float* const rstd[2],
                  const SjSwinBlockW* const w[2], int B, int H, int W, int shift, float* const mean2[2],
                  float* const rstd2[2]);

// Fused MLP half of a C = 96 Swin block (tc_mlp.cu): out = x + fc2(GELU(fc1(LN(x)))), bf16 [M, 96]; mean/rstd = norm2
// statistics of x; st_mean/st_rstd (optional) receive the eps-1e-5 LayerNorm statistics of the output rows
bool tc_mlp96_supported(int C, int hidden, const SjSwinBlockW& w);
void tc_mlp96(Ctx& c, const void* x, void* out, const float* mean, const float* rstd, const SjSwinBlockW& w, int M,
              float* st_mean, float* st_rstd);
void tc_mlp96_pair(Ctx& c, const void* const x[2], void* const out[2], const float* const mean[2], const float* const rstd[2],
                   const SjSwinBlockW* const w[2], int M, float* const st_mean[2], float* const st_rstd[2]);

// tcgen05 decoder head (tc_outconv.cu): bf16 inputs [B*8,256,256,48], fp32 logits out
void tc_out_conv(Ctx& c, const void* x_occ, const void* x_flow, const void* w_tc, const float* bias, int B,
                 int out_layout, void* out);

// fused last up-convolution + head projection (tc_upconv4h.cu): x bf16 [NB,H,W,96] -> z fp16 [NB,2H,2W,18] with
// z[q, tap*2+o] = sum_c ELU(upconv(x))[q,c] * head_w[tap,c,o]; head_w_tc = one head of SjDecoderW.out_w_tc
bool tc_upconv4h_supported(int H, int W, int Cin, int Cout);
void tc_upconv4h(Ctx& c, const void* x, void* z, const void* w_tc, const float* bias, const void* head_w_tc, int NB, int H,
                 int W);
// 9-tap shifted sum of both heads' projected columns + bias + final layout (headsum.cu); out_layout as out_conv
void head_tapsum(Ctx& c, const void* z_occ, const void* z_flow, const float* bias, int B, int out_layout, void* out);

// both 64x64 skip connections of the decoder in one kernel (tc_resadd2.cu): dst_a = src + ELU(skip_a . Wa[t] + ba) (may
// alias src), dst_b = dst_a + ELU(skip_b . Wb[t] + bb); skips bf16 [B,HW,96], src / dst bf16 [B,8,HW,128]
bool tc_resadd2_supported(int HW, int Cin, int Cout);
void tc_resadd2(Ctx& c, const void* skip_a, const void* skip_b, const void* src, void* dst_a, void* dst_b, const void* wa_tc,
                const float* bias_a, const void* wb_tc, const float* bias_b, int B, int HW);

// fused patch embedding (tc_patch_embed.cu): img0 [B,S0,S0,Cin0(,es0)] (+ img1 [B,S1,S1,Cin1] covering tokens
// [pad1, pad1 + S1/4) of the grid) -> y bf16 [B,(S0/4)^2,96] = LN_f(LN_0(conv_0) [+ LN_1(conv_1)]) + norm1 statistics
bool tc_patch_embed_supported(int B, int P, int Cin0, int Cin1, int P1, int pad1);
void tc_patch_embed(Ctx& c, const void* img0, int itype0, int S0, int Cin0, int es0, const SjPatchEmbedW& pw0,
                    const void* img1, int itype1, int S1, int Cin1, const SjPatchEmbedW* pw1, int pad1, const SjNorm& nf,
                    int B, void* y, float* st_mean, float* st_rstd);

// ---- validation-side loss / metrics (eval.cu) ----------------------------------------------------
size_t eval_workspace_bytes();
void eval_forward(Ctx& c, const float* pred, const float* gt_obs, const float* gt_occ, const float* gt_flow,
                  const float* origin, int B, int H, int W, int flags, float ogm_weight, float occ_weight,
                  float flow_origin_weight, float replica, float* out);

}  // namespace sj
