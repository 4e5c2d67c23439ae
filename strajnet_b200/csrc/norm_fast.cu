// LayerNorm family, bf16 fast path for the model's channel counts (96 / 192 / 384 = 24 x {4, 8, 16}): every lane owns 24
// consecutive channels of a row (three 16-byte loads, kept in registers for both statistics passes and the affine), so
// a warp covers 8 / 4 / 2 rows and every byte is read exactly once.  The generic kernels of norm.cu (one warp per row,
// 8-byte loads, three passes over the row) remain for fp32 and for other widths.
#include "kernels.h"
#include "tc_common.cuh"

namespace sj {
namespace {

using tc::ld8_bf16;
using tc::st8_bf16;

template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void load24(const bf16* p, float (&v)[24]) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float t[8];
    ld8_bf16(p + 8 * j, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[8 * j + i] = t[i];
  }
}

// mean / rstd of the 24*LPR values spread over the LPR lanes of a row group (two passes over registers)
template <int LPR>
__device__ __forceinline__ void stats24(const float (&v)[24], float eps, float& mu, float& rs) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) s += v[i];
  mu = group_sum<LPR>(s) * (1.0f / (24 * LPR));
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    const float d = v[i] - mu;
    q = fmaf(d, d, q);
  }
  rs = rsqrtf(group_sum<LPR>(q) * (1.0f / (24 * LPR)) + eps);
}

__device__ __forceinline__ void affine24(float (&v)[24], float mu, float rs, const float* g, const float* b) {
#pragma unroll
  for (int i = 0; i < 24; i += 4) {
    const float4 g4 = *reinterpret_cast<const float4*>(g + i), b4 = *reinterpret_cast<const float4*>(b + i);
    v[i] = (v[i] - mu) * rs * g4.x + b4.x;
    v[i + 1] = (v[i + 1] - mu) * rs * g4.y + b4.y;
    v[i + 2] = (v[i + 2] - mu) * rs * g4.z + b4.z;
    v[i + 3] = (v[i + 3] - mu) * rs * g4.w + b4.w;
  }
}

// STATS: mean/rstd only.  Otherwise y[row] = LN(x[gather(row)]) * g[grp] + b[grp] (+ res[row])
template <int LPR, bool STATS>
__global__ void __launch_bounds__(256) ln_rows_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int rows, int ld,
                                                      const float* __restrict__ g, const float* __restrict__ b, float eps,
                                                      const bf16* __restrict__ res, int g_div, int g_mod,
                                                      const int* __restrict__ map, int map_len, float* __restrict__ mean,
                                                      float* __restrict__ rstd) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  constexpr int C = 24 * LPR, RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR;
  const int row = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
  const bool valid = row < rows;
  long long src = valid ? row : 0;
  if (!STATS && map) src = (src / map_len) * map_len + map[src % map_len];
  float v[24];
  load24(x + src * ld + sub * 24, v);
  float mu, rs;
  stats24<LPR>(v, eps, mu, rs);
  if (!valid) return;
  if (STATS) {
    if (sub == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
    return;
  }
  const int grp = (row / g_div) % g_mod;
  affine24(v, mu, rs, g + (long long)grp * C + sub * 24, b + (long long)grp * C + sub * 24);
  if (res) {
    float r[24];
    load24(res + (long long)row * C + sub * 24, r);
#pragma unroll
    for (int i = 0; i < 24; ++i) v[i] += r[i];
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) st8_bf16(y + (long long)row * C + sub * 24 + 8 * j, v + 8 * j);
}

// statistics of the PatchMerging gather (modules.py:282-287): merged row (b,i,j) = x[b,2i,2j] | x[b,2i+1,2j] |
// x[b,2i,2j+1] | x[b,2i+1,2j+1]; LPS lanes per source token, 4*LPS lanes per merged row
template <int LPS>
__global__ void __launch_bounds__(256) ln_stats_merge_fast_kernel(const bf16* __restrict__ x, int B, int H, int W, float eps,
                                                                  float* __restrict__ mean, float* __restrict__ rstd) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  constexpr int C = 24 * LPS, LPR = 4 * LPS, RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR, q = sub / LPS, part = sub % LPS;
  const int ho = H / 2, wo = W / 2, rows = B * ho * wo;
  const int row = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
  const bool valid = row < rows;
  const int rr = valid ? row : 0;
  const int bb = rr / (ho * wo), rem = rr % (ho * wo), i = rem / wo, j = rem % wo;
  const bf16* p = x + (((long long)bb * H + 2 * i + (q & 1)) * W + 2 * j + (q >> 1)) * C + part * 24;
  float v[24];
  load24(p, v);
  float mu, rs;
  stats24<LPR>(v, eps, mu, rs);
  if (valid && sub == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
}

// pe_combine (misc.cu) with 4 lanes per 96-channel token
__global__ void __launch_bounds__(256) pe_combine_fast_kernel(const bf16* __restrict__ c0, const bf16* __restrict__ c1, int B,
                                                              int P, int pad1, SjNorm n0, SjNorm n1, SjNorm nf,
                                                              bf16* __restrict__ y, float* __restrict__ st_mean,
                                                              float* __restrict__ st_rstd) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  const int lane = threadIdx.x & 31, sub = lane & 3;
  const long long ntok = (long long)B * P * P;
  const long long tok = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 8 + (lane >> 2);
  const bool valid = tok < ntok;
  const long long t = valid ? tok : 0;
  float a[24], mu, rs;
  load24(c0 + t * 96 + sub * 24, a);
  stats24<4>(a, 1e-5f, mu, rs);
  affine24(a, mu, rs, n0.g + sub * 24, n0.b + sub * 24);
  if (c1) {
    const int pj = (int)(t % P) - pad1, pi = (int)((t / P) % P) - pad1, P1 = P - 2 * pad1;
    const bool in = pi >= 0 && pj >= 0 && pi < P1 && pj < P1;  // uniform over the 4 lanes of a token
    const long long t1 = in ? ((t / ((long long)P * P)) * P1 + pi) * P1 + pj : 0;
    float m[24];
    load24(c1 + t1 * 96 + sub * 24, m);
    stats24<4>(m, 1e-5f, mu, rs);
    affine24(m, mu, rs, n1.g + sub * 24, n1.b + sub * 24);
    if (in) {
#pragma unroll
      for (int i = 0; i < 24; ++i) a[i] += m[i];
    }
  }
  stats24<4>(a, 1e-5f, mu, rs);
  affine24(a, mu, rs, nf.g + sub * 24, nf.b + sub * 24);
#pragma unroll
  for (int i = 0; i < 24; ++i) a[i] = __bfloat162float(__float2bfloat16_rn(a[i]));
  if (valid) {
#pragma unroll
    for (int j = 0; j < 3; ++j) st8_bf16(y + t * 96 + sub * 24 + 8 * j, a + 8 * j);
  }
  if (st_mean) {  // eps-1e-5 LayerNorm statistics of the stored row: norm1 of the first Swin block
    stats24<4>(a, 1e-5f, mu, rs);
    if (valid && sub == 0) {
      st_mean[tok] = mu;
      st_rstd[tok] = rs;
    }
  }
}

int lanes_per_row(int C) { return C == 96 ? 4 : (C == 192 ? 8 : (C == 384 ? 16 : 0)); }
bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// Each returns false when the shape / alignment is not covered (the caller falls back to the generic kernel).
bool ln_fast(Ctx& c, bool stats_only, const void* x, void* y, int rows, int C, int ld, const float* g, const float* b,
             float eps, const void* res, int g_div, int g_mod, const int* map, int map_len, float* mean, float* rstd) {
  const int lpr = lanes_per_row(C);
  if (c.dtype != SJ_BF16 || !lpr || ld % 8 || !al16(x) || (y && !al16(y)) || (res && !al16(res))) return false;
  if (!stats_only && ld != C) return false;
  const int rpb = 8 * (32 / lpr), grid = cdiv(rows, rpb);
#define SJ_LNF(LPR_, ST_)                                                                                              \
  SJ_LAUNCH(c, ST_ ? "ln_stats_fast" : "layernorm_fast", (ln_rows_kernel<LPR_, ST_>), grid, 256, 0, (const bf16*)x,     \
            (bf16*)y, rows, ld, g, b, eps, (const bf16*)res, g_div < 1 ? 1 : g_div, g_mod < 1 ? 1 : g_mod, map, map_len, \
            mean, rstd)
  if (stats_only) {
    if (lpr == 4) SJ_LNF(4, true); else if (lpr == 8) SJ_LNF(8, true); else SJ_LNF(16, true);
  } else {
    if (lpr == 4) SJ_LNF(4, false); else if (lpr == 8) SJ_LNF(8, false); else SJ_LNF(16, false);
  }
#undef SJ_LNF
  return true;
}

bool ln_stats_merge_fast(Ctx& c, const void* x, int B, int H, int W, int C, float eps, float* mean, float* rstd) {
  if (c.dtype != SJ_BF16 || (C != 96 && C != 192) || !al16(x)) return false;
  const int rows = B * (H / 2) * (W / 2);
  if (C == 96) SJ_LAUNCH(c, "ln_stats_merge_fast", ln_stats_merge_fast_kernel<4>, cdiv(rows, 16), 256, 0, (const bf16*)x, B, H, W, eps, mean, rstd);
  else SJ_LAUNCH(c, "ln_stats_merge_fast", ln_stats_merge_fast_kernel<8>, cdiv(rows, 8), 256, 0, (const bf16*)x, B, H, W, eps, mean, rstd);
  return true;
}

bool pe_combine_fast(Ctx& c, const void* c0, const void* c1, int B, int P, int pad1, const SjNorm& n0, const SjNorm& n1,
                     const SjNorm& nf, void* y, float* st_mean, float* st_rstd) {
  if (!al16(c0) || (c1 && !al16(c1)) || !al16(y)) return false;
  const long long ntok = (long long)B * P * P;
  SJ_LAUNCH(c, "pe_combine_fast", pe_combine_fast_kernel, cdiv(ntok, 64), 256, 0, (const bf16*)c0, (const bf16*)c1, B, P, pad1,
            n0, n1, nf, (bf16*)y, st_mean, st_rstd);
  return true;
}

}  // namespace sj
