// Attention cores of the fp32 path: one query per thread, K/V of the (batch, head) staged in
// shared memory and read as warp-wide broadcasts.
#include <cstdlib>

#include "kernels.h"

namespace sj {
namespace {

// ------------------------------------------------------------------------------------------------
// Window attention (modules.py:109-131): 64 tokens per window, one block per (window, head).
// ------------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(64) window_attn_kernel(const T* __restrict__ qkv, T* __restrict__ out,
                                                         const float* __restrict__ table, int C, int heads,
                                                         int mask_mode, int H, int W, int shift,
                                                         const float* __restrict__ mask, int nW) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  constexpr int DS = D + 4;
  __shared__ __align__(16) float Ks[64][DS];
  __shared__ __align__(16) float Vs[64][DS];
  __shared__ float tbl[225];
  __shared__ int rid[64];

  const int win = blockIdx.x, h = blockIdx.y, n = threadIdx.x;
  const T* row = qkv + ((long long)win * 64 + n) * (3 * C) + h * D;
  float q[D];
  const float scale = (D == 32) ? 0.17677669529663687f : 0.25f;  // head_dim ** -0.5 (modules.py:73)
#pragma unroll
  for (int d = 0; d < D; d += 4) {
    float4 a = ld4<T>(row + d), b = ld4<T>(row + C + d), c = ld4<T>(row + 2 * C + d);
    q[d] = a.x * scale; q[d + 1] = a.y * scale; q[d + 2] = a.z * scale; q[d + 3] = a.w * scale;
    *reinterpret_cast<float4*>(&Ks[n][d]) = b;
    *reinterpret_cast<float4*>(&Vs[n][d]) = c;
  }
  for (int i = n; i < 225; i += 64) tbl[i] = table[i * heads + h];
  int wloc = 0;
  if (mask_mode == 1) {
    // region ids of modules.py:192-203 evaluated at this token's position in the shifted image
    int nws = (H / 8) * (W / 8);
    wloc = win % nws;
    int y = (wloc / (W / 8)) * 8 + n / 8, x = (wloc % (W / 8)) * 8 + n % 8;
    rid[n] = shift_region_id(H, W, 8, shift, y, x);
  } else if (mask_mode == 2) {
    wloc = win % nW;
  }
  __syncthreads();

  float s[64];
  float mx = -INFINITY;
  const int nr = n / 8, nc = n % 8;
  const int myrid = (mask_mode == 1) ? rid[n] : 0;
#pragma unroll
  for (int m = 0; m < 64; ++m) {
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < D; d += 4) {
      float4 k4 = *reinterpret_cast<const float4*>(&Ks[m][d]);
      acc = fmaf(q[d], k4.x, acc);
      acc = fmaf(q[d + 1], k4.y, acc);
      acc = fmaf(q[d + 2], k4.z, acc);
      acc = fmaf(q[d + 3], k4.w, acc);
    }
    acc += tbl[(nr - m / 8 + 7) * 15 + (nc - m % 8 + 7)];
    if (mask_mode == 1) acc += (rid[m] != myrid) ? -100.0f : 0.0f;
    else if (mask_mode == 2) acc += mask[((long long)wloc * 64 + n) * 64 + m];
    s[m] = acc;
    mx = fmaxf(mx, acc);
  }
  float sum = 0.f;
#pragma unroll
  for (int m = 0; m < 64; ++m) {
    s[m] = expf(s[m] - mx);
    sum += s[m];
  }
  float o[D];
#pragma unroll
  for (int d = 0; d < D; ++d) o[d] = 0.f;
#pragma unroll
  for (int m = 0; m < 64; ++m) {
#pragma unroll
    for (int d = 0; d < D; d += 4) {
      float4 v4 = *reinterpret_cast<const float4*>(&Vs[m][d]);
      o[d] = fmaf(s[m], v4.x, o[d]);
      o[d + 1] = fmaf(s[m], v4.y, o[d + 1]);
      o[d + 2] = fmaf(s[m], v4.z, o[d + 2]);
      o[d + 3] = fmaf(s[m], v4.w, o[d + 3]);
    }
  }
  const float inv = 1.0f / sum;
  T* orow = out + ((long long)win * 64 + n) * C + h * D;
#pragma unroll
  for (int d = 0; d < D; d += 4)
    st4<T>(orow + d, make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv));
}

// ------------------------------------------------------------------------------------------------
// Generic MHA core with tfa.layers.MultiHeadAttention semantics, and the FG-MSA
// variant with the bilinearly sampled relative-position bias (FG_MSA.py:147-176).
// grid (q chunks, heads, batch); one query per thread; online softmax over keys.
// ------------------------------------------------------------------------------------------------
template <typename T, int D, bool FG>
__global__ void mha_kernel(const MhaP p) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  constexpr int DP = (D + 3) & ~3;
  extern __shared__ __align__(16) float smem[];
  float* Ks = smem;                    // [Nk][DP]
  float* Vs = Ks + (size_t)p.Nk * DP;  // [Nk][DP]
  float* Tp = Vs + (size_t)p.Nk * DP;  // FG: [33][33] zero-padded table, then pos [Nk][2]
  float* Ps = Tp + 33 * 33;
  int* Km = reinterpret_cast<int*>(FG ? Ps + 2 * p.Nk : Tp);  // [Nk] key mask

  const int h = blockIdx.y, b = blockIdx.z;
  const T* K = reinterpret_cast<const T*>(p.k) + (long long)b * p.Nk * p.ldk + h * D;
  const T* V = reinterpret_cast<const T*>(p.v) + (long long)b * p.Nk * p.ldv + h * D;
  for (int i = threadIdx.x; i < p.Nk * DP; i += blockDim.x) {
    int m = i / DP, d = i % DP;
    Ks[i] = d < D ? ldf<T>(K + (long long)m * p.ldk + d) : 0.f;
    Vs[i] = d < D ? ldf<T>(V + (long long)m * p.ldv + d) : 0.f;
  }
  if (FG) {
    for (int i = threadIdx.x; i < 33 * 33; i += blockDim.x) {
      int r = i / 33 - 1, cc = i % 33 - 1;
      Tp[i] = (r >= 0 && r < 31 && cc >= 0 && cc < 31) ? p.fg_table[(r * 31 + cc) * p.heads + h] : 0.f;
    }
    const float* pos = p.fg_pos + ((long long)b * p.heads + h) * p.Nk * 2;
    for (int i = threadIdx.x; i < 2 * p.Nk; i += blockDim.x) Ps[i] = pos[i];
  }
  const int mb = b / p.mask_div;
  if (p.kmask)
    for (int i = threadIdx.x; i < p.Nk; i += blockDim.x) Km[i] = p.kmask[(long long)mb * p.Nk + i];
  __syncthreads();

  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= p.Nq) return;
  const T* Q = reinterpret_cast<const T*>(p.q) + ((long long)b * p.Nq + qi) * p.ldq + h * D;
  float q[DP];
  const float sq = sqrtf((float)D);
#pragma unroll
  for (int d = 0; d < DP; ++d) {
    float v = d < D ? ldf<T>(Q + d) : 0.f;
    q[d] = FG ? v : v / sq;  // tfa: query /= sqrt(head_size) before the dot; FG-MSA scales after (:148)
  }
  const int qvalid = p.qmask ? p.qmask[(long long)mb * p.Nq + qi] : 1;
  const float fg_scale = 0.14433756729740643f;  // 48 ** -0.5 (FG_MSA.py:31)
  const float iq = (float)(qi / 16), jq = (float)(qi % 16);

  float mx = -INFINITY, sum = 0.f;
  float o[DP];
#pragma unroll
  for (int d = 0; d < DP; ++d) o[d] = 0.f;
  for (int m = 0; m < p.Nk; ++m) {
    const float* kr = Ks + (size_t)m * DP;
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < DP; d += 4) {
      float4 k4 = *reinterpret_cast<const float4*>(kr + d);
      acc = fmaf(q[d], k4.x, acc);
      acc = fmaf(q[d + 1], k4.y, acc);
      acc = fmaf(q[d + 2], k4.z, acc);
      acc = fmaf(q[d + 3], k4.w, acc);
    }
    if (FG) {
      acc *= fg_scale;
      // zero-border bilinear sample of the 31x31 table (occu_metric.py:394-409, tfa_image.py:116-171)
      float r = jq - Ps[2 * m] + 1.0f, cc = iq - Ps[2 * m + 1] + 1.0f;
      float r0 = fminf(fmaxf(floorf(r), 0.f), 31.f), c0 = fminf(fmaxf(floorf(cc), 0.f), 31.f);
      float ar = fminf(fmaxf(r - r0, 0.f), 1.f), ac = fminf(fmaxf(cc - c0, 0.f), 1.f);
      const float* t0 = Tp + (int)r0 * 33 + (int)c0;
      float tl = t0[0], tr = t0[1], bl = t0[33], br = t0[34];
      float top = ac * (tr - tl) + tl, bot = ac * (br - bl) + bl;
      acc += ar * (bot - top) + top;
    } else {
      int keep = qvalid & (p.kmask ? Km[m] : 1);
      if (!keep) acc = acc + (-10e9f);  // fp32 add as in tfa: absorbs the logit (Q7)
    }
    if (acc > mx) {
      float f = expf(mx - acc);  // exp(-inf) = 0 on the first key
      sum *= f;
#pragma unroll
      for (int d = 0; d < DP; ++d) o[d] *= f;
      mx = acc;
    }
    float pw = expf(acc - mx);
    sum += pw;
    const float* vr = Vs + (size_t)m * DP;
#pragma unroll
    for (int d = 0; d < DP; d += 4) {
      float4 v4 = *reinterpret_cast<const float4*>(vr + d);
      o[d] = fmaf(pw, v4.x, o[d]);
      o[d + 1] = fmaf(pw, v4.y, o[d + 1]);
      o[d + 2] = fmaf(pw, v4.z, o[d + 2]);
      o[d + 3] = fmaf(pw, v4.w, o[d + 3]);
    }
  }
  const float inv = 1.0f / sum;
  T* orow = reinterpret_cast<T*>(p.out) + ((long long)b * p.Nq + qi) * p.ldo;
#pragma unroll
  for (int d = 0; d < D; ++d) stf<T>(orow + h * D + d, o[d] * inv);
  if (h == p.heads - 1)
    for (int col = p.heads * D; col < p.ldo; ++col) stf<T>(orow + col, 0.f);
}

template <typename T, int D, bool FG>
void launch_mha(Ctx& c, const MhaP& p) {
  constexpr int DP = (D + 3) & ~3;
  int block = p.Nq >= 128 ? 128 : (p.Nq > 32 ? 64 : 32);
  dim3 grid(cdiv(p.Nq, block), p.heads, p.batch);
  size_t smem = (size_t)2 * p.Nk * DP * 4 + (FG ? (33 * 33 + 2 * p.Nk) * 4 : 0) + (size_t)p.Nk * 4;
  if (smem > 48 * 1024) {
    if (!SJ_SMEM_LIMIT_OK((mha_kernel<T, D, FG>), smem)) { c.fail(SJ_ECUDA); return; }
  }
  SJ_LAUNCH(c, "mha_core", (mha_kernel<T, D, FG>), grid, block, smem, p);
}

template <typename T>
void dispatch_mha(Ctx& c, const MhaP& p) {
  bool fg = p.fg_pos != nullptr;
  if (fg) {
    if (p.D == 48 && p.Nq == 256 && p.Nk == 256) launch_mha<T, 48, true>(c, p);
    else c.fail(SJ_EUNSUPPORTED);
    return;
  }
  switch (p.D) {
    case 42: launch_mha<T, 42, false>(c, p); break;
    case 64: launch_mha<T, 64, false>(c, p); break;
    default: c.fail(SJ_EUNSUPPORTED);
  }
}

}  // namespace

// A/B switch for the warp-MMA attention cores of the bf16 path (attn_mma.cu)
static bool attn_mma_disabled() {
  static const bool off = getenv("SJ_DISABLE_ATTN_MMA") != nullptr;
  return off;
}

void window_attn_core(Ctx& c, const void* qkv, void* out, const float* rpb_table, int n_windows_total, int C,
                      int heads, int mask_mode, int H, int W, int shift, const float* mask, int nW) {
  if (!c.ok() || c.dry) return;
  int D = C / heads;
  if (D * heads != C || (D != 16 && D != 32)) { c.fail(SJ_EUNSUPPORTED); return; }
  if (c.dtype == SJ_BF16 && !attn_mma_disabled() &&
      attn_mma_window(c, qkv, out, rpb_table, n_windows_total, C, heads, mask_mode, H, W, shift, mask, nW))
    return;
  dim3 grid(n_windows_total, heads);
#define SJ_WA(T, DD)                                                                                              \
  SJ_LAUNCH(c, "window_attn_core", (window_attn_kernel<T, DD>), grid, 64, 0, (const T*)qkv, (T*)out, rpb_table, C, \
            heads, mask_mode, H, W, shift, mask, nW)
  if (c.dtype == SJ_BF16) { if (D == 32) SJ_WA(bf16, 32); else SJ_WA(bf16, 16); }
  else { if (D == 32) SJ_WA(float, 32); else SJ_WA(float, 16); }
#undef SJ_WA
}

void mha_core(Ctx& c, const MhaP& p) {
  if (!c.ok() || c.dry) return;
  if (p.heads * p.D > p.ldo || p.batch <= 0 || p.Nq <= 0 || p.Nk <= 0) { c.fail(SJ_EINVAL); return; }
  if (c.dtype == SJ_BF16 && !attn_mma_disabled() && attn_mma_mha(c, p)) return;
  if (c.dtype == SJ_BF16) dispatch_mha<bf16>(c, p);
  else dispatch_mha<float>(c, p);
}

}  // namespace sj
