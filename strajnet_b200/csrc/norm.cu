// LayerNormalization kernels (Keras semantics: last axis, biased variance).  One warp per row.
#include <cstdlib>

#include "kernels.h"

namespace sj {
namespace {

template <typename T>
__global__ void ln_stats_kernel(const T* __restrict__ x, int rows, int C, int ld, float eps, float* __restrict__ mean,
                                float* __restrict__ rstd) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  int lane = threadIdx.x % 32;
  if (row >= rows) return;
  const T* p = x + (long long)row * ld;
  float s = 0.f;
  for (int i = lane * 4; i < C; i += 128) {
    float4 v = ld4<T>(p + i);
    s += v.x + v.y + v.z + v.w;
  }
  float mu = warp_sum(s) / C;
  float q = 0.f;
  for (int i = lane * 4; i < C; i += 128) {
    float4 v = ld4<T>(p + i);
    float a = v.x - mu, b = v.y - mu, c2 = v.z - mu, d = v.w - mu;
    q += a * a + b * b + c2 * c2 + d * d;
  }
  float var = warp_sum(q) / C;
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rsqrtf(var + eps);
  }
}

// gathered row m = (b,i,j): concat of x[b,2i,2j], x[b,2i+1,2j], x[b,2i,2j+1], x[b,2i+1,2j+1]
template <typename T>
__global__ void ln_stats_merge_kernel(const T* __restrict__ x, int B, int H, int W, int C, float eps,
                                      float* __restrict__ mean, float* __restrict__ rstd) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  int lane = threadIdx.x % 32;
  int ho = H / 2, wo = W / 2;
  if (row >= B * ho * wo) return;
  int b = row / (ho * wo), rem = row % (ho * wo), i = rem / wo, j = rem % wo;
  const T* src[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) src[q] = x + (((long long)b * H + 2 * i + (q & 1)) * W + 2 * j + (q >> 1)) * C;
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    for (int k = lane * 4; k < C; k += 128) {
      float4 v = ld4<T>(src[q] + k);
      s += v.x + v.y + v.z + v.w;
    }
  float mu = warp_sum(s) / (4 * C);
  float qq = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    for (int k = lane * 4; k < C; k += 128) {
      float4 v = ld4<T>(src[q] + k);
      float a = v.x - mu, bb = v.y - mu, c2 = v.z - mu, d = v.w - mu;
      qq += a * a + bb * bb + c2 * c2 + d * d;
    }
  float var = warp_sum(qq) / (4 * C);
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rsqrtf(var + eps);
  }
}

template <typename T>
__global__ void layernorm_kernel(const T* __restrict__ x, T* __restrict__ y, int rows, int C, const float* __restrict__ g,
                                 const float* __restrict__ b, float eps, const T* __restrict__ res, int g_div, int g_mod,
                                 const int* __restrict__ map, int map_len) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  int lane = threadIdx.x % 32;
  if (row >= rows) return;
  long long src = row;
  if (map) src = (long long)(row / map_len) * map_len + map[row % map_len];
  const T* p = x + src * C;
  int grp = (row / g_div) % g_mod;
  const float* gg = g + (long long)grp * C;
  const float* bb = b + (long long)grp * C;
  float s = 0.f;
  for (int i = lane * 4; i < C; i += 128) {
    float4 v = ld4<T>(p + i);
    s += v.x + v.y + v.z + v.w;
  }
  float mu = warp_sum(s) / C;
  float q = 0.f;
  for (int i = lane * 4; i < C; i += 128) {
    float4 v = ld4<T>(p + i);
    float a = v.x - mu, bq = v.y - mu, c2 = v.z - mu, d = v.w - mu;
    q += a * a + bq * bq + c2 * c2 + d * d;
  }
  float rs = rsqrtf(warp_sum(q) / C + eps);
  for (int i = lane * 4; i < C; i += 128) {
    float4 v = ld4<T>(p + i);
    float4 g4 = *reinterpret_cast<const float4*>(gg + i), b4 = *reinterpret_cast<const float4*>(bb + i);
    float4 o;
    o.x = (v.x - mu) * rs * g4.x + b4.x;
    o.y = (v.y - mu) * rs * g4.y + b4.y;
    o.z = (v.z - mu) * rs * g4.z + b4.z;
    o.w = (v.w - mu) * rs * g4.w + b4.w;
    if (res) {
      float4 r = ld4<T>(res + (long long)row * C + i);
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    st4<T>(y + (long long)row * C + i, o);
  }
}

}  // namespace

static bool norm_fast_disabled() {
  static const bool off = getenv("SJ_DISABLE_NORM_FAST") != nullptr;
  return off;
}

void ln_stats(Ctx& c, const void* x, int rows, int C, int ld, float eps, float* mean, float* rstd) {
  if (!c.ok() || c.dry) return;
  if (C % 4 || ld % 4) { c.fail(SJ_EINVAL); return; }
  if (!norm_fast_disabled() && ln_fast(c, true, x, nullptr, rows, C, ld, nullptr, nullptr, eps, nullptr, 1, 1, nullptr, 0, mean, rstd)) return;
  int grid = cdiv(rows, 8);
  if (c.dtype == SJ_BF16) SJ_LAUNCH(c, "ln_stats", ln_stats_kernel<bf16>, grid, 256, 0, (const bf16*)x, rows, C, ld, eps, mean, rstd);
  else SJ_LAUNCH(c, "ln_stats", ln_stats_kernel<float>, grid, 256, 0, (const float*)x, rows, C, ld, eps, mean, rstd);
}

void ln_stats_merge(Ctx& c, const void* x, int B, int H, int W, int C, float eps, float* mean, float* rstd) {
  if (!c.ok() || c.dry) return;
  if (C % 4) { c.fail(SJ_EINVAL); return; }
  if (!norm_fast_disabled() && ln_stats_merge_fast(c, x, B, H, W, C, eps, mean, rstd)) return;
  int rows = B * (H / 2) * (W / 2);
  int grid = cdiv(rows, 8);
  if (c.dtype == SJ_BF16) SJ_LAUNCH(c, "ln_stats_merge", ln_stats_merge_kernel<bf16>, grid, 256, 0, (const bf16*)x, B, H, W, C, eps, mean, rstd);
  else SJ_LAUNCH(c, "ln_stats_merge", ln_stats_merge_kernel<float>, grid, 256, 0, (const float*)x, B, H, W, C, eps, mean, rstd);
}

void layernorm(Ctx& c, const void* x, void* y, int rows, int C, const float* g, const float* b, float eps,
               const void* res, int g_div, int g_mod) {
  if (!c.ok() || c.dry) return;
  if (C % 4) { c.fail(SJ_EINVAL); return; }
  if (g_div <= 0) g_div = 1;
  if (g_mod <= 0) g_mod = 1;
  if (!norm_fast_disabled() && ln_fast(c, false, x, y, rows, C, C, g, b, eps, res, g_div, g_mod, nullptr, 0, nullptr, nullptr)) return;
  int grid = cdiv(rows, 8);
  if (c.dtype == SJ_BF16)
    SJ_LAUNCH(c, "layernorm", layernorm_kernel<bf16>, grid, 256, 0, (const bf16*)x, (bf16*)y, rows, C, g, b, eps, (const bf16*)res, g_div, g_mod, (const int*)nullptr, 0);
  else
    SJ_LAUNCH(c, "layernorm", layernorm_kernel<float>, grid, 256, 0, (const float*)x, (float*)y, rows, C, g, b, eps, (const float*)res, g_div, g_mod, (const int*)nullptr, 0);
}

void layernorm_gather(Ctx& c, const void* x, void* y, int rows, int C, const float* g, const float* b, float eps,
                      const int* map, int map_len) {
  if (!c.ok() || c.dry) return;
  if (C % 4 || !map || map_len <= 0) { c.fail(SJ_EINVAL); return; }
  if (!norm_fast_disabled() && ln_fast(c, false, x, y, rows, C, C, g, b, eps, nullptr, 1, 1, map, map_len, nullptr, nullptr)) return;
  int grid = cdiv(rows, 8);
  if (c.dtype == SJ_BF16)
    SJ_LAUNCH(c, "layernorm_gather", layernorm_kernel<bf16>, grid, 256, 0, (const bf16*)x, (bf16*)y, rows, C, g, b, eps, (const bf16*)nullptr, 1, 1, map, map_len);
  else
    SJ_LAUNCH(c, "layernorm_gather", layernorm_kernel<float>, grid, 256, 0, (const float*)x, (float*)y, rows, C, g, b, eps, (const float*)nullptr, 1, 1, map, map_len);
}

}  // namespace sj
