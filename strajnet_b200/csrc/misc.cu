// Index maps, data reshuffles, patch embedding, FG-MSA offset network, trajectory glue and the
// decoder head.  All HBM-/latency-bound CUDA-core work.
#include <cstdlib>

#include "kernels.h"

namespace sj {
namespace {

// ---------------------------------------------------------------- integer maps (bit-exact rows)
__global__ void rel_pos_index_kernel(int ws, int64_t* out) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  int N = ws * ws;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * N) return;
  int n = i / N, m = i % N;
  out[i] = (int64_t)((n / ws - m / ws + ws - 1) * (2 * ws - 1) + (n % ws - m % ws + ws - 1));
}

__global__ void shift_mask_kernel(int H, int W, int ws, int shift, float* out) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  int N = ws * ws;
  long long total = (long long)(H / ws) * (W / ws) * N * N;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int m = i % N, n = (i / N) % N, w = i / ((long long)N * N);
  int wy = (w / (W / ws)) * ws, wx = (w % (W / ws)) * ws;
  int a = shift_region_id(H, W, ws, shift, wy + n / ws, wx + n % ws);
  int b = shift_region_id(H, W, ws, shift, wy + m / ws, wx + m % ws);
  out[i] = (a != b) ? -100.0f : 0.0f;
}

__global__ void window_token_map_kernel(int H, int W, int ws, int shift, int32_t* out) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  int N = ws * ws;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  int w = i / N, n = i % N;
  int y = ((w / (W / ws)) * ws + n / ws + shift) % H;
  int x = ((w % (W / ws)) * ws + n % ws + shift) % W;
  out[i] = y * W + x;
}

// window_partition (scatter = 0) / window_reverse (scatter = 1) on data, modules.py:49-63
template <typename T>
__global__ void window_permute_kernel(const T* __restrict__ x, T* __restrict__ y, long long rows, int C, int H, int W,
                                      int ws, int scatter) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int c4 = C / 4;
  if (i >= rows * c4) return;
  long long r = i / c4;  // row in window order: ((b*nW + w)*ws*ws + n)
  int c = (i % c4) * 4;
  int L = H * W, N = ws * ws;
  int l = r % L;
  int w = l / N, n = l % N;
  int yy = (w / (W / ws)) * ws + n / ws, xx = (w % (W / ws)) * ws + n % ws;
  long long mr = (r / L) * L + yy * W + xx;
  if (scatter) st4<T>(y + mr * C + c, ld4<T>(x + r * C + c));
  else st4<T>(y + r * C + c, ld4<T>(x + mr * C + c));
}

template <typename T>
__global__ void center_crop_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int P, int C) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  int h = P / 2, o = P / 4, c4 = C / 4;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * h * h * c4) return;
  int c = (i % c4) * 4;
  long long t = i / c4;
  int xx = t % h, yy = (t / h) % h, b = t / ((long long)h * h);
  st4<T>(y + t * C + c, ld4<T>(x + (((long long)b * P + yy + o) * P + xx + o) * C + c));
}

// ---------------------------------------------------------------- fused patch embedding (K3)
constexpr int PE_TOK = 16;
constexpr int PE_KMAX = 16 * 11;

template <typename T>
__global__ void patch_embed_kernel(const PatchEmbedP p) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  __shared__ __align__(16) float ins[PE_TOK][PE_KMAX];
  __shared__ float vals[PE_TOK][128];
  __shared__ float stat[PE_TOK][2];
  const int n = threadIdx.x, E = p.E;
  const int P = p.S[0] / 4;
  const long long tok0 = (long long)blockIdx.x * PE_TOK;
  const long long ntok = (long long)p.B * P * P;
  const int nwarps = blockDim.x / 32, warp = n / 32, lane = n % 32;
  float total[PE_TOK];
#pragma unroll
  for (int t = 0; t < PE_TOK; ++t) total[t] = 0.f;

  auto row_stats = [&]() {  // LN statistics of vals[t][0..E) for each token, one warp per token
    __syncthreads();
    for (int t = warp; t < PE_TOK; t += nwarps) {
      float s = 0.f;
      for (int i = lane; i < E; i += 32) s += vals[t][i];
      float mu = warp_sum(s) / E;
      float q = 0.f;
      for (int i = lane; i < E; i += 32) {
        float d = vals[t][i] - mu;
        q += d * d;
      }
      float var = warp_sum(q) / E;
      if (lane == 0) {
        stat[t][0] = mu;
        stat[t][1] = rsqrtf(var + 1e-5f);
      }
    }
    __syncthreads();
  };

  for (int in = 0; in < p.n_in; ++in) {
    const int Cin = p.Cin[in], es = p.es[in], S = p.S[in], K = 16 * Cin;
    const int pad = in == 1 ? p.pad1 : 0;
    __syncthreads();
    // gather: each (token, patch row) is one contiguous run of 4*Cin channels in the NHWC input; a warp takes
    // a run, lanes stride over it (no per-element index arithmetic)
    const int run = 4 * Cin;
    for (int pr = warp; pr < PE_TOK * 4; pr += nwarps) {
      const int t = pr >> 2, ky = pr & 3;
      const long long tok = tok0 + t;
      long long src = -1;  // element index of the run start
      if (tok < ntok) {
        const int pj = (int)(tok % P) - pad, pi = (int)((tok / P) % P) - pad;
        const long long b = tok / ((long long)P * P);
        if (pi >= 0 && pj >= 0 && pi < S / 4 && pj < S / 4) src = (((b * S + 4 * pi + ky) * S + 4 * pj) * Cin) * es;
      }
      for (int j = lane; j < run; j += 32)
        ins[t][ky * run + j] = src >= 0 ? load_input(p.img[in], src + (long long)j * es, p.itype[in]) : 0.f;
    }
    __syncthreads();
    float acc[PE_TOK];
#pragma unroll
    for (int t = 0; t < PE_TOK; ++t) acc[t] = 0.f;
    if (n < E) {
      const float* w = p.w[in] + n;
      for (int k = 0; k < K; k += 4) {  // 4 weights per step, each reused for all PE_TOK tokens (broadcast LDS.128)
        const float w0 = w[(long long)k * E], w1 = w[(long long)(k + 1) * E], w2 = w[(long long)(k + 2) * E],
                    w3 = w[(long long)(k + 3) * E];
#pragma unroll
        for (int t = 0; t < PE_TOK; ++t) {
          const float4 x4 = *reinterpret_cast<const float4*>(&ins[t][k]);
          acc[t] = fmaf(x4.x, w0, acc[t]);
          acc[t] = fmaf(x4.y, w1, acc[t]);
          acc[t] = fmaf(x4.z, w2, acc[t]);
          acc[t] = fmaf(x4.w, w3, acc[t]);
        }
      }
      float bv = p.bias[in][n];
#pragma unroll
      for (int t = 0; t < PE_TOK; ++t) vals[t][n] = acc[t] + bv;
    }
    row_stats();
    if (n < E) {
      float gv = p.g[in][n], bv = p.b[in][n];
#pragma unroll
      for (int t = 0; t < PE_TOK; ++t) {
        long long tok = tok0 + t;
        bool inside = true;
        if (pad > 0 && tok < ntok) {
          int pj = tok % P - pad, pi = (tok / P) % P - pad;
          inside = pi >= 0 && pj >= 0 && pi < S / 4 && pj < S / 4;
        }
        if (inside) total[t] += (vals[t][n] - stat[t][0]) * stat[t][1] * gv + bv;
      }
    }
  }
  if (p.gf) {
    __syncthreads();
    if (n < E)
#pragma unroll
      for (int t = 0; t < PE_TOK; ++t) vals[t][n] = total[t];
    row_stats();
    if (n < E) {
      float gv = p.gf[n], bv = p.bf[n];
#pragma unroll
      for (int t = 0; t < PE_TOK; ++t) total[t] = (vals[t][n] - stat[t][0]) * stat[t][1] * gv + bv;
    }
  }
  if (n < E) {
    T* y = reinterpret_cast<T*>(p.y);
#pragma unroll
    for (int t = 0; t < PE_TOK; ++t)
      if (tok0 + t < ntok) stf<T>(y + (tok0 + t) * E + n, total[t]);
  }
}

// ---------------------------------------------------------------- tensor-core patch embedding helpers (bf16 mode)
// im2col of the 4x4/stride-4 patches into a bf16 matrix A[token][Kpad] (k = (ky*4+kx)*Cin + c, zero padded to Kpad),
// so the conv becomes a plain tcgen05 GEMM.  One thread per 8 consecutive k (one 16-byte store).
__global__ void im2col4_kernel(const void* __restrict__ img, int itype, int B, int S, int Cin, int es, int Kpad,
                               bf16* __restrict__ A) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  const int k8n = Kpad / 8, P = S / 4, K = 16 * Cin;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * P * P * k8n) return;
  const long long tok = i / k8n;
  const int k0 = (int)(i % k8n) * 8;
  const int pj = (int)(tok % P), pi = (int)((tok / P) % P);
  const long long b = tok / ((long long)P * P);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = k0 + j;
    float x = 0.f;
    if (k < K) {
      const int c = k % Cin, kx = (k / Cin) & 3, ky = k / (4 * Cin);
      x = load_input(img, ((((b * S + 4 * pi + ky) * S + 4 * pj + kx) * Cin) + c) * (long long)es, itype);
    }
    v[j] = x;
  }
  __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
  u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
  *reinterpret_cast<uint4*>(A + tok * Kpad + k0) = u;
}

// Staged variant: one block = 64 consecutive tokens of one token row.  The four source image rows of those tokens are
// contiguous spans: they are read with 16-byte loads (fully coalesced), converted, and the kept elements (channel
// element stride es: plane 0 of the [..., 11, 2] occupancy raster) parked in shared memory as bf16; the A rows are then
// written with 16-byte stores.  The gather kernel above reads 8 scattered scalars per thread instead.
constexpr int IM2COL_TOK = 64;
template <int ITYPE>
__global__ void __launch_bounds__(256) im2col4_staged_kernel(const void* __restrict__ img, int B, int S, int Cin, int es,
                                                             int Kpad, bf16* __restrict__ A) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  extern __shared__ __align__(16) bf16 im2col_sm[];  // [4 ky][IM2COL_TOK * 4 * Cin]
  const int P = S / 4, blocks_per_row = P / IM2COL_TOK;
  const int pj0 = (blockIdx.x % blocks_per_row) * IM2COL_TOK, pi = (blockIdx.x / blocks_per_row) % P;
  const long long b = blockIdx.x / ((long long)blocks_per_row * P);
  const int keep_row = IM2COL_TOK * 4 * Cin;      // kept elements per image row
  const int span = keep_row * es;                 // source elements per image row
  constexpr int EPV = ITYPE == IN_F32 ? 4 : 16;   // source elements per 16-byte load
  for (int ky = 0; ky < 4; ++ky) {
    const long long base = (((b * S + 4 * pi + ky) * S) + 4 * pj0) * (long long)Cin * es;
    for (int v = threadIdx.x; v < span / EPV; v += 256) {
      const int e0 = v * EPV;
      if (ITYPE == IN_F32) {
        const float4 f = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(img) + base + e0);
        const float vals[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if ((e0 + i) % es == 0) im2col_sm[ky * keep_row + (e0 + i) / es] = __float2bfloat16_rn(vals[i]);
      } else {
        const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(img) + base + e0);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if ((e0 + i) % es == 0) {
            const uint32_t byte = (w[i >> 2] >> (8 * (i & 3))) & 0xff;
            const float x = ITYPE == IN_U8 ? (byte != 0 ? 1.0f : 0.0f) : (float)(int8_t)byte / 256.0f;
            im2col_sm[ky * keep_row + (e0 + i) / es] = __float2bfloat16_rn(x);
          }
        }
      }
    }
  }
  __syncthreads();
  const int k8n = Kpad / 8, K = 16 * Cin, kpr = 4 * Cin;  // kpr: k values per image row of a token
  for (int i = threadIdx.x; i < IM2COL_TOK * k8n; i += 256) {
    const int tokl = i / k8n, k0 = (i % k8n) * 8;
    __align__(16) bf16 o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + j;
      o[j] = k < K ? im2col_sm[(k / kpr) * keep_row + tokl * kpr + k % kpr] : __float2bfloat16_rn(0.f);
    }
    const long long tok = (b * P + pi) * P + pj0 + tokl;
    *reinterpret_cast<uint4*>(A + tok * Kpad + k0) = *reinterpret_cast<const uint4*>(o);
  }
}

// y[token] = LN_f( LN_0(c0[token]) + LN_1(c1[token']) ) over E = 96 channels, eps 1e-5 (modules.py:445, :580-587, :602);
// c1 is optional and may cover only the centre P1 x P1 tokens of the P x P grid (pad1 = (P - P1)/2), contributing 0
// elsewhere.  One warp per token, 3 channels per lane.
__global__ void pe_combine_kernel(const bf16* __restrict__ c0, const bf16* __restrict__ c1, int B, int P, int pad1,
                                  SjNorm n0, SjNorm n1, SjNorm nf, bf16* __restrict__ y, float* __restrict__ st_mean,
                                  float* __restrict__ st_rstd) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  const long long tok = (long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (tok >= (long long)B * P * P) return;
  auto ln3 = [&](float (&v)[3], const SjNorm& nm) {
    float mu = warp_sum(v[0] + v[1] + v[2]) / 96.f;
    float d0 = v[0] - mu, d1 = v[1] - mu, d2 = v[2] - mu;
    float rs = rsqrtf(warp_sum(d0 * d0 + d1 * d1 + d2 * d2) / 96.f + 1e-5f);
#pragma unroll
    for (int j = 0; j < 3; ++j) v[j] = (v[j] - mu) * rs * nm.g[lane + 32 * j] + nm.b[lane + 32 * j];
  };
  float a[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) a[j] = __bfloat162float(c0[tok * 96 + lane + 32 * j]);
  ln3(a, n0);
  if (c1) {
    const int pj = (int)(tok % P) - pad1, pi = (int)((tok / P) % P) - pad1, P1 = P - 2 * pad1;
    if (pi >= 0 && pj >= 0 && pi < P1 && pj < P1) {
      const long long t1 = ((tok / ((long long)P * P)) * P1 + pi) * P1 + pj;
      float m[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) m[j] = __bfloat162float(c1[t1 * 96 + lane + 32 * j]);
      ln3(m, n1);
#pragma unroll
      for (int j = 0; j < 3; ++j) a[j] += m[j];
    }
  }
  ln3(a, nf);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const bf16 r = __float2bfloat16_rn(a[j]);
    y[tok * 96 + lane + 32 * j] = r;
    a[j] = __bfloat162float(r);
  }
  if (st_mean) {  // eps-1e-5 LayerNorm statistics of the stored row: norm1 of the first Swin block
    const float mu = warp_sum(a[0] + a[1] + a[2]) / 96.f;
    const float d0 = a[0] - mu, d1 = a[1] - mu, d2 = a[2] - mu;
    const float var = warp_sum(d0 * d0 + d1 * d1 + d2 * d2) / 96.f;
    if (lane == 0) {
      st_mean[tok] = mu;
      st_rstd[tok] = rsqrtf(var + 1e-5f);
    }
  }
}

// ---------------------------------------------------------------- FG-MSA offset network
// one block (384 threads = output channels) per (b, image row i, 8-pixel half row): the grouped 3x3 conv
// weights are read once per block and reused for the 8 pixels
constexpr int FGP = 8;
template <typename T>
__global__ void __launch_bounds__(384) fg_offset_kernel(const T* __restrict__ q, int ldq, SjFgmsaW w,
                                                        float* __restrict__ off, float* __restrict__ pos) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  __shared__ __align__(16) float qs[3][FGP + 2][384];
  float (*us)[384] = reinterpret_cast<float (*)[384]>(&qs[0][0][0]);  // aliases qs: dead after the conv
  __shared__ float red[12][FGP];
  __shared__ float stat[FGP][2];
  const int n = threadIdx.x, b = blockIdx.x / 32, rem = blockIdx.x % 32, i = rem / 2, j0 = (rem % 2) * FGP;
  for (int r = 0; r < 3; ++r)
    for (int cc = 0; cc < FGP + 2; ++cc) {
      int yy = i + r - 1, xx = j0 + cc - 1;
      qs[r][cc][n] = (yy >= 0 && yy < 16 && xx >= 0 && xx < 16) ? ldf<T>(q + ((long long)b * 256 + yy * 16 + xx) * ldq + n) : 0.f;
    }
  __syncthreads();
  const int g = n / 48;
  float acc[FGP];
#pragma unroll
  for (int pz = 0; pz < FGP; ++pz) acc[pz] = w.conv0_b[n];
  for (int tap = 0; tap < 9; ++tap) {
    const float* wt = w.conv0_w + (long long)tap * 48 * 384 + n;
    const int r = tap / 3, dx = tap % 3;
    for (int cc = 0; cc < 48; cc += 4) {
      const float w0 = wt[cc * 384], w1 = wt[(cc + 1) * 384], w2 = wt[(cc + 2) * 384], w3 = wt[(cc + 3) * 384];
#pragma unroll
      for (int pz = 0; pz < FGP; ++pz) {
        const float4 x4 = *reinterpret_cast<const float4*>(&qs[r][pz + dx][g * 48 + cc]);
        acc[pz] = fmaf(x4.x, w0, acc[pz]);
        acc[pz] = fmaf(x4.y, w1, acc[pz]);
        acc[pz] = fmaf(x4.z, w2, acc[pz]);
        acc[pz] = fmaf(x4.w, w3, acc[pz]);
      }
    }
  }
  // LayerNorm over 384 channels, eps 1e-3 (Keras default), then tanh-GELU
  const int warp = n / 32, lane = n % 32;
#pragma unroll
  for (int pz = 0; pz < FGP; ++pz) {
    float s = warp_sum(acc[pz]);
    if (lane == 0) red[warp][pz] = s;
  }
  __syncthreads();
  if (n < FGP) {
    float mu = 0.f;
    for (int k = 0; k < 12; ++k) mu += red[k][n];
    stat[n][0] = mu / 384.f;
  }
  __syncthreads();
#pragma unroll
  for (int pz = 0; pz < FGP; ++pz) {
    float d = acc[pz] - stat[pz][0];
    float s2 = warp_sum(d * d);
    if (lane == 0) red[warp][pz] = s2;
  }
  __syncthreads();
  if (n < FGP) {
    float var = 0.f;
    for (int k = 0; k < 12; ++k) var += red[k][n];
    stat[n][1] = rsqrtf(var / 384.f + 1e-3f);
  }
  __syncthreads();
  const float gam = w.conv_norm.g[n], bet = w.conv_norm.b[n];
#pragma unroll
  for (int pz = 0; pz < FGP; ++pz) us[pz][n] = gelu_tanh((acc[pz] - stat[pz][0]) * stat[pz][1] * gam + bet);
  __syncthreads();
  if (n < FGP * 16) {
    const int pz = n / 16, gg = (n % 16) / 2, o = n % 2, j = j0 + pz, pix = i * 16 + j;
    float a = 0.f;
    for (int cc = 0; cc < 48; ++cc) a = fmaf(us[pz][gg * 48 + cc], w.offproj_w[cc * 2 + o], a);
    a = tanhf(a) * 8.0f;  // offset_range = (Hk/2, Wk/2) = (8, 8), FG_MSA.py:115-117
    long long idx = (((long long)b * 8 + gg) * 256 + pix) * 2 + o;
    off[idx] = a;
    pos[idx] = a + (o == 0 ? (float)j : (float)i);  // tf.meshgrid 'xy': ref[i,j] = (j, i)
  }
}

template <typename T>
__global__ void build_query_kernel(const T* __restrict__ q2, const float* __restrict__ off, const float* __restrict__ w2,
                                   const float* __restrict__ b2, int B, int fg, T* __restrict__ query) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over B*8*256*96 float4 groups
  if (i >= (long long)B * 8 * 256 * 96) return;
  int c = (i % 96) * 4;
  long long row = i / 96;  // (b*8 + t)*256 + l
  int l = row % 256;
  long long b = row / 2048;
  float4 v = q2 ? ld4<T>(q2 + (b * 256 + l) * 384 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  if (fg) {
    float o0 = off[row * 2], o1 = off[row * 2 + 1];
    float4 w0 = *reinterpret_cast<const float4*>(w2 + c), w1 = *reinterpret_cast<const float4*>(w2 + 384 + c);
    float4 bb = *reinterpret_cast<const float4*>(b2 + c);
    // flow_hidden = off . Wp2 + bp2 is formed first, then added to the query (modules.py:830-831)
    v.x += fmaf(o1, w1.x, o0 * w0.x) + bb.x;
    v.y += fmaf(o1, w1.y, o0 * w0.y) + bb.y;
    v.z += fmaf(o1, w1.z, o0 * w0.z) + bb.z;
    v.w += fmaf(o1, w1.w, o0 * w0.w) + bb.w;
  }
  st4<T>(query + row * 384 + c, v);
}

__device__ __forceinline__ void ld8_bf16(const bf16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}
__device__ __forceinline__ void st8_bf16(bf16* p, const float* f) {
  uint4 u;
  __nv_bfloat162 h;
  h = __floats2bfloat162_rn(f[0], f[1]); u.x = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(f[2], f[3]); u.y = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(f[4], f[5]); u.z = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(f[6], f[7]); u.w = *reinterpret_cast<uint32_t*>(&h);
  *reinterpret_cast<uint4*>(p) = u;
}

// bf16 form: a thread owns 8 channels of one (sample, token) and writes them for all 8 waypoints -- the encoder feature and
// the projection weights of those channels are loaded once instead of eight times (the generic kernel above: one element
// group per thread, five dependent loads each, 16 us for a 25 MB output)
__global__ void __launch_bounds__(256) build_query_bf16_kernel(const bf16* __restrict__ q2, const float* __restrict__ off,
                                                               const float* __restrict__ w2, const float* __restrict__ b2,
                                                               int B, int fg, bf16* __restrict__ query) {
  pdl_wait();
  pdl_trigger();
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;  // over B*256 tokens x 48 groups of 8 channels
  if (i >= (uint32_t)B * 256u * 48u) return;
  const uint32_t tok = i / 48u, c = (i - tok * 48u) * 8u;  // tok = b*256 + l
  const uint32_t b = tok >> 8, l = tok & 255u;
  float base[8], w0[8], w1[8];
  if (q2) ld8_bf16(q2 + (size_t)tok * 384 + c, base);
  else {
#pragma unroll
    for (int j = 0; j < 8; ++j) base[j] = 0.f;
  }
  if (fg) {
#pragma unroll
    for (int j = 0; j < 8; j += 4) {
      const float4 a = *reinterpret_cast<const float4*>(w2 + c + j), d = *reinterpret_cast<const float4*>(w2 + 384 + c + j);
      w0[j] = a.x; w0[j + 1] = a.y; w0[j + 2] = a.z; w0[j + 3] = a.w;
      w1[j] = d.x; w1[j + 1] = d.y; w1[j + 2] = d.z; w1[j + 3] = d.w;
    }
  }
  float bb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bb[j] = fg ? b2[c + j] : 0.f;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const size_t row = ((size_t)b * 8 + t) * 256 + l;
    float v[8];
    if (fg) {
      const float2 o = *reinterpret_cast<const float2*>(off + row * 2);
      // flow_hidden = off . Wp2 + bp2 is formed first, then added to the query (modules.py:830-831)
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = base[j] + (fmaf(o.y, w1[j], o.x * w0[j]) + bb[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = base[j];
    }
    st8_bf16(query + row * 384 + c, v);
  }
}

// ---------------------------------------------------------------- trajectory glue
// one block (64 threads) per actor: node features, step masks, type embedding
template <typename T>
__global__ void traj_node_kernel(const float* __restrict__ obs, const float* __restrict__ occ, SjTrajW w, int B,
                                 T* __restrict__ node, int* __restrict__ stepmask, int* __restrict__ cmask,
                                 float* __restrict__ vec) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  __shared__ float X[88];
  const int a = blockIdx.x, b = a / 64, i = a % 64, n = threadIdx.x;
  const float* src = i < 48 ? obs + ((long long)b * 48 + i) * 88 : occ + ((long long)b * 16 + (i - 48)) * 88;
  for (int k = n; k < 88; k += 64) X[k] = src[k];
  __syncthreads();
  if (n < 11) stepmask[a * 11 + n] = X[n * 8] != 0.f;
  if (n == 0) {
    int any = 0;
    for (int t = 0; t < 11; ++t) any |= X[t * 8] != 0.f;
    cmask[a] = any;
  }
  float wn[5];
#pragma unroll
  for (int cc = 0; cc < 5; ++cc) wn[cc] = w.node_w[cc * 64 + n];
  float bn = w.node_b[n];
  for (int t = 0; t < 11; ++t) {
    float acc = 0.f;
#pragma unroll
    for (int cc = 0; cc < 5; ++cc) acc = fmaf(X[t * 8 + cc], wn[cc], acc);
    stf<T>(node + ((long long)a * 11 + t) * 64 + n, elu1(acc + bn));
  }
  float v = 0.f;
#pragma unroll
  for (int cc = 0; cc < 3; ++cc) v = fmaf(X[5 + cc], w.vec_w[cc * 64 + n], v);
  vec[(long long)a * 64 + n] = v;
}

template <typename T>
__global__ void traj_pool_concat_kernel(const T* __restrict__ proj, const float* __restrict__ vec, T* __restrict__ cat) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  const int a = blockIdx.x, n = threadIdx.x;  // 384 threads
  float v;
  if (n < 320) {
    v = -INFINITY;
    for (int t = 0; t < 11; ++t) v = fmaxf(v, ldf<T>(proj + ((long long)a * 11 + t) * 320 + n));  // mask-unaware (Q6)
  } else {
    v = vec[(long long)a * 64 + n - 320];
  }
  stf<T>(cat + (long long)a * 384 + n, v);
}

template <typename T>
__global__ void traj_prep_kernel(const T* __restrict__ E, const int* __restrict__ cmask, const float* __restrict__ seg_w,
                                 T* __restrict__ A, T* __restrict__ Q) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  const int a = blockIdx.x, n = threadIdx.x;  // 384 threads
  float e = ldf<T>(E + (long long)a * 384 + n) * (float)cmask[a];
  float s = seg_w[((a % 64) < 48 ? 0 : 384) + n];
  stf<T>(A + (long long)a * 384 + n, e);
  stf<T>(Q + (long long)a * 384 + n, e + s);
}

// key[a] = LN(E[a] + LN(F2[a]; ia_norm2) + seg; obs_norm | occ_norm), eps 1e-3.  One warp per actor.
template <typename T>
__global__ void traj_final_kernel(const T* __restrict__ E, const T* __restrict__ F2, SjTrajW w, int n_actors,
                                  T* __restrict__ key) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  int a = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (a >= n_actors) return;
  float f[12], s = 0.f;
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    f[k] = ldf<T>(F2 + (long long)a * 384 + lane + 32 * k);
    s += f[k];
  }
  float mu = warp_sum(s) / 384.f, q = 0.f;
#pragma unroll
  for (int k = 0; k < 12; ++k) q += (f[k] - mu) * (f[k] - mu);
  float rs = rsqrtf(warp_sum(q) / 384.f + 1e-3f);
  const bool is_obs = (a % 64) < 48;
  const SjNorm nm = is_obs ? w.obs_norm : w.occ_norm;
  s = 0.f;
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    int n = lane + 32 * k;
    float val = (f[k] - mu) * rs * w.ia_norm2.g[n] + w.ia_norm2.b[n];
    // obs = obs + val_obs (trajNet.py:179), then obs + embed (:184)
    f[k] = (ldf<T>(E + (long long)a * 384 + n) + val) + w.seg_w[(is_obs ? 0 : 384) + n];
    s += f[k];
  }
  mu = warp_sum(s) / 384.f;
  q = 0.f;
#pragma unroll
  for (int k = 0; k < 12; ++k) q += (f[k] - mu) * (f[k] - mu);
  rs = rsqrtf(warp_sum(q) / 384.f + 1e-3f);
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    int n = lane + 32 * k;
    stf<T>(key + (long long)a * 384 + n, (f[k] - mu) * rs * nm.g[n] + nm.b[n]);
  }
}

// ---------------------------------------------------------------- decoder head (K10)
// Two 3x3 48->2 convs (output_layer on x, output_layer_f on fx; modules.py:767-770) + the final
// [B,8,256,256,4] -> [B,256,256,32] transpose (:838).  HBM/FMA-bound CUDA-core kernel: one block per
// (sample, 16x32 pixel tile) loops over the 8 waypoints x 2 heads, staging each 18x34 halo tile in
// shared memory (112-byte pixel stride: conflict-free 16-byte reads), 4 pixels per thread, and writes
// every pixel's 32 output channels as one contiguous 128-byte line.
constexpr int OC_TY = 16, OC_TX = 32, OC_PSTRIDE = 56;  // pixel stride in bf16 elements (112 B)

template <typename T>
__global__ void __launch_bounds__(128) out_conv_kernel(const T* __restrict__ xo, const T* __restrict__ xf,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       int out_layout, void* __restrict__ outv) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  extern __shared__ __align__(16) uint8_t oc_smem[];
  T* tile = reinterpret_cast<T*>(oc_smem);                                                       // [18][34][56]
  float2* ws = reinterpret_cast<float2*>(oc_smem + (OC_TY + 2) * (OC_TX + 2) * OC_PSTRIDE * sizeof(T));  // [2][432]
  const int tid = threadIdx.x, ty = tid / 8, tx = tid % 8;
  const int b = blockIdx.z, y0 = blockIdx.y * OC_TY, x0 = blockIdx.x * OC_TX;
  for (int i = tid; i < 864; i += 128) ws[i] = make_float2(w[2 * i], w[2 * i + 1]);
  float acc[4][32];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[j][k] = 0.f;

#pragma unroll
  for (int th = 0; th < 16; ++th) {  // (waypoint t, head): fully unrolled so acc[][] stays in registers
    const int t = th >> 1, head = th & 1;
    const T* src = (head == 0 ? xo : xf) + ((long long)(b * 8 + t) * 65536) * 48;
    __syncthreads();
    constexpr int VE = 16 / sizeof(T);  // elements per 16-byte vector
    constexpr int VPP = 48 / VE;        // vectors per pixel
    for (int i = tid; i < (OC_TY + 2) * (OC_TX + 2) * VPP; i += 128) {
      const int ch = i % VPP, pix = i / VPP, px = pix % (OC_TX + 2), py = pix / (OC_TX + 2);
      const int yy = y0 + py - 1, xx = x0 + px - 1;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (yy >= 0 && yy < 256 && xx >= 0 && xx < 256)
        v = *reinterpret_cast<const uint4*>(src + ((long long)yy * 256 + xx) * 48 + ch * VE);
      *reinterpret_cast<uint4*>(tile + pix * OC_PSTRIDE + ch * VE) = v;
    }
    __syncthreads();
    float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
    const float2* wh = ws + head * 432;
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      const T* row = tile + ((ty + tap / 3) * (OC_TX + 2) + tx + tap % 3) * OC_PSTRIDE;
#pragma unroll
      for (int ch = 0; ch < 48; ch += 8) {
        float xv[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 lo = ld4<T>(row + j * 8 * OC_PSTRIDE + ch), hi = ld4<T>(row + j * 8 * OC_PSTRIDE + ch + 4);
          xv[j][0] = lo.x; xv[j][1] = lo.y; xv[j][2] = lo.z; xv[j][3] = lo.w;
          xv[j][4] = hi.x; xv[j][5] = hi.y; xv[j][6] = hi.z; xv[j][7] = hi.w;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float2 wv = wh[tap * 48 + ch + e];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            a0[j] = fmaf(xv[j][e], wv.x, a0[j]);
            a1[j] = fmaf(xv[j][e], wv.y, a1[j]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[j][t * 4 + head * 2] = a0[j] + bias[head * 2];
      acc[j][t * 4 + head * 2 + 1] = a1[j] + bias[head * 2 + 1];
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int y = y0 + ty, x = x0 + tx + 8 * j;
    float* out = reinterpret_cast<float*>(outv);
    if (out_layout == 2) {
      uint32_t q[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) q[t] = quantize_waypoint(acc[j][4 * t], acc[j][4 * t + 1], acc[j][4 * t + 2], acc[j][4 * t + 3]);
      uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(outv) + (((long long)b * 256 + y) * 256 + x) * 32);
      o[0] = make_uint4(q[0], q[1], q[2], q[3]);
      o[1] = make_uint4(q[4], q[5], q[6], q[7]);
    } else if (out_layout == 1) {
      float* o = out + (((long long)b * 256 + y) * 256 + x) * 32;
#pragma unroll
      for (int k = 0; k < 32; k += 4)
        *reinterpret_cast<float4*>(o + k) = make_float4(acc[j][k], acc[j][k + 1], acc[j][k + 2], acc[j][k + 3]);
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t)
        *reinterpret_cast<float4*>(out + ((((long long)b * 8 + t) * 256 + y) * 256 + x) * 4) =
            make_float4(acc[j][t * 4], acc[j][t * 4 + 1], acc[j][t * 4 + 2], acc[j][t * 4 + 3]);
    }
  }
}

}  // namespace

// ---------------------------------------------------------------- launchers
void relative_position_index(Ctx& c, int ws, int64_t* out) {
  if (!c.ok() || c.dry) return;
  int n = ws * ws * ws * ws;
  SJ_LAUNCH(c, "rel_pos_index", rel_pos_index_kernel, cdiv(n, 256), 256, 0, ws, out);
}
void shift_attn_mask(Ctx& c, int H, int W, int ws, int shift, float* out) {
  if (!c.ok() || c.dry) return;
  long long total = (long long)(H / ws) * (W / ws) * ws * ws * ws * ws;
  SJ_LAUNCH(c, "shift_mask", shift_mask_kernel, cdiv(total, 256), 256, 0, H, W, ws, shift, out);
}
void window_token_map(Ctx& c, int H, int W, int ws, int shift, int32_t* out) {
  if (!c.ok() || c.dry) return;
  SJ_LAUNCH(c, "window_token_map", window_token_map_kernel, cdiv(H * W, 256), 256, 0, H, W, ws, shift, out);
}
void window_permute(Ctx& c, const void* x, void* y, int B, int H, int W, int C, int ws, int scatter) {
  if (!c.ok() || c.dry) return;
  if (C % 4) { c.fail(SJ_EINVAL); return; }
  long long rows = (long long)B * H * W, n = rows * (C / 4);
  if (c.dtype == SJ_BF16) SJ_LAUNCH(c, "window_permute", window_permute_kernel<bf16>, cdiv(n, 256), 256, 0, (const bf16*)x, (bf16*)y, rows, C, H, W, ws, scatter);
  else SJ_LAUNCH(c, "window_permute", window_permute_kernel<float>, cdiv(n, 256), 256, 0, (const float*)x, (float*)y, rows, C, H, W, ws, scatter);
}
void center_crop(Ctx& c, const void* x, void* y, int B, int P, int C) {
  if (!c.ok() || c.dry) return;
  long long n = (long long)B * (P / 2) * (P / 2) * (C / 4);
  if (c.dtype == SJ_BF16) SJ_LAUNCH(c, "center_crop", center_crop_kernel<bf16>, cdiv(n, 256), 256, 0, (const bf16*)x, (bf16*)y, B, P, C);
  else SJ_LAUNCH(c, "center_crop", center_crop_kernel<float>, cdiv(n, 256), 256, 0, (const float*)x, (float*)y, B, P, C);
}

void patch_embed(Ctx& c, const PatchEmbedP& p) {
  if (!c.ok() || c.dry) return;
  if (p.E > 128 || p.E % 32 || p.n_in < 1 || p.n_in > 2) { c.fail(SJ_EUNSUPPORTED); return; }
  for (int i = 0; i < p.n_in; ++i)
    if (p.Cin[i] > 11 || p.S[i] % 4) { c.fail(SJ_EUNSUPPORTED); return; }
  int P = p.S[0] / 4;
  long long ntok = (long long)p.B * P * P;
  int grid = cdiv(ntok, PE_TOK);
  if (c.dtype == SJ_BF16) SJ_LAUNCH(c, "patch_embed", patch_embed_kernel<bf16>, grid, p.E, 0, p);
  else SJ_LAUNCH(c, "patch_embed", patch_embed_kernel<float>, grid, p.E, 0, p);
}

void im2col4(Ctx& c, const void* img, int itype, int B, int S, int Cin, int es, int Kpad, void* A) {
  if (!c.ok() || c.dry) return;
  if (Kpad % 8 || Kpad < 16 * Cin) { c.fail(SJ_EINVAL); return; }
  static const bool staged_off = getenv("SJ_DISABLE_IM2COL_STAGED") != nullptr;
  const int P = S / 4, epv = itype == IN_F32 ? 4 : 16;
  const size_t esz = itype == IN_F32 ? 4 : 1;
  if (!staged_off && P % IM2COL_TOK == 0 && (IM2COL_TOK * 4 * Cin * es) % epv == 0 && ((size_t)S * Cin * es * esz) % 16 == 0 &&
      (reinterpret_cast<uintptr_t>(img) & 15) == 0) {
    const size_t smem = (size_t)4 * IM2COL_TOK * 4 * Cin * 2;
    const int grid = B * P * (P / IM2COL_TOK);
    if (itype == IN_F32) SJ_LAUNCH(c, "im2col4_staged", im2col4_staged_kernel<IN_F32>, grid, 256, smem, img, B, S, Cin, es, Kpad, (bf16*)A);
    else if (itype == IN_U8) SJ_LAUNCH(c, "im2col4_staged", im2col4_staged_kernel<IN_U8>, grid, 256, smem, img, B, S, Cin, es, Kpad, (bf16*)A);
    else SJ_LAUNCH(c, "im2col4_staged", im2col4_staged_kernel<IN_I8_DIV256>, grid, 256, smem, img, B, S, Cin, es, Kpad, (bf16*)A);
    return;
  }
  const long long n = (long long)B * (S / 4) * (S / 4) * (Kpad / 8);
  SJ_LAUNCH(c, "im2col4", im2col4_kernel, cdiv(n, 256), 256, 0, img, itype, B, S, Cin, es, Kpad, (bf16*)A);
}
void pe_combine(Ctx& c, const void* c0, const void* c1, int B, int P, int pad1, const SjNorm& n0, const SjNorm& n1,
                const SjNorm& nf, void* y, float* st_mean, float* st_rstd) {
  if (!c.ok() || c.dry) return;
  static const bool fast_off = getenv("SJ_DISABLE_NORM_FAST") != nullptr;
  if (!fast_off && pe_combine_fast(c, c0, c1, B, P, pad1, n0, n1, nf, y, st_mean, st_rstd)) return;
  const long long ntok = (long long)B * P * P;
  SJ_LAUNCH(c, "pe_combine", pe_combine_kernel, cdiv(ntok, 8), 256, 0, (const bf16*)c0, (const bf16*)c1, B, P, pad1, n0, n1,
            nf, (bf16*)y, st_mean, st_rstd);
}

void fg_offset(Ctx& c, const void* q, int ldq, const SjFgmsaW* w, int B, float* off, float* pos) {
  if (!c.ok() || c.dry) return;
  static const bool mma_off = getenv("SJ_DISABLE_FG_OFFSET_MMA") != nullptr;
  if (c.dtype == SJ_BF16 && !mma_off && fg_offset_mma(c, q, ldq, w, B, off, pos)) return;
  if (c.dtype == SJ_BF16) SJ_LAUNCH(c, "fg_offset", fg_offset_kernel<bf16>, B * 32, 384, 0, (const bf16*)q, ldq, *w, off, pos);
  else SJ_LAUNCH(c, "fg_offset", fg_offset_kernel<float>, B * 32, 384, 0, (const float*)q, ldq, *w, off, pos);
}
void build_query(Ctx& c, const void* q2, const float* off, const SjFgmsaW* w, int B, int fg, void* query) {
  if (!c.ok() || c.dry) return;
  long long n = (long long)B * 8 * 256 * 96;
  const float* w2 = fg ? w->offproj2_w : nullptr;
  const float* b2 = fg ? w->offproj2_b : nullptr;
  if (fg && (!w2 || !b2 || !off)) { c.fail(SJ_EINVAL); return; }
  if (c.dtype == SJ_BF16) SJ_LAUNCH(c, "build_query", build_query_bf16_kernel, cdiv(n / 16, 256), 256, 0, (const bf16*)q2, off, w2, b2, B, fg, (bf16*)query);
  else SJ_LAUNCH(c, "build_query", build_query_kernel<float>, cdiv(n, 256), 256, 0, (const float*)q2, off, w2, b2, B, fg, (float*)query);
}
void fg_flow_hidden(Ctx& c, const float* off, const SjFgmsaW* w, int B, void* out) {
  build_query(c, nullptr, off, w, B, 1, out);
}

void traj_node(Ctx& c, const float* obs, const float* occ, const SjTrajW* w, int B, void* node, int* stepmask,
               int* cmask, float* vec) {
  if (!c.ok() || c.dry) return;
  if (c.dtype == SJ_BF16) SJ_LAUNCH(c, "traj_node", traj_node_kernel<bf16>, B * 64, 64, 0, obs, occ, *w, B, (bf16*)node, stepmask, cmask, vec);
  else SJ_LAUNCH(c, "traj_node", traj_node_kernel<float>, B * 64, 64, 0, obs, occ, *w, B, (float*)node, stepmask, cmask, vec);
}
void traj_pool_concat(Ctx& c, const void* proj, const float* vec, int n_actors, void* cat) {
  if (!c.ok() || c.dry) return;
  if (c.dtype == SJ_BF16) SJ_LAUNCH(c, "traj_pool_concat", traj_pool_concat_kernel<bf16>, n_actors, 384, 0, (const bf16*)proj, vec, (bf16*)cat);
  else SJ_LAUNCH(c, "traj_pool_concat", traj_pool_concat_kernel<float>, n_actors, 384, 0, (const float*)proj, vec, (float*)cat);
}
void traj_prep(Ctx& c, const void* E, const int* cmask, const float* seg_w, int n_actors, void* A, void* Q) {
  if (!c.ok() || c.dry) return;
  if (c.dtype == SJ_BF16) SJ_LAUNCH(c, "traj_prep", traj_prep_kernel<bf16>, n_actors, 384, 0, (const bf16*)E, cmask, seg_w, (bf16*)A, (bf16*)Q);
  else SJ_LAUNCH(c, "traj_prep", traj_prep_kernel<float>, n_actors, 384, 0, (const float*)E, cmask, seg_w, (float*)A, (float*)Q);
}
void traj_final(Ctx& c, const void* E, const void* F2, const SjTrajW* w, int n_actors, void* key) {
  if (!c.ok() || c.dry) return;
  if (c.dtype == SJ_BF16) SJ_LAUNCH(c, "traj_final", traj_final_kernel<bf16>, cdiv(n_actors, 8), 256, 0, (const bf16*)E, (const bf16*)F2, *w, n_actors, (bf16*)key);
  else SJ_LAUNCH(c, "traj_final", traj_final_kernel<float>, cdiv(n_actors, 8), 256, 0, (const float*)E, (const float*)F2, *w, n_actors, (float*)key);
}

void out_conv(Ctx& c, const void* x_occ, const void* x_flow, const float* w, const float* b, int B, int out_layout,
              void* out) {
  if (!c.ok() || c.dry) return;
  dim3 grid(256 / OC_TX, 256 / OC_TY, B);
  const size_t smem = (size_t)(OC_TY + 2) * (OC_TX + 2) * OC_PSTRIDE * c.esize() + 864 * sizeof(float2);
  if (c.dtype == SJ_BF16) {
    if (!SJ_SMEM_LIMIT_OK((out_conv_kernel<bf16>), (int)smem)) { c.fail(SJ_ECUDA); return; }
    SJ_LAUNCH(c, "out_conv", out_conv_kernel<bf16>, grid, 128, smem, (const bf16*)x_occ, (const bf16*)x_flow, w, b, out_layout, out);
  } else {
    if (!SJ_SMEM_LIMIT_OK((out_conv_kernel<float>), (int)smem)) { c.fail(SJ_ECUDA); return; }
    SJ_LAUNCH(c, "out_conv", out_conv_kernel<float>, grid, 128, smem, (const float*)x_occ, (const float*)x_flow, w, b, out_layout, out);
  }
}

}  // namespace sj
