// fp32-accumulate SIMT GEMM / implicit-GEMM used by the fp32 parity path (config 2) and as the
// shape-generic companion of the tcgen05 kernels.  C = [R +] act(LN?(gatherA) . W + bias).
//
// Tile 128 x BN x 16, 256 threads, 8 x (BN/16) register micro-tile, smem double buffering.
// A-operand modes: plain rows (with optional gather map / LayerNorm-on-load), implicit 3x3 SAME
// conv over NHWC with optional fused nearest x2 upsample (modules.py:746-749), and the
// PatchMerging 2x2 gather (modules.py:282-287).
#include "kernels.h"

namespace sj {
namespace {

constexpr int BM = 128, BK = 16, NT = 256;

__device__ __forceinline__ long long map_row(const RowMap& r, int m, int g) {
  long long row = m;
  if (r.inner > 0) row = (long long)(m / r.inner) * r.outer + (m % r.inner);
  row += (long long)g * r.gstride;
  if (r.map != nullptr) row = (row / r.map_len) * r.map_len + r.map[row % r.map_len];
  return row;
}

template <typename T, int AMODE, int BN>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(const GemmP p) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  constexpr int TN = BN / 16;  // columns per thread (8 or 4)
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int t = threadIdx.x;
  const int g = blockIdx.z;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const T* __restrict__ A = reinterpret_cast<const T*>(p.A);
  const float* __restrict__ Wg = p.W + (long long)g * p.w_gstride;

  // ---- A loader state: this thread always loads row (t % 128), k-slice (t / 128) * 8 ----
  const int arow = t % BM, akk = (t / BM) * 8;
  const int am = m0 + arow;
  const bool avalid = am < p.M;
  const T* aptr = A;
  float mean = 0.f, rstd = 1.f;
  int ci = 0, cy = 0, cx = 0;  // conv: image, y, x; merge: batch, i, j
  if (avalid) {
    if (AMODE == A_PLAIN) {
      long long r = map_row(p.am, am, g);
      aptr = A + r * p.lda;
      if (p.ln_mean) {
        mean = p.ln_mean[r];
        rstd = p.ln_rstd[r];
      }
    } else if (AMODE == A_CONV3) {
      int hw = p.H * p.Wd;
      ci = am / hw;
      int rem = am - ci * hw;
      cy = rem / p.Wd;
      cx = rem - cy * p.Wd;
    } else {
      int ho = p.H / 2, wo = p.Wd / 2;
      ci = am / (ho * wo);
      int rem = am - ci * ho * wo;
      cy = rem / wo;
      cx = rem - cy * wo;
      if (p.ln_mean) {
        mean = p.ln_mean[am];
        rstd = p.ln_rstd[am];
      }
    }
  }
  const float* lng = p.ln_g ? p.ln_g + (long long)g * p.ln_gstride : nullptr;
  const float* lnb = p.ln_b ? p.ln_b + (long long)g * p.ln_gstride : nullptr;

  auto load_a = [&](int kt, float (&a)[8]) {
    const int k0 = kt * BK;
    const T* src = nullptr;
    if (avalid) {
      if (AMODE == A_PLAIN) {
        src = aptr + k0 + akk;
      } else if (AMODE == A_CONV3) {
        int tap = k0 / p.Cin;
        int c0 = k0 - tap * p.Cin + akk;
        int yy = cy + tap / 3 - 1, xx = cx + tap % 3 - 1;
        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.Wd) {
          int hi = p.H >> p.up, wi = p.Wd >> p.up;
          src = A + (((long long)ci * hi + (yy >> p.up)) * wi + (xx >> p.up)) * p.Cin + c0;
        }
      } else {
        int q = k0 / p.Cin;
        int c0 = k0 - q * p.Cin + akk;
        src = A + (((long long)ci * p.H + 2 * cy + (q & 1)) * p.Wd + 2 * cx + (q >> 1)) * p.Cin + c0;
      }
    }
    if (src) {
      float4 v0 = ld4<T>(src), v1 = ld4<T>(src + 4);
      a[0] = v0.x; a[1] = v0.y; a[2] = v0.z; a[3] = v0.w;
      a[4] = v1.x; a[5] = v1.y; a[6] = v1.z; a[7] = v1.w;
      if (AMODE != A_CONV3 && lng) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = (a[i] - mean) * rstd * lng[k0 + akk + i] + lnb[k0 + akk + i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = 0.f;
    }
  };

  // ---- B loader: BK x BN floats = BK*BN/4 float4, NT threads ----
  constexpr int BV = BK * BN / 4 / NT;  // float4 per thread (2 or 1)
  auto load_b = [&](int kt, float4 (&b)[BV]) {
#pragma unroll
    for (int i = 0; i < BV; ++i) {
      int idx = t + i * NT;
      int k = idx / (BN / 4), n4 = idx % (BN / 4);
      int n = n0 + n4 * 4;
      if (n < p.N) b[i] = *reinterpret_cast<const float4*>(Wg + (long long)(kt * BK + k) * p.ldw + n);
      else b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tiles = [&](int buf, const float (&a)[8], const float4 (&b)[BV]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[buf][akk + i][arow] = a[i];
#pragma unroll
    for (int i = 0; i < BV; ++i) {
      int idx = t + i * NT;
      int k = idx / (BN / 4), n4 = idx % (BN / 4);
      *reinterpret_cast<float4*>(&Bs[buf][k][n4 * 4]) = b[i];
    }
  };

  const int tx = t % 16, ty = t / 16;
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int KT = p.K / BK;
  float ra[8];
  float4 rb[BV];
  load_a(0, ra);
  load_b(0, rb);
  store_tiles(0, ra, rb);
  __syncthreads();
  for (int kt = 0; kt < KT; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < KT) {
      load_a(kt + 1, ra);
      load_b(kt + 1, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[TN];
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      if (TN == 8) {
        float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][(BN / 2) % BN + tx * 4]);
        b[TN - 4] = b1.x; b[TN - 3] = b1.y; b[TN - 2] = b1.z; b[TN - 1] = b1.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < KT) {
      store_tiles(buf ^ 1, ra, rb);
      __syncthreads();
    }
  }

  // ---- epilogue ----
  T* __restrict__ C = reinterpret_cast<T*>(p.C);
  const T* __restrict__ R = reinterpret_cast<const T*>(p.R);
  const float* bias = p.bias ? p.bias + (long long)g * p.bias_gstride : nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
    long long crow = map_row(p.cm, m, g);
#pragma unroll
    for (int h = 0; h < TN / 4; ++h) {
      int n = n0 + h * (BN / 2) + tx * 4;
      if (n >= p.N) continue;
      float4 v = make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
      if (bias) {
        float4 bb = *reinterpret_cast<const float4*>(bias + n);
        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
      }
      v.x = apply_act(v.x, p.act); v.y = apply_act(v.y, p.act);
      v.z = apply_act(v.z, p.act); v.w = apply_act(v.w, p.act);
      if (R) {
        float4 r = ld4<T>(R + crow * p.ldr + n);
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      }
      st4<T>(C + crow * p.ldc + n, v);
    }
  }
}

template <typename T, int AMODE>
void launch_bn(Ctx& c, const GemmP& p) {
  bool bn64 = (p.N <= 64) || ((p.N % 128) != 0 && (p.N % 128) <= 64);
  if (bn64) {
    dim3 grid(cdiv(p.M, BM), cdiv(p.N, 64), p.groups);
    SJ_LAUNCH(c, "gemm_simt", (gemm_simt_kernel<T, AMODE, 64>), grid, NT, 0, p);
  } else {
    dim3 grid(cdiv(p.M, BM), cdiv(p.N, 128), p.groups);
    SJ_LAUNCH(c, "gemm_simt", (gemm_simt_kernel<T, AMODE, 128>), grid, NT, 0, p);
  }
}

template <typename T>
void launch_mode(Ctx& c, const GemmP& p) {
  switch (p.amode) {
    case A_PLAIN: launch_bn<T, A_PLAIN>(c, p); break;
    case A_CONV3: launch_bn<T, A_CONV3>(c, p); break;
    case A_MERGE: launch_bn<T, A_MERGE>(c, p); break;
    default: c.fail(SJ_EINVAL);
  }
}

}  // namespace

void gemm_simt(Ctx& c, const GemmP& p) {
  if (!c.ok()) return;
  if (p.M <= 0 || p.N <= 0 || p.K <= 0 || (p.K % BK) != 0 || (p.N % 4) != 0 || (p.ldw % 4) != 0 ||
      (p.ldc % 4) != 0 || (p.R && (p.ldr % 4) != 0)) {
    c.fail(SJ_EINVAL);
    return;
  }
  if (p.amode == A_PLAIN && (p.lda % 4) != 0) { c.fail(SJ_EINVAL); return; }
  if (p.amode != A_PLAIN && (p.Cin % BK) != 0) { c.fail(SJ_EUNSUPPORTED); return; }
  if (c.dry) return;
  if (c.dtype == SJ_BF16) launch_mode<bf16>(c, p);
  else launch_mode<float>(c, p);
}

}  // namespace sj
