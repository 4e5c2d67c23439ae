// K1 -- fused window multi-head self-attention on the tensor cores (bf16, C = 96, 3 heads x 32, window 8):
//   norm1 -> cyclic shift -> window partition -> QKV (+bias) -> q*scale -> q k^T + relative-position bias
//   (+ shifted-window mask) -> softmax -> . v -> head merge -> proj (+bias) -> window reverse -> shift back
//   -> + shortcut                                   (modules.py:225-258 with WindowAttention.call :103-134)
// as ONE kernel: the only HBM traffic is the token tile in and out (plus per-token LayerNorm statistics).
//
// One tile = two 8x8 windows = 128 tokens = the M of every tcgen05.mma.
//   * TMA: each window arrives as four 4x4-token quadrant boxes of a 5-D view {C, 4, W/4, 4, B*H/4} of the
//     NHWC token grid; the cyclic shift (4 = half a window) only changes the quadrant coordinates modulo the
//     grid, so wrap-around needs no special case.  Tokens sit in shared memory in (window, quadrant, y, x)
//     order; attention is permutation-equivariant inside a window, and bias / mask / output addresses are
//     computed from each row's true coordinates.
//   * LayerNorm is folded (tc_gemm.cu): raw tokens feed the QKV MMA, the epilogue applies
//     rstd*(acc - mean*colsum) + bias'.
//   * per head: QKV_h MMA (N=96) -> rows to bf16 q|k|v tiles in smem -> S = q k^T (M=128, N=128: both windows'
//     keys, the cross-window half is never read) -> 128 row-threads: bias, mask, exp2 softmax -> un-normalised
//     P (bf16) -> O_h = P . v_h (v_h is an MN-major B operand, no transpose) -> O_h/sum to smem ->
//     proj accumulates O_h . Wproj[h] into a persistent TMEM accumulator.
//   * TMEM: region A (128 columns) is reused for QKV_h -> S -> O_h, region B (96 columns) holds proj.
//   * the weights of ONE head (norm1-folded Wqkv_h 18 KB, Wproj_h 6 KB) are streamed per head through two
//     single-slot mbarrier-guarded buffers (freed right after the MMA that reads them, so the next head's weights
//     arrive under the softmax); everything is staged in 32-channel SWIZZLE_64B chunks (K = 96 = 3 chunks, no
//     padding).  110 KB of shared memory and 256 TMEM columns per CTA: TWO CTAs are resident per SM, so while one
//     tile waits on an MMA / TMEM round trip the other tile's row threads run.
// Warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warps 2..5 = the 128 row threads.
#include <cstdio>

#include "kernels.h"
#include "tc_common.cuh"

namespace sj {
namespace {

using namespace tc;

constexpr int NTHREADS = 192;
constexpr int X_CHUNK = 128 * 64;         // 8 KB: 128 rows x 32 channels
constexpr int WQ_CHUNK = 96 * 64;         // 6 KB: (q|k|v of one head) x 32 channels
constexpr int QKV_TILE = 128 * 64;        // 8 KB: 128 rows x 32 dims
constexpr int P_CHUNK = 128 * 128;        // 16 KB: 128 rows x 64 keys
// shared-memory plan for C channels (NH = C / 32 heads, K of the QKV GEMM in NC = C / 32 chunks)
template <int C>
struct Plan {
  static constexpr int NH = C / 32, NC = C / 32;
  static constexpr int WP_TILE = C * 64;  // proj rows x 32 channels of one head
  static constexpr int OFF_X = 0;
  static constexpr int OFF_WQ = OFF_X + NC * X_CHUNK;
  static constexpr int OFF_WP = OFF_WQ + NC * WQ_CHUNK;
  static constexpr int OFF_Q = OFF_WP + WP_TILE;
  static constexpr int OFF_K = OFF_Q + QKV_TILE;
  static constexpr int OFF_V = OFF_K + QKV_TILE;
  static constexpr int OFF_P = OFF_V + QKV_TILE;       // 1024-aligned for the 128-byte swizzle
  static constexpr int OFF_AO = OFF_Q;                 // O_h aliases q_h (dead once S_h has been computed)
  static constexpr int OFF_TBL = OFF_P + 2 * P_CHUNK;  // float [NH][228]
  static constexpr int OFF_LN = OFF_TBL + NH * 228 * 4;  // float colsum[3C], biasf[3C]
  static constexpr int OFF_RID = OFF_LN + 2 * 3 * C * 4;  // int [128]
  static constexpr int OFF_BAR = (OFF_RID + 128 * 4 + 63) & ~63;
  static constexpr int SMEM_BYTES = OFF_BAR + 256;
  static constexpr int TMEM_COLS = 128 + C <= 256 ? 256 : 512;  // region A (128) + the proj accumulator (C)
  static constexpr int CTAS_PER_SM = C == 96 ? 2 : 1;
  static_assert(OFF_P % 1024 == 0, "P tile alignment");
  static_assert(CTAS_PER_SM * (SMEM_BYTES + 1024 + 1024) <= 228 * 1024, "shared memory per SM");
  static_assert(128 + C <= 512 && C % 32 == 0 && C <= 256, "one proj MMA (N = C <= 256), TMEM budget");
};
constexpr float QSCALE = 0.17677669529663687f * 1.4426950408889634f;  // head_dim^-0.5 * log2(e)
constexpr float LOG2E = 1.4426950408889634f;

enum { B_XFULL = 0, B_XEMPTY, B_WQFULL, B_WQEMPTY, B_WPFULL, B_WPEMPTY, B_QKV, B_QKR, B_S, B_PR, B_O, B_AOR, B_PROJ, B_EPI, B_COUNT };

struct WmsaP {
  int B, H, W, shift, num_tiles, nW, wpr;  // wpr = windows per row (W/8)
  int ctas;  // CTAs serving this group (a launch may carry two independent groups: different x / weights)
  const bf16* x;
  bf16* out;
  const float* mean;
  const float* rstd;
  const float* colsum;
  const float* biasf;
  const float* table;  // [225, NH]
  const float* bproj;  // [C]
  float* mean2;        // optional: norm2 statistics of the output rows
  float* rstd2;
};

struct WmsaPPair { WmsaP g[2]; };
struct PairMaps { CUtensorMap m[2][3]; };

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint4 pack8(const float* f) {
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(f[4], f[5]), d = __floats2bfloat162_rn(f[6], f[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
  return u;
}
// 32 bf16 (64 B) of row r into a [128 x 64 B] SWIZZLE_64B tile
__device__ __forceinline__ void store_row64(uint8_t* tile, int r, const float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(tile + r * 64 + ((j ^ ((r >> 1) & 3)) << 4)) = pack8(v + 8 * j);
}

template <int C>
__global__ void __launch_bounds__(NTHREADS, Plan<C>::CTAS_PER_SM)
tc_wmsa_kernel(const __grid_constant__ PairMaps maps, const WmsaPPair pp) {
  using PL = Plan<C>;
  constexpr int NH = PL::NH, NC = PL::NC, WP_TILE = PL::WP_TILE;
  constexpr int OFF_X = PL::OFF_X, OFF_WQ = PL::OFF_WQ, OFF_WP = PL::OFF_WP, OFF_Q = PL::OFF_Q, OFF_K = PL::OFF_K;
  constexpr int OFF_V = PL::OFF_V, OFF_P = PL::OFF_P, OFF_AO = PL::OFF_AO, OFF_TBL = PL::OFF_TBL, OFF_LN = PL::OFF_LN;
  constexpr int OFF_RID = PL::OFF_RID, OFF_BAR = PL::OFF_BAR;
  // group of this CTA: CTAs [0, pp.g[0].ctas) run group 0, the rest group 1 (same shapes, other tensors: the flow / raster
  // branches of the encoder in lock step).  Both parameter sets sit in one array so that the choice is a constant-bank
  // offset, not a select per field.
  const int grp = blockIdx.x >= (unsigned)pp.g[0].ctas ? 1 : 0;
  const WmsaP& p = pp.g[grp];
  const CUtensorMap* mapX = &maps.m[grp][0];
  const CUtensorMap* mapWq = &maps.m[grp][1];
  const CUtensorMap* mapWp = &maps.m[grp][2];
  const int cta = (int)blockIdx.x - (grp ? pp.g[0].ctas : 0), nctas = p.ctas;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + B_COUNT);
  float* tbl = reinterpret_cast<float*>(smem + OFF_TBL);
  float* lnc = reinterpret_cast<float*>(smem + OFF_LN);
  int* rid_s = reinterpret_cast<int*>(smem + OFF_RID);

  const int warp = uniform_warp_idx(), lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(mapX);
    prefetch_tmap(mapWq);
    prefetch_tmap(mapWp);
    const int counts[B_COUNT] = {1, 1, 1, 1, 1, 1, 1, 4, 1, 4, 1, 4, 1, 4};
    for (int i = 0; i < B_COUNT; ++i) mbar_init(&bar[i], counts[i]);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, PL::TMEM_COLS);
  // constants: bias table (pre-multiplied by log2 e), LayerNorm-fold vectors; zero the P tile once (the
  // cross-window halves are never written again)
  for (int i = threadIdx.x; i < NH * 225; i += NTHREADS) tbl[(i % NH) * 228 + i / NH] = p.table[i] * LOG2E;
  for (int i = threadIdx.x; i < 3 * C; i += NTHREADS) {
    lnc[i] = p.colsum[i];
    lnc[3 * C + i] = p.biasf[i];
  }
  for (int i = threadIdx.x; i < 2 * P_CHUNK / 16; i += NTHREADS)
    reinterpret_cast<uint4*>(smem + OFF_P)[i] = make_uint4(0, 0, 0, 0);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_a = *tmem_slot, tmem_proj = tmem_a + 128;
  pdl_wait();  // everything above (barriers, TMEM, folded constants) overlaps the previous kernel's tail
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      const int s4 = p.shift >> 2, wq = p.W >> 2, hq = p.H >> 2;
      uint32_t it = 0, hc = 0;
      for (int tile = cta; tile < p.num_tiles; tile += nctas, ++it) {
        mbar_wait(&bar[B_XEMPTY], (it & 1) ^ 1);
        mbar_expect_tx(&bar[B_XFULL], NC * X_CHUNK);
        for (int wt = 0; wt < 2; ++wt) {
          const int gw = 2 * tile + wt, b = gw / p.nW, wl = gw % p.nW, wi = wl / p.wpr, wj = wl % p.wpr;
          for (int quad = 0; quad < 4; ++quad) {
            const int yq = (2 * wi + (quad >> 1) + s4) % hq, xq = (2 * wj + (quad & 1) + s4) % wq;
            for (int c = 0; c < NC; ++c)
              tma_load_5d(smem + OFF_X + c * X_CHUNK + (wt * 64 + quad * 16) * 64, mapX, &bar[B_XFULL], c * 32, 0, xq,
                          0, b * hq + yq);
          }
        }
        // weights of each head: the slots are released by the MMA warp as soon as the MMAs reading them have completed
        for (int h = 0; h < NH; ++h, ++hc) {
          mbar_wait(&bar[B_WQEMPTY], (hc & 1) ^ 1);
          mbar_expect_tx(&bar[B_WQFULL], NC * WQ_CHUNK);
          for (int c = 0; c < NC; ++c)
            for (int part = 0; part < 3; ++part)
              tma_load_2d(smem + OFF_WQ + c * WQ_CHUNK + part * 32 * 64, mapWq, &bar[B_WQFULL], c * 32, part * C + h * 32);
          mbar_wait(&bar[B_WPEMPTY], (hc & 1) ^ 1);
          mbar_expect_tx(&bar[B_WPFULL], WP_TILE);
          tma_load_2d(smem + OFF_WP, mapWp, &bar[B_WPFULL], h * 32, 0);
        }
      }
    }
  } else if (warp == 1) {
    // the whole warp runs the loop (uniform control flow); one elected lane issues each batch of MMAs
    constexpr uint32_t HI128 = desc_hi(128, 1024), HI64 = desc_hi(64, 512);
    const uint32_t id_qkv = make_idesc_bf16(128, 96), id_s = make_idesc_bf16(128, 128), id_proj = make_idesc_bf16(128, C);
    const uint32_t id_pv = make_idesc_bf16(128, 32) | (1u << 16);  // B (= v_h) is MN-major
    const uint32_t x_lo = desc_lo(smem_u32(smem + OFF_X)), wq_lo = desc_lo(smem_u32(smem + OFF_WQ));
    const uint32_t wp_lo = desc_lo(smem_u32(smem + OFF_WP)), q_lo = desc_lo(smem_u32(smem + OFF_Q));
    const uint32_t k_lo = desc_lo(smem_u32(smem + OFF_K)), v_lo = desc_lo(smem_u32(smem + OFF_V));
    const uint32_t p_lo = desc_lo(smem_u32(smem + OFF_P)), ao_lo = desc_lo(smem_u32(smem + OFF_AO));
    uint32_t it = 0, hc = 0;  // tiles / heads processed by this CTA
    for (int tile = cta; tile < p.num_tiles; tile += nctas, ++it) {
      mbar_wait(&bar[B_XFULL], it & 1);
      for (int h = 0; h < NH; ++h, ++hc) {
        const uint32_t hp = hc & 1;
        // region A is free: the row threads finished reading O of the previous head (B_AOR waited below)
        mbar_wait(&bar[B_WQFULL], hp);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int s = 0; s < 2 * NC; ++s) {  // K = C: 32-channel chunks x two K steps
            const uint32_t c = s >> 1, k = s & 1;
            umma_bf16_w(tmem_a, x_lo + (c * X_CHUNK >> 4) + 2 * k, HI64, wq_lo + (c * WQ_CHUNK >> 4) + 2 * k, HI64, id_qkv,
                        s != 0);
          }
          umma_commit(&bar[B_QKV]);
          umma_commit(&bar[B_WQEMPTY]);
          if (h == NH - 1) umma_commit(&bar[B_XEMPTY]);
        }
        __syncwarp();
        mbar_wait(&bar[B_QKR], hp);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_bf16_w(tmem_a, q_lo + 2 * k, HI64, k_lo + 2 * k, HI64, id_s, k != 0);
          umma_commit(&bar[B_S]);
        }
        __syncwarp();
        mbar_wait(&bar[B_PR], hp);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int s = 0; s < 8; ++s)  // 128 keys: A = P chunk s/4 (K-major), B = v rows 16s.. (MN-major, 64 B rows)
            umma_bf16_w(tmem_a, p_lo + ((s >> 2) * (P_CHUNK >> 4)) + 2 * (s & 3), HI128, v_lo + s * (1024 >> 4), HI64,
                        id_pv, s != 0);
          umma_commit(&bar[B_O]);
        }
        __syncwarp();
        mbar_wait(&bar[B_AOR], hp);
        if (h == 0) mbar_wait(&bar[B_EPI], (it & 1) ^ 1);  // previous tile's epilogue has drained the proj accumulator
        mbar_wait(&bar[B_WPFULL], hp);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma_bf16_w(tmem_proj, ao_lo + 2 * k, HI64, wp_lo + 2 * k, HI64, id_proj, (h | k) != 0);
          umma_commit(&bar[B_WPEMPTY]);
          if (h == NH - 1) umma_commit(&bar[B_PROJ]);
        }
        __syncwarp();
      }
    }
  } else {
    const int quarter = warp % 4, r = quarter * 32 + lane;
    const int wt = r >> 6, rho = r & 63, quad = rho >> 4;
    const int rw = (quad >> 1) * 4 + ((rho >> 2) & 3), cw = (quad & 1) * 4 + (rho & 3);
    const int base_q = (rw + 7) * 15 + cw + 7;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    uint32_t it = 0, hc = 0;
    for (int tile = cta; tile < p.num_tiles; tile += nctas, ++it) {
      const int gw = 2 * tile + wt, b = gw / p.nW, wl = gw % p.nW, wi = wl / p.wpr, wj = wl % p.wpr;
      const int ys = 8 * wi + rw, xs = 8 * wj + cw;
      const long long tok = ((long long)b * p.H + (ys + p.shift) % p.H) * p.W + (xs + p.shift) % p.W;
      const float mean = p.mean[tok], rstd = p.rstd[tok];
      int myrid = 0;
      if (p.shift > 0) {
        myrid = shift_region_id(p.H, p.W, 8, p.shift, ys, xs);
        rid_s[r] = myrid;
        asm volatile("bar.sync 1, 128;" ::: "memory");  // the 128 row threads only
      }
      float inv_sum[NH];
#pragma unroll 1
      for (int h = 0; h < NH; ++h, ++hc) {
        const uint32_t hp = hc & 1;
        // ---- q | k | v of this head: TMEM -> LayerNorm fold -> bf16 operand tiles ----
        mbar_wait(&bar[B_QKV], hp);
        tc_fence_after();
#pragma unroll
        for (int part = 0; part < 3; ++part) {
          float v[32];
          tmem_ld32(tmem_a + lane_addr + part * 32, v);
          const float* cs = lnc + part * C + h * 32;
          const float* bf = lnc + 3 * C + part * C + h * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float t = rstd * (v[i] - mean * cs[i]) + bf[i];
            v[i] = part == 0 ? t * QSCALE : t;
          }
          store_row64(smem + (part == 0 ? OFF_Q : (part == 1 ? OFF_K : OFF_V)), r, v);
        }
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar[B_QKR]);
        // ---- softmax over this row's own window (64 keys) ----
        mbar_wait(&bar[B_S], hp);
        tc_fence_after();
        float s[64];
        {
          float a[32], c2[32];
          tmem_ld32(tmem_a + lane_addr + wt * 64, a);
          tmem_ld32(tmem_a + lane_addr + wt * 64 + 32, c2);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            s[i] = a[i];
            s[32 + i] = c2[i];
          }
        }
        const float* tb = tbl + h * 228 + base_q;
        float mx = -INFINITY;
#pragma unroll
        for (int m = 0; m < 64; ++m) {
          // key m sits at window coordinates (rk, ck) in the quadrant order
          const int rk = ((m >> 4) >> 1) * 4 + ((m >> 2) & 3), ck = ((m >> 4) & 1) * 4 + (m & 3);
          float t = s[m] + tb[-(rk * 15 + ck)];
          if (p.shift > 0) t += (rid_s[wt * 64 + m] != myrid) ? -100.0f * LOG2E : 0.0f;
          s[m] = t;
          mx = fmaxf(mx, t);
        }
        float sum = 0.f;
#pragma unroll
        for (int m = 0; m < 64; ++m) {
          s[m] = ex2(s[m] - mx);
          sum += s[m];
        }
        inv_sum[h] = 1.0f / sum;
        {
          uint8_t* prow = smem + OFF_P + wt * P_CHUNK + r * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(prow + ((j ^ (r & 7)) << 4)) = pack8(s + 8 * j);
        }
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar[B_PR]);
        // ---- O_h / sum -> bf16 A tile of the projection ----
        mbar_wait(&bar[B_O], hp);
        tc_fence_after();
        {
          float o[32];
          tmem_ld32(tmem_a + lane_addr, o);
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] *= inv_sum[h];
          store_row64(smem + OFF_AO, r, o);
        }
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar[B_AOR]);
      }
      // ---- proj + bias + shortcut -> out (window reverse + shift back = this row's own token) ----
      mbar_wait(&bar[B_PROJ], it & 1);
      tc_fence_after();
      const bf16* xin = p.x + tok * C;
      bf16* dst = p.out + tok * C;
      float st_s = 0.f, st_q = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < C; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_proj + lane_addr + c0, v);
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          float xr[8];
          ld8_bf16(xin + c0 + i, xr);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[i + j] += p.bproj[c0 + i + j] + xr[j];
            const float rr = __bfloat162float(__float2bfloat16_rn(v[i + j]));  // statistics of the stored values
            st_s += rr;
            st_q = fmaf(rr, rr, st_q);
          }
          st8_bf16(dst + c0 + i, v + i);
        }
      }
      if (p.mean2) {
        const float mu = st_s * (1.0f / C);
        p.mean2[tok] = mu;
        p.rstd2[tok] = rsqrtf(fmaxf(st_q * (1.0f / C) - mu * mu, 0.f) + 1e-5f);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[B_EPI]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_a, PL::TMEM_COLS);
  }
}

}  // namespace

bool tc_wmsa_supported(int B, int H, int W, int Cc, int heads, int ws, int shift) {
  if ((Cc != 96 && Cc != 192) || heads * 32 != Cc || ws != 8 || (shift != 0 && shift != 4)) return false;
  if (H % 8 || W % 8 || H < 8 || W < 8) return false;
  return ((long long)B * (H / 8) * (W / 8)) % 2 == 0;
}

// x, out: bf16 [B, H*W, 96]; mean/rstd: fp32 [B*H*W] (norm1 statistics of x); w: norm1-folded qkv (tensor-core copy).
// groups = 2: two independent problems of the same geometry in ONE launch (half of the CTAs each)
template <int C>
static void tc_wmsa_launch_c(Ctx& c, int groups, const void* const* x, void* const* out, const float* const* mean,
                           const float* const* rstd, const SjSwinBlockW* const* w, int B, int H, int W, int shift,
                           float* const* mean2, float* const* rstd2) {
  if (!c.ok() || c.dry) return;
  PairMaps pm;
  CUtensorMap (&maps)[2][3] = pm.m;
  WmsaPPair ppair = {};
  WmsaP (&pp)[2] = ppair.g;
  const int tiles = B * (H / 8) * (W / 8) / 2;
  const int cap = Plan<C>::CTAS_PER_SM * num_sms() / groups;
  const int per = tiles < cap ? tiles : cap;
  for (int g = 0; g < groups; ++g) {
    const SjSwinBlockW& wg = *w[g];
    if (!wg.qkv_ln.w_tc || !wg.qkv_ln.tc_colsum || !wg.qkv_ln.tc_bias || !wg.proj.w_tc || !wg.proj.b || !wg.rpb_table) {
      c.fail(SJ_EINVAL);
      return;
    }
    uint64_t dx[5] = {(uint64_t)C, 4, (uint64_t)W / 4, 4, (uint64_t)B * H / 4};
    uint64_t sx[4] = {(uint64_t)C * 2, (uint64_t)4 * C * 2, (uint64_t)W * C * 2, (uint64_t)4 * W * C * 2};
    uint32_t bx[5] = {32, 4, 1, 4, 1};
    uint64_t dq[2] = {(uint64_t)C, (uint64_t)3 * C};
    uint64_t sq[1] = {(uint64_t)C * 2};
    uint32_t bq[2] = {32, 32};
    uint64_t dp[2] = {(uint64_t)C, (uint64_t)C};
    uint32_t bp[2] = {32, (uint32_t)C};
    if (!encode_tmap(&maps[g][0], x[g], 5, dx, sx, bx, 64) || !encode_tmap(&maps[g][1], wg.qkv_ln.w_tc, 2, dq, sq, bq, 64) ||
        !encode_tmap(&maps[g][2], wg.proj.w_tc, 2, dp, sq, bp, 64)) {
      snprintf(tls().cuda_err, sizeof(tls().cuda_err), "cuTensorMapEncodeTiled failed (tc_wmsa)");
      c.fail(SJ_ECUDA);
      return;
    }
    WmsaP& p = pp[g];
    p.B = B; p.H = H; p.W = W; p.shift = shift;
    p.wpr = W / 8; p.nW = (H / 8) * (W / 8);
    p.num_tiles = tiles; p.ctas = per;
    p.x = (const bf16*)x[g]; p.out = (bf16*)out[g]; p.mean = mean[g]; p.rstd = rstd[g];
    p.mean2 = mean2 ? mean2[g] : nullptr; p.rstd2 = rstd2 ? rstd2[g] : nullptr;
    p.colsum = wg.qkv_ln.tc_colsum; p.biasf = wg.qkv_ln.tc_bias; p.table = wg.rpb_table; p.bproj = wg.proj.b;
  }
  if (groups == 1) {
    pp[1] = pp[0];
    for (int i = 0; i < 3; ++i) maps[1][i] = maps[0][i];
  }
  const size_t smem = 1024 + Plan<C>::SMEM_BYTES;
  if (!SJ_SMEM_LIMIT_OK((tc_wmsa_kernel<C>), 227 * 1024)) {
    c.fail(SJ_ECUDA);
    return;
  }
  SJ_LAUNCH(c, "tc_wmsa", tc_wmsa_kernel<C>, groups * per, NTHREADS, smem, pm, ppair);
}

static void tc_wmsa_launch(Ctx& c, int Cc, int groups, const void* const* x, void* const* out, const float* const* mean,
                           const float* const* rstd, const SjSwinBlockW* const* w, int B, int H, int W, int shift,
                           float* const* mean2, float* const* rstd2) {
  if (Cc == 96) tc_wmsa_launch_c<96>(c, groups, x, out, mean, rstd, w, B, H, W, shift, mean2, rstd2);
  else if (Cc == 192) tc_wmsa_launch_c<192>(c, groups, x, out, mean, rstd, w, B, H, W, shift, mean2, rstd2);
  else c.fail(SJ_EUNSUPPORTED);
}

void tc_wmsa(Ctx& c, const void* x, void* out, const float* mean, const float* rstd, const SjSwinBlockW& w, int B,
             int H, int W, int Cc, int shift, float* mean2, float* rstd2) {
  const SjSwinBlockW* wp = &w;
  tc_wmsa_launch(c, Cc, 1, &x, &out, &mean, &rstd, &wp, B, H, W, shift, &mean2, &rstd2);
}

void tc_wmsa_pair(Ctx& c, const void* const x[2], void* const out[2], const float* const mean[2], const float* const rstd[2],
                  const SjSwinBlockW* const w[2], int B, int H, int W, int shift, float* const mean2[2],
                  float* const rstd2[2]) {
  tc_wmsa_launch(c, 96, 2, x, out, mean, rstd, w, B, H, W, shift, mean2, rstd2);
}

}  // namespace sj
