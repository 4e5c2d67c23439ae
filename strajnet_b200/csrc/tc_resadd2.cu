// Decoder skip connections at 64x64 (i = 1 of modules.py:750-757 and the flow branch's :762-765), both in ONE kernel:
//   x   = x_pre + ELU(res0     . W_eff[t]  + b )        (res_layer[1]: (8,1,1) Conv3D over the 8x-repeated skip, collapsed)
//   fx  = x     + ELU(flow_res . Wf_eff[t] + bf)        (res_f; uses x AFTER the res0 add)
// for every waypoint t.  As two grouped tc_gemm launches these cost 2 x 101 us at batch 16 for 26 GFLOP: the GEMMs are
// nothing, the op is 134 MB read + 268 MB written, and tc_gemm's thread-per-row epilogue issues 16-byte accesses at a
// 256-byte stride (32 L1 wavefronts per instruction: the LSU wavefront pipe was the busiest unit, 54 %).  Here:
//   * x_pre is read ONCE (the second launch re-read x), as TMA boxes (SWIZZLE_128B) prefetched a tile ahead, and read back
//     conflict-free by the row threads;
//   * results leave through a per-warp shared-memory transpose: every store instruction writes 8 rows x 64 contiguous bytes
//     (full sectors, 8 wavefronts instead of 32);
//   * weight-stationary: CTA c serves waypoint t = c % 8 and keeps W_eff[t], Wf_eff[t] (48 KB) resident; the eight CTAs of
//     a tile sequence walk the same (sample, pixel-block) order, so each skip tile is fetched from HBM once and hits L2 for
//     the other seven waypoints.
// Both 128 x 128 x 96 products of a tile go to two TMEM accumulators (double buffered: 512 columns).
// Warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warps 2..9 = epilogue (TMEM lane quarter = warp % 4, column half =
// (warp - 2) / 4).
#include <cstdio>

#include "kernels.h"
#include "tc_common.cuh"

namespace sj {
namespace {

using namespace tc;

constexpr int CIN = 96, COUT = 128, KC = 32, NCH = CIN / KC, BM = 128;
constexpr int NTHREADS = 320;
constexpr int A_CHUNK = BM * KC * 2;            // 8 KB: 128 rows x 32 channels (SWIZZLE_64B)
constexpr int A_STAGE = 2 * NCH * A_CHUNK;      // 48 KB: both skip tiles
constexpr int NA = 2;
constexpr int W_CHUNK = COUT * KC * 2;          // 8 KB: 128 output channels x 32 input channels
constexpr int W_BYTES = 2 * NCH * W_CHUNK;      // 48 KB: W_eff[t] and Wf_eff[t]
constexpr int R_BOX = BM * 128;                 // 16 KB: 128 rows x 64 columns (SWIZZLE_128B)
constexpr int R_STAGE = 2 * R_BOX;              // 32 KB
constexpr int NR = 2;
constexpr int STG_WARP = 32 * 64;               // 2 KB: 32 rows x 32 bf16 columns
constexpr int OFF_W = NA * A_STAGE;
constexpr int OFF_R = OFF_W + W_BYTES;
constexpr int OFF_STG = OFF_R + NR * R_STAGE;
constexpr int OFF_BAR = OFF_STG + 8 * STG_WARP;
constexpr int SMEM_BYTES = OFF_BAR + 2048;      // barriers + two bias vectors
static_assert(OFF_R % 1024 == 0, "residual tile alignment (128-byte swizzle)");
static_assert(SMEM_BYTES + 1024 <= 227 * 1024, "shared memory");

struct Res2P {
  int B, HW, tiles;      // tiles = B * HW / 128 per waypoint
  const float* bias_a;   // [COUT]
  const float* bias_b;
  bf16* dst_a;           // [B, 8, HW, COUT]
  bf16* dst_b;
};

__global__ void __launch_bounds__(NTHREADS, 1)
tc_resadd2_kernel(const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapA2,
                  const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2,
                  const __grid_constant__ CUtensorMap mapR, const Res2P p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* afull = bars;             // [NA]
  uint64_t* aempty = bars + NA;       // [NA]
  uint64_t* rfull = bars + 2 * NA;    // [NR]
  uint64_t* rempty = rfull + NR;      // [NR] (8 epilogue warps)
  uint64_t* wfull = rempty + NR;
  uint64_t* tfull = wfull + 1;        // [2]
  uint64_t* tempty = tfull + 2;       // [2] (8 epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* bias_s = reinterpret_cast<float*>(bars + 32);  // [2][COUT]

  const int warp = uniform_warp_idx(), lane = threadIdx.x % 32;
  for (int i = threadIdx.x; i < COUT; i += NTHREADS) {
    bias_s[i] = p.bias_a[i];
    bias_s[COUT + i] = p.bias_b[i];
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA1);
    prefetch_tmap(&mapA2);
    prefetch_tmap(&mapW1);
    prefetch_tmap(&mapW2);
    prefetch_tmap(&mapR);
    for (int s = 0; s < NA; ++s) {
      mbar_init(&afull[s], 1);
      mbar_init(&aempty[s], 1);
    }
    for (int s = 0; s < NR; ++s) {
      mbar_init(&rfull[s], 1);
      mbar_init(&rempty[s], 8);
    }
    mbar_init(wfull, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // CTA -> waypoint t and a strided sequence of (sample, pixel-block) tiles
  const int t = blockIdx.x & 7, first = blockIdx.x >> 3, stride = gridDim.x >> 3;
  const int blocks_per_img = p.HW / BM;
  if (warp != 0) {  // the producer lane waits after it has issued the (constant) resident weights
    pdl_wait();
    pdl_trigger();
  }

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(wfull, W_BYTES);
      for (int ch = 0; ch < NCH; ++ch) {
        tma_load_2d(smem + OFF_W + ch * W_CHUNK, &mapW1, wfull, ch * KC, t * COUT);
        tma_load_2d(smem + OFF_W + (NCH + ch) * W_CHUNK, &mapW2, wfull, ch * KC, t * COUT);
      }
      pdl_wait();
      int as = 0, rs = 0;
      uint32_t aph = 0, rph = 0;
      for (int tt = first; tt < p.tiles; tt += stride) {
        const int b = tt / blocks_per_img, m0 = (tt % blocks_per_img) * BM;
        mbar_wait(&aempty[as], aph ^ 1);
        mbar_expect_tx(&afull[as], A_STAGE);
        for (int ch = 0; ch < NCH; ++ch) {
          tma_load_2d(smem + as * A_STAGE + ch * A_CHUNK, &mapA1, &afull[as], ch * KC, b * p.HW + m0);
          tma_load_2d(smem + as * A_STAGE + (NCH + ch) * A_CHUNK, &mapA2, &afull[as], ch * KC, b * p.HW + m0);
        }
        if (++as == NA) { as = 0; aph ^= 1; }
        mbar_wait(&rempty[rs], rph ^ 1);
        mbar_expect_tx(&rfull[rs], R_STAGE);
        const int row0 = (b * 8 + t) * p.HW + m0;
        tma_load_2d(smem + OFF_R + rs * R_STAGE, &mapR, &rfull[rs], 0, row0);
        tma_load_2d(smem + OFF_R + rs * R_STAGE + R_BOX, &mapR, &rfull[rs], 64, row0);
        if (++rs == NR) { rs = 0; rph ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t HI64 = desc_hi(64, 512);
    const uint32_t idesc = make_idesc_bf16(BM, COUT);
    const uint32_t w_lo = desc_lo(smem_u32(smem + OFF_W));
    int as = 0, acc = 0;
    uint32_t aph = 0, tph = 0;
    mbar_wait(wfull, 0);
    tc_fence_after();
    for (int tt = first; tt < p.tiles; tt += stride) {
      mbar_wait(&tempty[acc], tph ^ 1);
      mbar_wait(&afull[as], aph);
      tc_fence_after();
      const uint32_t a_lo = desc_lo(smem_u32(smem + as * A_STAGE));
      const uint32_t d = tmem_base + acc * 2 * COUT;
      if (elect_one()) {
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
          for (int s = 0; s < 2 * NCH; ++s) {
            const uint32_t c = s >> 1, k = s & 1;
            umma_bf16_w(d + g * COUT, a_lo + (((g * NCH + c) * A_CHUNK) >> 4) + 2 * k, HI64,
                        w_lo + (((g * NCH + c) * W_CHUNK) >> 4) + 2 * k, HI64, idesc, s != 0);
          }
        umma_commit(&aempty[as]);
        umma_commit(&tfull[acc]);
      }
      __syncwarp();
      if (++as == NA) { as = 0; aph ^= 1; }
      if (++acc == 2) { acc = 0; tph ^= 1; }
    }
  } else {
    const int quarter = warp % 4, half = (warp - 2) / 4, ew = warp - 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    uint8_t* stg = smem + OFF_STG + ew * STG_WARP;
    // cooperative store of the staged 32 x 64-byte block: lane j moves the 16-byte piece (row 8i + j/4, piece j%4)
    auto store_block = [&](bf16* dst_rows, int c0) {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = 8 * i + (lane >> 2), pc = lane & 3;
        const uint4 v = *reinterpret_cast<const uint4*>(stg + row * 64 + ((pc ^ ((row >> 1) & 3)) << 4));
        *reinterpret_cast<uint4*>(dst_rows + (long long)row * COUT + c0 + pc * 8) = v;
      }
      __syncwarp();
    };
    int rs = 0, acc = 0;
    uint32_t rph = 0, tph = 0;
    for (int tt = first; tt < p.tiles; tt += stride) {
      const int b = tt / blocks_per_img, m0 = (tt % blocks_per_img) * BM;
      const long long row0 = (long long)(b * 8 + t) * p.HW + m0 + quarter * 32;  // first row of this warp
      mbar_wait(&rfull[rs], rph);
      mbar_wait(&tfull[acc], tph);
      tc_fence_after();
      const uint8_t* rbox = smem + OFF_R + rs * R_STAGE + half * R_BOX + r * 128;
      const uint32_t t_addr = tmem_base + lane_addr + acc * 2 * COUT + half * 64;
#pragma unroll
      for (int ck = 0; ck < 2; ++ck) {
        const int c0 = half * 64 + ck * 32;  // first output column of this chunk
        float x[32];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 u = *reinterpret_cast<const uint4*>(rbox + (((4 * ck + j) ^ (r & 7)) << 4));
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            x[8 * j + 2 * i] = __uint_as_float(w[i] << 16);
            x[8 * j + 2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
          }
        }
        uint32_t q[16];
        {
          float v[32];
          tmem_ld32(t_addr + ck * 32, v);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + i);
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int e = 0; e < 4; e += 2) {
              const float a0 = x[i + e] + act_fast(v[i + e] + bb[e], ACT_ELU);
              const float a1 = x[i + e + 1] + act_fast(v[i + e + 1] + bb[e + 1], ACT_ELU);
              __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
              const uint32_t w = *reinterpret_cast<uint32_t*>(&h);
              q[(i + e) >> 1] = w;
              // the flow branch adds onto the STORED (bf16) value, as the un-fused pair of launches did
              x[i + e] = __uint_as_float(w << 16);
              x[i + e + 1] = __uint_as_float(w & 0xffff0000u);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
              make_uint4(q[4 * j], q[4 * j + 1], q[4 * j + 2], q[4 * j + 3]);
        store_block(p.dst_a + row0 * COUT, c0);
        {
          float v[32];
          tmem_ld32(t_addr + COUT + ck * 32, v);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + COUT + c0 + i);
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int e = 0; e < 4; e += 2) {
              const float a0 = x[i + e] + act_fast(v[i + e] + bb[e], ACT_ELU);
              const float a1 = x[i + e + 1] + act_fast(v[i + e + 1] + bb[e + 1], ACT_ELU);
              __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
              q[(i + e) >> 1] = *reinterpret_cast<uint32_t*>(&h);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
              make_uint4(q[4 * j], q[4 * j + 1], q[4 * j + 2], q[4 * j + 3]);
        store_block(p.dst_b + row0 * COUT, c0);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&rempty[rs]);
        mbar_arrive(&tempty[acc]);
      }
      if (++rs == NR) { rs = 0; rph ^= 1; }
      if (++acc == 2) { acc = 0; tph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

bool tc_resadd2_supported(int HW, int Cin, int Cout) { return Cin == CIN && Cout == COUT && HW % BM == 0; }

// skip_a, skip_b: bf16 [B, HW, 96]; src: bf16 [B, 8, HW, 128]; dst_a = src + ELU(skip_a . Wa[t] + ba) (may alias src),
// dst_b = dst_a + ELU(skip_b . Wb[t] + bb); wa_tc / wb_tc: bf16 [8][128][96] (K-major copies of the collapsed kernels)
void tc_resadd2(Ctx& c, const void* skip_a, const void* skip_b, const void* src, void* dst_a, void* dst_b, const void* wa_tc,
                const float* bias_a, const void* wb_tc, const float* bias_b, int B, int HW) {
  if (!c.ok() || c.dry) return;
  if (!wa_tc || !wb_tc || !bias_a || !bias_b || HW % BM) { c.fail(SJ_EUNSUPPORTED); return; }
  CUtensorMap mapA1, mapA2, mapW1, mapW2, mapR;
  uint64_t da[2] = {(uint64_t)CIN, (uint64_t)B * HW};
  uint64_t sa[1] = {(uint64_t)CIN * 2};
  uint32_t ba[2] = {KC, BM};
  uint64_t dw[2] = {(uint64_t)CIN, (uint64_t)8 * COUT};
  uint32_t bw[2] = {KC, COUT};
  uint64_t dr[2] = {(uint64_t)COUT, (uint64_t)B * 8 * HW};
  uint64_t sr[1] = {(uint64_t)COUT * 2};
  uint32_t br[2] = {64, BM};
  if (!encode_tmap(&mapA1, skip_a, 2, da, sa, ba, 64) || !encode_tmap(&mapA2, skip_b, 2, da, sa, ba, 64) ||
      !encode_tmap(&mapW1, wa_tc, 2, dw, sa, bw, 64) || !encode_tmap(&mapW2, wb_tc, 2, dw, sa, bw, 64) ||
      !encode_tmap(&mapR, src, 2, dr, sr, br, 128)) {
    snprintf(tls().cuda_err, sizeof(tls().cuda_err), "cuTensorMapEncodeTiled failed (tc_resadd2)");
    c.fail(SJ_ECUDA);
    return;
  }
  Res2P p{};
  p.B = B; p.HW = HW; p.tiles = B * (HW / BM);
  p.bias_a = bias_a; p.bias_b = bias_b;
  p.dst_a = (bf16*)dst_a; p.dst_b = (bf16*)dst_b;
  if (!SJ_SMEM_LIMIT_OK(tc_resadd2_kernel, 227 * 1024)) { c.fail(SJ_ECUDA); return; }
  int per_t = num_sms() / 8;
  if (per_t > p.tiles) per_t = p.tiles;
  if (per_t < 1) per_t = 1;
  SJ_LAUNCH(c, "tc_resadd2", tc_resadd2_kernel, 8 * per_t, NTHREADS, 1024 + SMEM_BYTES, mapA1, mapA2, mapW1, mapW2, mapR, p);
}

}  // namespace sj
