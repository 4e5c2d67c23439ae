// Decoder up-convolution, small-channel variant (K8, the 96 -> 48 layers that are 43 % of the forward's nominal FLOPs):
// nearest x2 upsample + 3x3 SAME conv + bias + ELU (modules.py:746-749) as four 2x2 sub-pixel convolutions on the
// low-res input (weights.fold_upconv_subpixel), with ALL FOUR output phases of a tile produced by one CTA.
//
// Why: with Cout = 48 a [128 x 16] x [16 x 48] MMA is bound by the shared-memory read of its A slice (32 cycles) rather
// than by its math (24 cycles), and tc_upconv_kernel issues 16 such MMAs per 16 input channels (4 phases x 4 taps).
// The 16 (phase, tap) products only touch 9 distinct shifted views of the staged input patch -- view (ro, dx) feeds
// every phase (py, px) with tap (a, b) = (ro - py, dx - px) in {0,1}^2 -- so the phases that share a view are stacked
// along N:
//   * accumulators of the four phases sit side by side in TMEM in ring order  p00 | p01 | p11 | p10  (48 columns each),
//     so that three of the four edge views address two ADJACENT phases;
//   * centre view: one MMA with N = 192; views (0,1), (1,2), (2,1): one MMA with N = 96; view (1,0): two N = 48 MMAs;
//     corners: N = 48.  10 MMAs instead of 16 per 16 channels, 10 instead of 16 A-slice reads;
//   * the [48 x Cin] weight tiles are laid out in shared memory per view in the same ring order (each is its own
//     TMA box of the folded weights [4 phases][Cout][4*Cin]), and stay resident for the whole persistent CTA;
//   * the input patch (tile + 1-pixel halo, 18 x 10 pixels x 96 channels) is staged ONCE per tile by three 4-D TMA
//     boxes (32-channel SWIZZLE_64B chunks); out-of-image pixels are zero-filled by TMA = SAME padding.  Every view
//     is a shifted UMMA descriptor over that patch (start + (ro*10 + dx) pixel rows, SBO = 10 pixel rows).
// Warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warps 2..9 = epilogue (warps 2..5: output rows 2y, warps 6..9:
// rows 2y+1; each thread writes its two horizontally adjacent output pixels = 192 contiguous bytes).  Accumulators
// are double-buffered in TMEM (2 x 192 columns).
#include <cstdio>

#include "kernels.h"
#include "tc_common.cuh"

namespace sj {
namespace {

using namespace tc;

constexpr int CIN = 96, COUT = 48, KC = 32, NCH = CIN / KC;  // 32-channel chunks, 64-byte swizzled rows
constexpr int TH = 16, TW = 8, PH = TH + 2, PW = TW + 2;
constexpr int NTHREADS = 320;
constexpr int A_SUB = (PH * PW * KC * 2 + 1023) & ~1023;  // one chunk of the patch: 11520 -> 12288 B
constexpr int A_SLOT = NCH * A_SUB;                       // 36 KB
constexpr int NA = 2;
constexpr int B_TILE = COUT * KC * 2;                     // 3072 B: [48 rows][64 B]
constexpr int B_BYTES = NCH * 16 * B_TILE;                // 144 KB
constexpr int OFF_B = NA * A_SLOT;
constexpr int OFF_BAR = OFF_B + B_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 1024;
constexpr int ACC_COLS = 4 * COUT;                        // 192 TMEM columns per accumulator stage

// shared-memory weight slots: (view ro, dx) x ring phase; ring 0..3 = (py,px) 00, 01, 11, 10
struct Slot { int ro, dx, ring; };
__host__ __device__ constexpr int ring_py(int r) { return r >> 1; }
__host__ __device__ constexpr int ring_px(int r) { return (r == 1 || r == 2) ? 1 : 0; }

struct Up4P {
  int NB, H, W, tiles_x, tiles_y, num_tiles;
  const float* bias;
  bf16* out;  // [NB, 2H, 2W, COUT]
};

__global__ void __launch_bounds__(NTHREADS, 1)
tc_upconv4_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const Up4P p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem + OFF_B;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* afull = bars;            // [NA]
  uint64_t* aempty = bars + NA;      // [NA]
  uint64_t* bfull = bars + 2 * NA;   // weights resident
  uint64_t* tfull = bfull + 1;       // [2]
  uint64_t* tempty = tfull + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* bias_s = reinterpret_cast<float*>(bars + 16);  // [COUT]

  const int warp = uniform_warp_idx(), lane = threadIdx.x % 32;
  for (int i = threadIdx.x; i < COUT; i += NTHREADS) bias_s[i] = p.bias[i];
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    for (int s = 0; s < NA; ++s) {
      mbar_init(&afull[s], 1);
      mbar_init(&aempty[s], 1);
    }
    mbar_init(bfull, 1);
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
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  if (warp != 0) {  // the producer lane waits after it has issued the (constant) resident weights
    pdl_wait();
    pdl_trigger();
  }

  // slot table: 0-3 centre (ring 0..3); 4,5 view (0,1); 6,7 view (1,2); 8,9 view (2,1); 10 view (1,0) ring 0;
  // 11 view (1,0) ring 3; 12..15 corners (0,0) r0, (0,2) r1, (2,2) r2, (2,0) r3
  constexpr Slot SLOTS[16] = {{1, 1, 0}, {1, 1, 1}, {1, 1, 2}, {1, 1, 3}, {0, 1, 0}, {0, 1, 1}, {1, 2, 1}, {1, 2, 2},
                              {2, 1, 2}, {2, 1, 3}, {1, 0, 0}, {1, 0, 3}, {0, 0, 0}, {0, 2, 1}, {2, 2, 2}, {2, 0, 3}};

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bfull, B_BYTES);
      for (int ch = 0; ch < NCH; ++ch)
        for (int s = 0; s < 16; ++s) {
          const int py = ring_py(SLOTS[s].ring), px = ring_px(SLOTS[s].ring);
          const int a = SLOTS[s].ro - py, b = SLOTS[s].dx - px;
          tma_load_2d(smem_b + (ch * 16 + s) * B_TILE, &mapB, bfull, (a * 2 + b) * CIN + ch * KC, (py * 2 + px) * COUT);
        }
      pdl_wait();
      int as_ = 0;
      uint32_t aph = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int n = t / tiles_per_img, tr = t % tiles_per_img;
        const int y0 = (tr / p.tiles_x) * TH, x0 = (tr % p.tiles_x) * TW;
        mbar_wait(&aempty[as_], aph ^ 1);
        mbar_expect_tx(&afull[as_], NCH * PH * PW * KC * 2);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
          tma_load_4d(smem + as_ * A_SLOT + ch * A_SUB, &mapA, &afull[as_], ch * KC, x0 - 1, y0 - 1, n);
        if (++as_ == NA) { as_ = 0; aph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // the whole warp runs the loop (uniform control flow, operands in uniform registers); one elected lane issues
    // one entry per MMA of a 16-channel step: view, first weight slot, first ring phase, stacked phases
    struct Op { int ro, dx, slot, ring, cnt; };
    constexpr Op OPS[10] = {{1, 1, 0, 0, 4}, {0, 1, 4, 0, 2}, {1, 2, 6, 1, 2}, {2, 1, 8, 2, 2}, {1, 0, 10, 0, 1},
                            {1, 0, 11, 3, 1}, {0, 0, 12, 0, 1}, {0, 2, 13, 1, 1}, {2, 2, 14, 2, 1}, {2, 0, 15, 3, 1}};
    constexpr uint32_t A_HI = desc_hi(KC * 2, PW * KC * 2), B_HI = desc_hi(KC * 2, 8 * KC * 2);
    const uint32_t idesc1 = make_idesc_bf16(128, COUT), idesc2 = make_idesc_bf16(128, 2 * COUT),
                   idesc4 = make_idesc_bf16(128, 4 * COUT);
    const uint32_t b_lo = desc_lo(smem_u32(smem_b));
    int as_ = 0, acc = 0;
    uint32_t aph = 0, tph = 0;
    mbar_wait(bfull, 0);
    tc_fence_after();
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      mbar_wait(&tempty[acc], tph ^ 1);
      mbar_wait(&afull[as_], aph);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
      const uint32_t a_lo = desc_lo(smem_u32(smem + as_ * A_SLOT));
      if (elect_one()) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) {
#pragma unroll
            for (int o = 0; o < 10; ++o) {
              const uint32_t va = a_lo + ((ch * A_SUB + (OPS[o].ro * PW + OPS[o].dx) * KC * 2) >> 4) + 2 * k;
              const uint32_t vb = b_lo + (((ch * 16 + OPS[o].slot) * B_TILE) >> 4) + 2 * k;
              const uint32_t idesc = OPS[o].cnt == 4 ? idesc4 : (OPS[o].cnt == 2 ? idesc2 : idesc1);
              // the centre view covers all 192 columns and is issued first: it alone initialises the accumulators
              umma_bf16_w(d_tmem + OPS[o].ring * COUT, va, A_HI, vb, B_HI, idesc, (o != 0 || (ch | k) != 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(&aempty[as_]);
        umma_commit(&tfull[acc]);
      }
      __syncwarp();
      if (++as_ == NA) { as_ = 0; aph ^= 1; }
      if (++acc == 2) { acc = 0; tph ^= 1; }
    }
  } else {
    const int quarter = warp % 4, py = (warp - 2) / 4;
    const int r = quarter * 32 + lane, ty = r / TW, tx = r % TW;
    // ring order p00 | p01 | p11 | p10: row phase 0 reads columns [0, 96) as (px0, px1), row phase 1 reads
    // [96, 192) as (px1, px0)
    float bias_r[COUT];  // bias lives in registers (both phases of this thread use the same 48 channels)
#pragma unroll
    for (int i = 0; i < COUT; ++i) bias_r[i] = bias_s[i];
    int acc = 0;
    uint32_t tph = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      const int n = t / tiles_per_img, tr = t % tiles_per_img;
      const int yy = (tr / p.tiles_x) * TH + ty, xx = (tr % p.tiles_x) * TW + tx;
      mbar_wait(&tfull[acc], tph);
      tc_fence_after();
      bf16* dst0 = p.out + (((long long)n * (2 * p.H) + 2 * yy + py) * (2 * p.W) + 2 * xx) * COUT;
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * ACC_COLS + py * 2 * COUT;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int px = py ? 1 - half : half;
        bf16* dst = dst0 + px * COUT;
        {
          float v[32];
          tmem_ld32(t_addr + half * COUT, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = act_fast(v[i] + bias_r[i], ACT_ELU);
#pragma unroll
          for (int i = 0; i < 32; i += 8) st8_bf16(dst + i, v + i);
        }
        {
          float v[16];
          tmem_ld16(t_addr + half * COUT + 32, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = act_fast(v[i] + bias_r[32 + i], ACT_ELU);
#pragma unroll
          for (int i = 0; i < 16; i += 8) st8_bf16(dst + 32 + i, v + i);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
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

bool tc_upconv4_supported(int H, int W, int Cin, int Cout) {
  return Cin == CIN && Cout == COUT && H % TH == 0 && W % TW == 0;
}

// x bf16 [NB,H,W,96] -> y bf16 [NB,2H,2W,48]; w_tc = folded kernels [4 phases][48][4*96] bf16
void tc_upconv4(Ctx& c, const void* x, void* y, const void* w_tc, const float* bias, int NB, int H, int W) {
  if (!c.ok() || c.dry) return;
  if (!w_tc || !bias || H % TH || W % TW) { c.fail(SJ_EUNSUPPORTED); return; }
  Up4P p{};
  p.NB = NB; p.H = H; p.W = W;
  p.tiles_x = W / TW; p.tiles_y = H / TH;
  p.num_tiles = NB * p.tiles_x * p.tiles_y;
  p.bias = bias;
  p.out = (bf16*)y;
  CUtensorMap mapA, mapB;
  uint64_t da[4] = {(uint64_t)CIN, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
  uint64_t sa[3] = {(uint64_t)CIN * 2, (uint64_t)W * CIN * 2, (uint64_t)H * W * CIN * 2};
  uint32_t ba[4] = {KC, PW, PH, 1};
  uint64_t db[2] = {(uint64_t)4 * CIN, (uint64_t)4 * COUT};
  uint64_t sb[1] = {(uint64_t)4 * CIN * 2};
  uint32_t bb[2] = {KC, COUT};
  if (!encode_tmap(&mapA, x, 4, da, sa, ba, KC * 2) || !encode_tmap(&mapB, w_tc, 2, db, sb, bb, KC * 2)) {
    snprintf(tls().cuda_err, sizeof(tls().cuda_err), "cuTensorMapEncodeTiled failed (tc_upconv4)");
    c.fail(SJ_ECUDA);
    return;
  }
  const size_t smem = 1024 + SMEM_BYTES;
  if (!SJ_SMEM_LIMIT_OK((tc_upconv4_kernel), 227 * 1024)) {
    c.fail(SJ_ECUDA);
    return;
  }
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  SJ_LAUNCH(c, "tc_upconv4", tc_upconv4_kernel, grid, NTHREADS, smem, mapA, mapB, p);
}

}  // namespace sj
