// Decoder up-convolution, mid-channel variant (K8, the 128 -> 96 layers): nearest x2 upsample + 3x3 SAME conv + bias +
// ELU (modules.py:746-749) as four 2x2 sub-pixel convolutions on the low-res input (weights.fold_upconv_subpixel),
// ONE output phase (py, px) per CTA.
//
// Why: tc_upconv_kernel streams the folded weights of a row phase (8 tiles of [96 x 128], 196 KB) from L2 for every
// 128-pixel tile -- 1.6 GB of L2 -> SM traffic per launch, and the tensor pipe idles half the time waiting for it.  The
// four 2x2 taps of ONE phase are only 96 KB, so with a phase per CTA they stay resident in shared memory for the whole
// persistent kernel and the only streamed operand is the input patch (17 x 9 pixels x 128 channels per tile, 2
// SWIZZLE_128B chunks, one 4-D TMA box each; out-of-image pixels zero-filled = SAME padding).  The taps are shifted
// UMMA-descriptor views of the patch (start + (a*9 + b) pixel rows, SBO = 9 pixel rows).
// CTA i serves phase i % 4 and strides over the tiles with the other CTAs of that phase.
// Warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (32 MMAs of N = 96 per tile), warps 2..9 = epilogue (two warps per
// TMEM lane quarter, 48 columns each); four accumulator stages in TMEM.
#include <cstdio>

#include "kernels.h"
#include "tc_common.cuh"

namespace sj {
namespace {

using namespace tc;

constexpr int CIN = 128, COUT = 96, KC = 64, NCH = CIN / KC;  // 64-channel chunks, 128-byte swizzled rows
constexpr int TH = 16, TW = 8, PH = TH + 1, PW = TW + 1;
constexpr int NTHREADS = 320;
constexpr int A_SUB = (PH * PW * KC * 2 + 1023) & ~1023;  // 19584 -> 20480 B
constexpr int A_SLOT = NCH * A_SUB;                       // 40 KB
constexpr int NA = 2;                                     // (3 before the output staging buffer took 28 KB)
constexpr int B_TILE = COUT * KC * 2;                     // 12 KB: [96 rows][128 B]
constexpr int B_BYTES = NCH * 4 * B_TILE;                 // 96 KB
constexpr int STG_ROW = 112;                              // 96 bytes of one pixel (48 channels) + 16: conflict-free rows
constexpr int STG_WARP = 32 * STG_ROW;                    // 3.5 KB per epilogue warp
constexpr int OFF_B = NA * A_SLOT;
constexpr int OFF_STG = OFF_B + B_BYTES;
constexpr int OFF_BAR = OFF_STG + 8 * STG_WARP;
constexpr int SMEM_BYTES = OFF_BAR + 1024;
constexpr int NACC = 4, ACC_COLS = 128;                   // accumulator stages (96 columns used of each 128)

struct Up1P {
  int NB, H, W, tiles_x, tiles_y, num_tiles;
  const float* bias;
  bf16* out;  // [NB, 2H, 2W, COUT]
};

__global__ void __launch_bounds__(NTHREADS, 1)
tc_upconv1p_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const Up1P p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem + OFF_B;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* afull = bars;               // [NA]
  uint64_t* aempty = bars + NA;         // [NA]
  uint64_t* bfull = bars + 2 * NA;      // weights resident
  uint64_t* tfull = bfull + 1;          // [NACC]
  uint64_t* tempty = tfull + NACC;      // [NACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + NACC);
  float* bias_s = reinterpret_cast<float*>(bars + 32);  // [COUT]

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
    for (int a = 0; a < NACC; ++a) {
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
  const int phase = blockIdx.x & 3, py = phase >> 1, px = phase & 1;
  const int t_first = blockIdx.x >> 2, t_step = gridDim.x >> 2;
  if (warp != 0) {  // the producer lane waits after it has issued the (constant) resident weights
    pdl_wait();
    pdl_trigger();
  }

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bfull, B_BYTES);
      for (int ch = 0; ch < NCH; ++ch)
        for (int tap = 0; tap < 4; ++tap)  // tap = a*2 + b
          tma_load_2d(smem_b + (ch * 4 + tap) * B_TILE, &mapB, bfull, tap * CIN + ch * KC, phase * COUT);
      pdl_wait();
      int as_ = 0;
      uint32_t aph = 0;
      for (int t = t_first; t < p.num_tiles; t += t_step) {
        const int n = t / tiles_per_img, tr = t % tiles_per_img;
        const int y0 = (tr / p.tiles_x) * TH, x0 = (tr % p.tiles_x) * TW;
        mbar_wait(&aempty[as_], aph ^ 1);
        mbar_expect_tx(&afull[as_], NCH * PH * PW * KC * 2);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
          tma_load_4d(smem + as_ * A_SLOT + ch * A_SUB, &mapA, &afull[as_], ch * KC, x0 - 1 + px, y0 - 1 + py, n);
        if (++as_ == NA) { as_ = 0; aph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // the whole warp runs the loop (uniform control flow, operands in uniform registers); one elected lane issues
    constexpr uint32_t A_HI = desc_hi(KC * 2, PW * KC * 2), B_HI = desc_hi(KC * 2, 8 * KC * 2);
    const uint32_t idesc = make_idesc_bf16(128, COUT);
    const uint32_t b_lo = desc_lo(smem_u32(smem_b));
    int as_ = 0, acc = 0;
    uint32_t aph = 0, tph = 0;
    mbar_wait(bfull, 0);
    tc_fence_after();
    for (int t = t_first; t < p.num_tiles; t += t_step) {
      mbar_wait(&tempty[acc], tph ^ 1);
      mbar_wait(&afull[as_], aph);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
      const uint32_t a_lo = desc_lo(smem_u32(smem + as_ * A_SLOT));
      if (elect_one()) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
#pragma unroll
          for (int tap = 0; tap < 4; ++tap) {
            // shifted view of the staged patch: MMA row 8*ty+tx -> patch pixel (ty + a, tx + b)
            const uint32_t va = a_lo + ((ch * A_SUB + ((tap >> 1) * PW + (tap & 1)) * KC * 2) >> 4);
            const uint32_t vb = b_lo + (((ch * 4 + tap) * B_TILE) >> 4);
#pragma unroll
            for (int k = 0; k < KC / 16; ++k)
              umma_bf16_w(d_tmem, va + 2 * k, A_HI, vb + 2 * k, B_HI, idesc, (ch | tap | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&aempty[as_]);
        umma_commit(&tfull[acc]);
      }
      __syncwarp();
      if (++as_ == NA) { as_ = 0; aph ^= 1; }
      if (++acc == NACC) { acc = 0; tph ^= 1; }
    }
  } else {
    const int quarter = warp % 4, half = (warp - 2) / 4;  // two warps per lane quarter, 48 columns each
    const int r = quarter * 32 + lane, ty = r / TW, tx = r % TW;
    const int c0 = half * (COUT / 2);
    float bias_r[COUT / 2];  // this thread's 48 output channels never change: bias lives in registers
#pragma unroll
    for (int i = 0; i < COUT / 2; ++i) bias_r[i] = bias_s[c0 + i];
    // Output rows leave through a per-warp shared-memory transpose: a thread's 96 bytes (its pixel, 48 channels) are 384
    // bytes away from its neighbour's in global memory, so direct 16-byte stores cost 32 L1 wavefronts per instruction;
    // after the transpose six consecutive lanes cover one pixel's 96 bytes (6-8 wavefronts per instruction).
    uint8_t* stg = smem + OFF_STG + (warp - 2) * STG_WARP;
    int s_off[6], g_off[6];  // piece g = 32 i + lane of the warp's 32 x 6 sixteen-byte pieces: pixel g / 6, piece g % 6
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int g = 32 * i + lane, pix = g / 6, pc = g % 6;
      const int pty = quarter * 4 + pix / TW, ptx = pix % TW;  // tile coordinates of that pixel
      s_off[i] = pix * STG_ROW + pc * 16;
      g_off[i] = ((2 * pty) * (2 * p.W) + 2 * ptx) * COUT * 2 + pc * 16;  // bytes from the tile's first output pixel
    }
    int acc = 0;
    uint32_t tph = 0;
    for (int t = t_first; t < p.num_tiles; t += t_step) {
      const int n = t / tiles_per_img, tr = t % tiles_per_img;
      const int y0 = (tr / p.tiles_x) * TH, x0 = (tr % p.tiles_x) * TW;
      mbar_wait(&tfull[acc], tph);
      tc_fence_after();
      uint8_t* tile0 = reinterpret_cast<uint8_t*>(
          p.out + (((long long)n * (2 * p.H) + 2 * y0 + py) * (2 * p.W) + 2 * x0 + px) * COUT + c0);
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * ACC_COLS + c0;
      uint4* row = reinterpret_cast<uint4*>(stg + lane * STG_ROW);
      {
        float v[32];
        tmem_ld32(t_addr, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = act_fast(v[i] + bias_r[i], ACT_ELU);
#pragma unroll
        for (int i = 0; i < 32; i += 8) st8_bf16(reinterpret_cast<bf16*>(row + i / 8), v + i);
      }
      {
        float v[16];
        tmem_ld16(t_addr + 32, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = act_fast(v[i] + bias_r[32 + i], ACT_ELU);
#pragma unroll
        for (int i = 0; i < 16; i += 8) st8_bf16(reinterpret_cast<bf16*>(row + 4 + i / 8), v + i);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);  // the accumulator is drained: the stores below overlap the next MMAs
#pragma unroll
      for (int i = 0; i < 6; ++i)
        *reinterpret_cast<uint4*>(tile0 + g_off[i]) = *reinterpret_cast<const uint4*>(stg + s_off[i]);
      __syncwarp();
      if (++acc == NACC) { acc = 0; tph ^= 1; }
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

bool tc_upconv1p_supported(int H, int W, int Cin, int Cout) {
  return Cin == CIN && Cout == COUT && H % TH == 0 && W % TW == 0;
}

// x bf16 [NB,H,W,128] -> y bf16 [NB,2H,2W,96]; w_tc = folded kernels [4 phases][96][4*128] bf16
void tc_upconv1p(Ctx& c, const void* x, void* y, const void* w_tc, const float* bias, int NB, int H, int W) {
  if (!c.ok() || c.dry) return;
  if (!w_tc || !bias || H % TH || W % TW) { c.fail(SJ_EUNSUPPORTED); return; }
  Up1P p{};
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
    snprintf(tls().cuda_err, sizeof(tls().cuda_err), "cuTensorMapEncodeTiled failed (tc_upconv1p)");
    c.fail(SJ_ECUDA);
    return;
  }
  const size_t smem = 1024 + SMEM_BYTES;
  if (!SJ_SMEM_LIMIT_OK((tc_upconv1p_kernel), 227 * 1024)) {
    c.fail(SJ_ECUDA);
    return;
  }
  int grid = num_sms() & ~3;  // CTA index mod 4 = output phase
  if (grid > 4 * p.num_tiles) grid = 4 * p.num_tiles;
  SJ_LAUNCH(c, "tc_upconv1p", tc_upconv1p_kernel, grid, NTHREADS, smem, mapA, mapB, p);
}

}  // namespace sj
