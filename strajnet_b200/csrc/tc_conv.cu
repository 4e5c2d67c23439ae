// Decoder up-convolution on the tensor cores (K8): nearest x2 upsample + 3x3 SAME conv + bias + ELU
// (modules.py:746-749), executed as four 2x2 sub-pixel convolutions on the LOW-RES input with
// pre-summed taps (weights.fold_upconv_subpixel): 2.25x fewer FLOPs, 4x fewer input bytes.
//
// Implicit GEMM, no im2col buffer: one work item = (image, 8x16 low-res pixel tile, output row phase py).
// It produces both column phases px in two TMEM accumulators [128 pixels x Cout].  The K loop walks
// the 2 row taps x 3 column taps x Cin chunks; for every tap the A tile is ONE 4-D TMA box
// {channels, 16, 8, 1} of the NHWC input at a shifted origin -- out-of-image rows/columns are
// zero-filled by TMA, which is exactly the SAME padding.  A column tap feeds one (dx = +-1) or two
// (dx = 0) column phases, so the matching folded weight tiles [Cout x chunk] ride in the same stage.
//
// Same skeleton as tc_gemm.cu: persistent CTAs, warp 0 = TMA, warp 1 = tcgen05.mma, warps 2..5 =
// epilogue, mbarrier ring, accumulators double-buffered in TMEM when 4*Cout fits 512 columns.
#include <cstdio>

#include "kernels.h"
#include "tc_common.cuh"

namespace sj {
namespace {

using namespace tc;

constexpr int TH = 8, TW = 16, BM = 128, NTHREADS = 192;

struct ConvP {
  int NB, H, W, Cin, Cout;  // low-res input geometry
  int tiles_x, tiles_y, num_items;
  int acs;                  // TMEM column stride between accumulators
  int nacc;                 // accumulator stages (1 or 2)
  int stages;
  const float* bias;
  bf16* out;                // [NB, 2H, 2W, Cout]
};

// KC: channels per swizzled smem row (64 -> SWIZZLE_128B, 32 -> SWIZZLE_64B); NCH: chunks per pipeline stage
template <int KC, int NCH>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_upconv_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const ConvP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int ROWB = KC * 2;                 // bytes per smem row
  constexpr int A_SUB = BM * ROWB;             // one A chunk
  const int b_sub = p.Cout * ROWB;             // one B chunk of one phase
  const int a_stage = NCH * A_SUB;
  const int stage_bytes = a_stage + 2 * NCH * b_sub;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + p.stages;
  uint64_t* tfull_bar = bars + 2 * p.stages;
  uint64_t* tempty_bar = bars + 2 * p.stages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 4);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int groups_per_tap = p.Cin / (KC * NCH);
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int py = item & 1, t = item >> 1;
        const int n = t / tiles_per_img, tr = t % tiles_per_img;
        const int y0 = (tr / p.tiles_x) * TH, x0 = (tr % p.tiles_x) * TW;
        for (int a = 0; a < 2; ++a) {
          const int dy = a - 1 + py;
          for (int dxi = 0; dxi < 3; ++dxi) {
            const int dx = dxi - 1;
            for (int cg = 0; cg < groups_per_tap; ++cg) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* sa = smem + stage * stage_bytes;
              uint8_t* sb = sa + a_stage;
              const int nb = dx == 0 ? 2 : 1;
              mbar_expect_tx(&full_bar[stage], a_stage + nb * NCH * b_sub);
#pragma unroll
              for (int ch = 0; ch < NCH; ++ch)
                tma_load_4d(sa + ch * A_SUB, &mapA, &full_bar[stage], (cg * NCH + ch) * KC, x0 + dx, y0 + dy, n);
              // slot 0: first phase fed by this column tap, slot 1: second (dx == 0 only)
              for (int s = 0; s < nb; ++s) {
                const int px = dx < 0 ? 0 : (dx > 0 ? 1 : s);
                const int b = dx - px + 1;  // column tap index inside phase px: low-res offset = b - 1 + px
                const int krow = ((a * 2 + b) * p.Cin);
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch)
                  tma_load_2d(sb + (s * NCH + ch) * b_sub, &mapB, &full_bar[stage], krow + (cg * NCH + ch) * KC,
                              (py * 2 + px) * p.Cout);
              }
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(BM, p.Cout);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        uint32_t started[2] = {0, 0};
        for (int a = 0; a < 2; ++a) {
          for (int dxi = 0; dxi < 3; ++dxi) {
            const int dx = dxi - 1;
            for (int cg = 0; cg < groups_per_tap; ++cg) {
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              const uint32_t sa = smem_u32(smem + stage * stage_bytes);
              const uint32_t sb = sa + a_stage;
              const int nb = dx == 0 ? 2 : 1;
              for (int s = 0; s < nb; ++s) {
                const int px = dx < 0 ? 0 : (dx > 0 ? 1 : s);
                const uint32_t d_tmem = tmem_base + (as * 2 + px) * p.acs;
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                  const uint64_t da = make_smem_desc(sa + ch * A_SUB, ROWB);
                  const uint64_t db = make_smem_desc(sb + (s * NCH + ch) * b_sub, ROWB);
#pragma unroll
                  for (int k = 0; k < KC / 16; ++k) {
                    umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, started[px]);
                    started[px] = 1;
                  }
                }
              }
              umma_commit(&empty_bar[stage]);
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
          }
        }
        umma_commit(&tfull_bar[as]);
        if (++as == p.nacc) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    const int quarter = warp % 4;
    const int r = quarter * 32 + lane, ty = r / TW, tx = r % TW;
    int as = 0;
    uint32_t aphase = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int py = item & 1, t = item >> 1;
      const int n = t / tiles_per_img, tr = t % tiles_per_img;
      const int yy = (tr / p.tiles_x) * TH + ty, xx = (tr % p.tiles_x) * TW + tx;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int px = 0; px < 2; ++px) {
        bf16* dst = p.out + (((long long)n * (2 * p.H) + 2 * yy + py) * (2 * p.W) + 2 * xx + px) * p.Cout;
        const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (as * 2 + px) * p.acs;
        int c = 0;
        for (; c + 32 <= p.Cout; c += 32) {
          float v[32];
          tmem_ld32(t_addr + c, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = act_fast(v[i] + p.bias[c + i], ACT_ELU);
#pragma unroll
          for (int i = 0; i < 32; i += 8) st8_bf16(dst + c + i, v + i);
        }
        if (c < p.Cout) {
          float v[16];
          tmem_ld16(t_addr + c, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = act_fast(v[i] + p.bias[c + i], ACT_ELU);
#pragma unroll
          for (int i = 0; i < 16; i += 8) st8_bf16(dst + c + i, v + i);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == p.nacc) { as = 0; aphase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int KC, int NCH>
void launch_upconv(Ctx& c, const void* x, const void* w_tc, ConvP p) {
  const int rowb = KC * 2;
  CUtensorMap mapA, mapB;
  uint64_t da[4] = {(uint64_t)p.Cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.NB};
  uint64_t sa[3] = {(uint64_t)p.Cin * 2, (uint64_t)p.W * p.Cin * 2, (uint64_t)p.H * p.W * p.Cin * 2};
  uint32_t ba[4] = {KC, TW, TH, 1};
  uint64_t db[2] = {(uint64_t)4 * p.Cin, (uint64_t)4 * p.Cout};
  uint64_t sb[1] = {(uint64_t)4 * p.Cin * 2};
  uint32_t bb[2] = {KC, (uint32_t)p.Cout};
  if (!encode_tmap(&mapA, x, 4, da, sa, ba, rowb) || !encode_tmap(&mapB, w_tc, 2, db, sb, bb, rowb)) {
    snprintf(tls().cuda_err, sizeof(tls().cuda_err), "cuTensorMapEncodeTiled failed (tc_upconv Cin=%d Cout=%d)", p.Cin, p.Cout);
    c.fail(SJ_ECUDA);
    return;
  }
  const int stage_bytes = NCH * BM * rowb + 2 * NCH * p.Cout * rowb;
  p.stages = (200 * 1024) / stage_bytes;
  if (p.stages > 8) p.stages = 8;
  if (p.stages < 2) { c.fail(SJ_EUNSUPPORTED); return; }
  size_t smem = 1024 + (size_t)p.stages * stage_bytes + 256;
  if (smem < 120 * 1024) smem = 120 * 1024;  // one CTA per SM: each CTA owns all 512 TMEM columns
  if (cudaFuncSetAttribute(tc_upconv_kernel<KC, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
    c.fail(SJ_ECUDA);
    return;
  }
  const int grid = p.num_items < num_sms() ? p.num_items : num_sms();
  SJ_LAUNCH(c, "tc_upconv", (tc_upconv_kernel<KC, NCH>), grid, NTHREADS, smem, mapA, mapB, p);
}

}  // namespace

bool tc_upconv_supported(int H, int W, int Cin, int Cout) {
  if (H % TH || W % TW || Cout % 16 || Cout < 16 || Cout > 256) return false;
  return Cin % 64 == 0 || Cin == 96;
}

// x bf16 [NB,H,W,Cin] -> y bf16 [NB,2H,2W,Cout]; w_tc = folded kernels [4][Cout][4*Cin] bf16
void tc_upconv(Ctx& c, const void* x, void* y, const void* w_tc, const float* bias, int NB, int H, int W, int Cin,
               int Cout) {
  if (!c.ok() || c.dry) return;
  if (!tc_upconv_supported(H, W, Cin, Cout) || !w_tc || !bias) { c.fail(SJ_EUNSUPPORTED); return; }
  ConvP p{};
  p.NB = NB; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.tiles_x = W / TW; p.tiles_y = H / TH;
  p.num_items = NB * p.tiles_x * p.tiles_y * 2;
  p.acs = (Cout + 63) / 64 * 64;
  p.nacc = 4 * p.acs <= 512 ? 2 : 1;
  p.bias = bias;
  p.out = (bf16*)y;
  if (Cin % 64 == 0) launch_upconv<64, 1>(c, x, w_tc, p);
  else launch_upconv<32, 3>(c, x, w_tc, p);
}

}  // namespace sj
