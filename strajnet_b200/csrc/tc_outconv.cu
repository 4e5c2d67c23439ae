// Decoder head on the tensor cores (K10): output_layer (on x) and output_layer_f (on fx), two 3x3 SAME
// 48->2 convolutions (modules.py:767-770), plus the [B,8,256,256,4] -> [B,256,256,32] transpose (:838).
//
// HBM-bound: 1.6 GB of bf16 activations are read once, 134 MB of fp32 logits are written.  One work item =
// (sample b, 16x8 pixel tile); it loops over the 8 waypoints x 2 heads.  For each of those 16 sub-items one
// 4-D TMA box {64 ch (48 real, rest zero-filled), 10, 18, 1} stages the tile + halo; the nine taps are shifted
// UMMA descriptor views of that patch (see tc_conv.cu), accumulated into a 16-column TMEM accumulator
// (N = 16 is the smallest M=128 MMA; 2 columns are real).  The epilogue gathers the 16 x 2 logits of a pixel
// and writes its 32 output channels as one contiguous 128-byte line.  Weights (36 KB) stay resident.
#include <cstdio>

#include "kernels.h"
#include "tc_common.cuh"

namespace sj {
namespace {

using namespace tc;

constexpr int TH = 16, TW = 8, PH = TH + 2, PW = TW + 2, NTHREADS = 192;
constexpr int A_SLOT = ((PH * PW * 128) + 1023) & ~1023;  // 23552
constexpr int NA = 6;
constexpr int W_TAP = 16 * 128;                             // one tap of one head: 16 rows x 128 B
constexpr int W_BYTES = 2 * 9 * W_TAP;                      // 36 KB

struct OutP {
  int B, num_tiles, out_layout;
  const float* bias;  // [2][2]
  float* out;
};

__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float& a, float& b) {
  uint32_t r0, r1;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  a = __uint_as_float(r0);
  b = __uint_as_float(r1);
}

__global__ void __launch_bounds__(NTHREADS, 1)
tc_outconv_kernel(const __grid_constant__ CUtensorMap mapO, const __grid_constant__ CUtensorMap mapF,
                  const __grid_constant__ CUtensorMap mapW, const OutP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_w = smem + NA * A_SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_w + W_BYTES);
  uint64_t* afull = bars;
  uint64_t* aempty = bars + NA;
  uint64_t* wfull = bars + 2 * NA;
  uint64_t* tfull_bar = wfull + 1;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapO);
    prefetch_tmap(&mapF);
    prefetch_tmap(&mapW);
    for (int s = 0; s < NA; ++s) {
      mbar_init(&afull[s], 1);
      mbar_init(&aempty[s], 1);
    }
    mbar_init(wfull, 1);
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
  constexpr int TILES_X = 256 / TW, TILES_PER_IMG = TILES_X * (256 / TH);

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(wfull, W_BYTES);
      for (int head = 0; head < 2; ++head)
        for (int tap = 0; tap < 9; ++tap)
          tma_load_2d(smem_w + (head * 9 + tap) * W_TAP, &mapW, wfull, tap * 64, head * 16);
      int slot = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < p.num_tiles; item += gridDim.x) {
        const int b = item / TILES_PER_IMG, tr = item % TILES_PER_IMG;
        const int y0 = (tr / TILES_X) * TH, x0 = (tr % TILES_X) * TW;
        for (int th = 0; th < 16; ++th) {
          mbar_wait(&aempty[slot], ph ^ 1);
          mbar_expect_tx(&afull[slot], PH * PW * 128);
          tma_load_4d(smem_a + slot * A_SLOT, (th & 1) ? &mapF : &mapO, &afull[slot], 0, x0 - 1, y0 - 1, b * 8 + (th >> 1));
          if (++slot == NA) { slot = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 16);
      constexpr uint32_t A_HI = desc_hi(128, PW * 128), B_HI = desc_hi(128, 8 * 128);
      const uint32_t w_lo = desc_lo(smem_u32(smem_w));
      int slot = 0, acc = 0;
      uint32_t ph = 0, tph = 0;
      mbar_wait(wfull, 0);
      tc_fence_after();
      for (int item = blockIdx.x; item < p.num_tiles; item += gridDim.x) {
        mbar_wait(&tempty_bar[acc], tph ^ 1);
        tc_fence_after();
#pragma unroll 1
        for (int th = 0; th < 16; ++th) {
          mbar_wait(&afull[slot], ph);
          tc_fence_after();
          const uint32_t a_lo = desc_lo(smem_u32(smem_a + slot * A_SLOT));
          const uint32_t d_tmem = tmem_base + acc * 256 + th * 16;
          const uint32_t wh = w_lo + (uint32_t)((th & 1) * 9 * W_TAP >> 4);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t va = a_lo + ((((tap / 3) * PW + tap % 3) * 128) >> 4);
            const uint32_t vb = wh + ((tap * W_TAP) >> 4);
#pragma unroll
            for (int k = 0; k < 3; ++k)  // channels 48..63 of the box are zero-filled: skip the 4th K step
              umma_bf16_w(d_tmem, va + 2 * k, A_HI, vb + 2 * k, B_HI, idesc, (tap | k) != 0);
          }
          umma_commit(&aempty[slot]);
          if (++slot == NA) { slot = 0; ph ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; tph ^= 1; }
      }
    }
  } else {
    const int quarter = warp % 4;
    const int r = quarter * 32 + lane, ty = r / TW, tx = r % TW;
    const float b00 = p.bias[0], b01 = p.bias[1], b10 = p.bias[2], b11 = p.bias[3];
    int acc = 0;
    uint32_t tph = 0;
    for (int item = blockIdx.x; item < p.num_tiles; item += gridDim.x) {
      const int b = item / TILES_PER_IMG, tr = item % TILES_PER_IMG;
      const int y = (tr / TILES_X) * TH + ty, x = (tr % TILES_X) * TW + tx;
      mbar_wait(&tfull_bar[acc], tph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 256;
      float v[32];
#pragma unroll
      for (int th = 0; th < 16; ++th) {
        tmem_ld2(t_addr + th * 16, v[2 * th], v[2 * th + 1]);
        v[2 * th] += (th & 1) ? b10 : b00;
        v[2 * th + 1] += (th & 1) ? b11 : b01;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (p.out_layout == 1) {  // [B,256,256,32], channel = t*4 + head*2 + o
        float4* o = reinterpret_cast<float4*>(p.out + (((long long)b * 256 + y) * 256 + x) * 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
      } else {  // [B,8,256,256,4]
#pragma unroll
        for (int t = 0; t < 8; ++t)
          *reinterpret_cast<float4*>(p.out + ((((long long)b * 8 + t) * 256 + y) * 256 + x) * 4) =
              make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
      }
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

// x_occ, x_flow: bf16 [B*8,256,256,48]; w_tc: bf16 [2 heads x 16 rows][9 taps x 64] (rows >= 2 and channels >= 48
// zero); bias fp32 [2][2]; out fp32
void tc_out_conv(Ctx& c, const void* x_occ, const void* x_flow, const void* w_tc, const float* bias, int B,
                 int out_layout, float* out) {
  if (!c.ok() || c.dry) return;
  CUtensorMap mapO, mapF, mapW;
  uint64_t da[4] = {48, 256, 256, (uint64_t)B * 8};
  uint64_t sa[3] = {96, 96 * 256, 96 * 65536};
  uint32_t ba[4] = {64, PW, PH, 1};
  uint64_t dw[2] = {9 * 64, 32};
  uint64_t sw[1] = {9 * 64 * 2};
  uint32_t bw[2] = {64, 16};
  if (!encode_tmap(&mapO, x_occ, 4, da, sa, ba, 128) || !encode_tmap(&mapF, x_flow, 4, da, sa, ba, 128) ||
      !encode_tmap(&mapW, w_tc, 2, dw, sw, bw, 128)) {
    snprintf(tls().cuda_err, sizeof(tls().cuda_err), "cuTensorMapEncodeTiled failed (tc_out_conv)");
    c.fail(SJ_ECUDA);
    return;
  }
  OutP p{};
  p.B = B; p.num_tiles = B * (256 / TW) * (256 / TH); p.out_layout = out_layout; p.bias = bias; p.out = out;
  const size_t smem = 1024 + (size_t)NA * A_SLOT + W_BYTES + 512;
  if (cudaFuncSetAttribute(tc_outconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
    c.fail(SJ_ECUDA);
    return;
  }
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  SJ_LAUNCH(c, "tc_out_conv", tc_outconv_kernel, grid, NTHREADS, smem, mapO, mapF, mapW, p);
}

}  // namespace sj
