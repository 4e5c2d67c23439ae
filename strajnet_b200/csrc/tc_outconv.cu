// Decoder head on the tensor cores (K10): output_layer (on x) and output_layer_f (on fx), two 3x3 SAME
// 48->2 convolutions (modules.py:767-770), plus the [B,8,256,256,4] -> [B,256,256,32] transpose (:838).
//
// HBM-bound: 1.6 GB of bf16 activations are read once, 134 MB of fp32 logits are written.  A 3x3 conv with
// 2 output channels is a poor MMA shape (every MMA re-reads its 128x16 A slice from shared memory whatever N
// is), so the contraction is re-associated:  out[p,o] = sum_tap Z[p + off(tap), tap, o]  with the pointwise
// projection  Z[q, (tap,o)] = sum_c x[q,c] * W[tap,c,o]  (18 columns, padded to N = 32).
//   * work item = (sample b, 16x8 pixel tile), looping over the 8 waypoints x 2 heads (16 sub-items);
//   * per sub-item ONE 4-D TMA box {64 ch (48 real, rest zero-filled), 10, 18, 1} stages the tile + halo
//     (180 pixels); two M=128 MMA blocks x 3 K-steps project all of them to Z in TMEM (fp32);
//   * the 128 epilogue threads move Z to shared memory, then each thread (= output pixel) adds the nine
//     shifted Z entries of its two logits; after 16 sub-items it writes its pixel's 32 output channels as
//     one contiguous 128-byte line.
// Weights (8 KB) stay resident.  Warp 0 = TMA, warp 1 = tcgen05.mma, warps 2..5 = epilogue.
#include <cstdio>

#include "kernels.h"
#include "tc_common.cuh"

namespace sj {
namespace {

using namespace tc;

constexpr int TH = 16, TW = 8, PH = TH + 2, PW = TW + 2, NPIX = PH * PW, NTHREADS = 192;
constexpr int A_SLOT = ((NPIX * 128) + 1023) & ~1023;  // 23552 B; the second MMA block reads 9 KB past it (valid smem)
constexpr int NA = 6, NZ = 4;
constexpr int W_HEAD = 32 * 128;                         // one head: 32 rows (tap*2+o, 18 real) x 128 B
constexpr int ZS_STRIDE = 20;                            // floats per pixel in the Z staging buffer
constexpr int ZS_BYTES = NPIX * ZS_STRIDE * 4;           // 14400
constexpr int OFF_W = NA * A_SLOT;
constexpr int OFF_ZS = OFF_W + 2 * W_HEAD + 9216;        // slack so block 1 of the last slot never reads Zs/barriers
constexpr int OFF_BAR = OFF_ZS + 2 * ZS_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256;

struct OutP {
  int B, num_tiles, out_layout;
  const float* bias;  // [2][2]
  void* out;
};

__global__ void __launch_bounds__(NTHREADS, 1)
tc_outconv_kernel(const __grid_constant__ CUtensorMap mapO, const __grid_constant__ CUtensorMap mapF,
                  const __grid_constant__ CUtensorMap mapW, const OutP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* afull = bars;
  uint64_t* aempty = bars + NA;
  uint64_t* wfull = bars + 2 * NA;
  uint64_t* zfull = wfull + 1;
  uint64_t* zempty = zfull + NZ;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(zempty + NZ);

  const int warp = uniform_warp_idx(), lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapO);
    prefetch_tmap(&mapF);
    prefetch_tmap(&mapW);
    for (int s = 0; s < NA; ++s) {
      mbar_init(&afull[s], 1);
      mbar_init(&aempty[s], 1);
    }
    mbar_init(wfull, 1);
    for (int z = 0; z < NZ; ++z) {
      mbar_init(&zfull[z], 1);
      mbar_init(&zempty[z], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr int TILES_X = 256 / TW, TILES_PER_IMG = TILES_X * (256 / TH);
  if (warp != 0) {  // the producer lane waits after it has issued the (constant) resident weights
    pdl_wait();
    pdl_trigger();
  }

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(wfull, 2 * W_HEAD);
      for (int head = 0; head < 2; ++head) tma_load_2d(smem + OFF_W + head * W_HEAD, &mapW, wfull, 0, head * 32);
      pdl_wait();
      int slot = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < p.num_tiles; item += gridDim.x) {
        const int b = item / TILES_PER_IMG, tr = item % TILES_PER_IMG;
        const int y0 = (tr / TILES_X) * TH, x0 = (tr % TILES_X) * TW;
        for (int th = 0; th < 16; ++th) {
          mbar_wait(&aempty[slot], ph ^ 1);
          mbar_expect_tx(&afull[slot], NPIX * 128);
          tma_load_4d(smem + slot * A_SLOT, (th & 1) ? &mapF : &mapO, &afull[slot], 0, x0 - 1, y0 - 1, b * 8 + (th >> 1));
          if (++slot == NA) { slot = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // uniform loop over the whole warp (descriptors in uniform registers); one elected lane issues
    const uint32_t idesc = make_idesc_bf16(128, 32);
    constexpr uint32_t HI = desc_hi(128, 1024);
    const uint32_t w_lo = desc_lo(smem_u32(smem + OFF_W));
    int slot = 0, zb = 0;
    uint32_t ph = 0, zph = 0;
    mbar_wait(wfull, 0);
    tc_fence_after();
    for (int item = blockIdx.x; item < p.num_tiles; item += gridDim.x) {
#pragma unroll 1
      for (int th = 0; th < 16; ++th) {
        mbar_wait(&afull[slot], ph);
        mbar_wait(&zempty[zb], zph ^ 1);
        tc_fence_after();
        const uint32_t a_lo = desc_lo(smem_u32(smem + slot * A_SLOT));
        const uint32_t wh = w_lo + (uint32_t)((th & 1) * (W_HEAD >> 4));
        if (elect_one()) {
#pragma unroll
          for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int k = 0; k < 3; ++k)  // channels 48..63 of the box are zero-filled: skip the 4th K step
              umma_bf16_w(tmem_base + zb * 64 + mb * 32, a_lo + mb * (128 * 128 >> 4) + 2 * k, HI, wh + 2 * k, HI, idesc,
                          k != 0);
          umma_commit(&aempty[slot]);
          umma_commit(&zfull[zb]);
        }
        __syncwarp();
        if (++slot == NA) { slot = 0; ph ^= 1; }
        if (++zb == NZ) { zb = 0; zph ^= 1; }
      }
    }
  } else {
    const int quarter = warp % 4;
    const int r = quarter * 32 + lane, ty = r / TW, tx = r % TW;
    float* zs = reinterpret_cast<float*>(smem + OFF_ZS);
    const float b00 = p.bias[0], b01 = p.bias[1], b10 = p.bias[2], b11 = p.bias[3];
    int zb = 0, sub = 0;
    uint32_t zph = 0;
    for (int item = blockIdx.x; item < p.num_tiles; item += gridDim.x) {
      const int b = item / TILES_PER_IMG, tr = item % TILES_PER_IMG;
      const int y = (tr / TILES_X) * TH + ty, x = (tr % TILES_X) * TW + tx;
      float v[32];
#pragma unroll
      for (int th = 0; th < 16; ++th, ++sub) {
        float* zbuf = zs + (sub & 1) * (ZS_BYTES / 4);
        mbar_wait(&zfull[zb], zph);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + zb * 64;
        float z0[32], z1[32];
        tmem_ld32(t_addr, z0);       // patch pixel r
        tmem_ld32(t_addr + 32, z1);  // patch pixel 128 + r (valid below 180)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&zempty[zb]);
        if (++zb == NZ) { zb = 0; zph ^= 1; }
        {
          float4* d0 = reinterpret_cast<float4*>(zbuf + r * ZS_STRIDE);
#pragma unroll
          for (int j = 0; j < 5; ++j) d0[j] = make_float4(z0[4 * j], z0[4 * j + 1], z0[4 * j + 2], z0[4 * j + 3]);
          if (128 + r < NPIX) {
            float4* d1 = reinterpret_cast<float4*>(zbuf + (128 + r) * ZS_STRIDE);
#pragma unroll
            for (int j = 0; j < 5; ++j) d1[j] = make_float4(z1[4 * j], z1[4 * j + 1], z1[4 * j + 2], z1[4 * j + 3]);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // the 128 epilogue threads only
        float a0 = (th & 1) ? b10 : b00, a1 = (th & 1) ? b11 : b01;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const float2 zz = *reinterpret_cast<const float2*>(zbuf + ((ty + tap / 3) * PW + tx + tap % 3) * ZS_STRIDE + 2 * tap);
          a0 += zz.x;
          a1 += zz.y;
        }
        v[2 * th] = a0;
        v[2 * th + 1] = a1;
      }
      float* outf = reinterpret_cast<float*>(p.out);
      if (p.out_layout == 2) {  // submission bytes (inference.py:160-182), 32 per pixel
        uint32_t q[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) q[t] = quantize_waypoint(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.out) + (((long long)b * 256 + y) * 256 + x) * 32);
        o[0] = make_uint4(q[0], q[1], q[2], q[3]);
        o[1] = make_uint4(q[4], q[5], q[6], q[7]);
      } else if (p.out_layout == 1) {  // [B,256,256,32], channel = t*4 + head*2 + o
        float4* o = reinterpret_cast<float4*>(outf + (((long long)b * 256 + y) * 256 + x) * 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
      } else {  // [B,8,256,256,4]
#pragma unroll
        for (int t = 0; t < 8; ++t)
          *reinterpret_cast<float4*>(outf + ((((long long)b * 8 + t) * 256 + y) * 256 + x) * 4) =
              make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

// x_occ, x_flow: bf16 [B*8,256,256,48]; w_tc: bf16 [2 heads][32 rows = tap*2+o (18 real)][64 ch (48 real)];
// bias fp32 [2][2]; out fp32
void tc_out_conv(Ctx& c, const void* x_occ, const void* x_flow, const void* w_tc, const float* bias, int B,
                 int out_layout, void* out) {
  if (!c.ok() || c.dry) return;
  CUtensorMap mapO, mapF, mapW;
  uint64_t da[4] = {48, 256, 256, (uint64_t)B * 8};
  uint64_t sa[3] = {96, 96 * 256, 96 * 65536};
  uint32_t ba[4] = {64, PW, PH, 1};
  uint64_t dw[2] = {64, 64};
  uint64_t sw[1] = {128};
  uint32_t bw[2] = {64, 32};
  if (!encode_tmap(&mapO, x_occ, 4, da, sa, ba, 128) || !encode_tmap(&mapF, x_flow, 4, da, sa, ba, 128) ||
      !encode_tmap(&mapW, w_tc, 2, dw, sw, bw, 128)) {
    snprintf(tls().cuda_err, sizeof(tls().cuda_err), "cuTensorMapEncodeTiled failed (tc_out_conv)");
    c.fail(SJ_ECUDA);
    return;
  }
  OutP p{};
  p.B = B; p.num_tiles = B * (256 / TW) * (256 / TH); p.out_layout = out_layout; p.bias = bias; p.out = out;
  const size_t smem = 1024 + SMEM_BYTES;
  if (!SJ_SMEM_LIMIT_OK((tc_outconv_kernel), 227 * 1024)) {
    c.fail(SJ_ECUDA);
    return;
  }
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  SJ_LAUNCH(c, "tc_out_conv", tc_outconv_kernel, grid, NTHREADS, smem, mapO, mapF, mapW, p);
}

}  // namespace sj
