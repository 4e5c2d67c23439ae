// Second half of the fused decoder head (K10): the 9-tap shifted sum over the projected columns written by
// tc_upconv4h_kernel, for both heads, plus the [B,8,256,256,4] -> [B,256,256,32] transpose (modules.py:767-770, :838)
// and optionally the submission quantisation (inference.py:124-136, :160-182):
//   out[b, y, x, t*4 + head*2 + o] = bias[head][o] + sum_{dy,dx} Z_head[b*8+t, y+dy-1, x+dx-1, (dy*3+dx)*2 + o]
// with zero outside the image (SAME padding of the 3x3 convolution).
//
// HBM-bound by construction: 2 x 36 B read per (waypoint, pixel), 4 B (fp32) or 1 B (quantised) written per output
// element; no tensor cores (9 adds per output).  Work item = (sample, 8 x 32 pixel tile); its 16 (waypoint, head) halo
// tiles stream through a 4-stage ring filled by TMA: one thread issues two 3-D box loads per stage (the image rows are
// viewed as 2304 32-bit words = 256 pixels x 9 fp16 pairs; a 34-pixel halo row is 306 words, more than one box dimension
// may hold, so it arrives as two overlapping 184-word boxes whose starts are 16-byte aligned), out-of-image rows / columns
// are zero-filled by the TMA unit (= the SAME padding of the 3x3 convolution).  Each thread owns one pixel, reads its nine
// fp16 pairs per sub-item conflict-free (pixel stride 9 words; the second box starts 128 words into the row and 7424 bytes
// into the stage, which keeps the two halves of a warp on disjoint banks) and finally writes its 32 output channels as one 128-byte line.
// (The first version filled the ring with per-thread 16-byte cp.async: address selects and border predicates were ~55 %
// of the instructions of an instruction-bound kernel -- issue slots 58 % busy at 0.66 of the HBM peak.)
#include <cuda_fp16.h>

#include <cstdio>

#include "kernels.h"
#include "tc_common.cuh"

namespace sj {
namespace {

using namespace tc;

constexpr int TW = 32, TH = 8, NTHREADS = TW * TH;
constexpr int PIX_W = 9;                               // 32-bit words per pixel (18 fp16 columns)
constexpr int ROW_W = 256 * PIX_W;                     // 2304 words per image row
constexpr int LEAD_W = 3;                              // the box starts 3 words before the halo pixel x0-1: 16-byte aligned
constexpr int BOX_W = 184;                             // words per box row (A: words [0,184) from there, B: [128,312))
constexpr int B_WORD0 = 128;                           // first word of box B (a multiple of 32: both halves of a warp on the same bank map)
constexpr int SPLIT_TX = 17;                           // pixels tx < 17 read box A (needs words < 3 + 9*16 + 27 = 174)
constexpr int SROW = BOX_W * 4;                        // 736 bytes per staged row
constexpr int SUB_A = 0, SUB_B = 7424;                 // box offsets inside a stage (128-byte aligned)
constexpr int STAGE = 14848;                           // 7424 + 7360, rounded up to 128
constexpr int NST = 4;
static_assert(16 % NST == 0, "stage index must follow the sub-item index");
static_assert(SUB_B % 128 == 0 && STAGE % 128 == 0 && SUB_B >= (TH + 2) * SROW && STAGE >= SUB_B + (TH + 2) * SROW, "stage layout");
static_assert(LEAD_W + PIX_W * (SPLIT_TX - 1) + 3 * PIX_W <= BOX_W && LEAD_W + PIX_W * SPLIT_TX >= B_WORD0 &&
              LEAD_W + PIX_W * (TW - 1) + 3 * PIX_W <= B_WORD0 + BOX_W && (TW * PIX_W) % 4 == 0 && (PIX_W + LEAD_W) % 4 == 0 &&
              B_WORD0 % 32 == 0 && BOX_W % 4 == 0, "box coverage, 16-byte aligned box starts");
constexpr int TILES_X = 256 / TW, TILES_PER_IMG = TILES_X * (256 / TH);

struct HeadSumP {
  const float* bias;    // [2][2]
  void* out;
  int num_items, out_layout;
};

__global__ void __launch_bounds__(NTHREADS, 3)
head_tapsum_kernel(const __grid_constant__ CUtensorMap mapO, const __grid_constant__ CUtensorMap mapF, const HeadSumP p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  __shared__ uint64_t full[NST];
  const int tid = threadIdx.x, ty = tid / TW, tx = tid % TW;
  const int n_my = p.num_items > (int)blockIdx.x ? (p.num_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int total = n_my * 16;
  if (tid == 0) {
    prefetch_tmap(&mapO);
    prefetch_tmap(&mapF);
    for (int i = 0; i < NST; ++i) mbar_init(&full[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();

  // sub-item s = 16 * (local item) + 2 * waypoint + head; executed by thread 0 only.  A macro, not a lambda: the descriptor
  // operand of the TMA instruction must be the kernel parameter itself (an out-of-line closure holding references to the
  // maps faulted with "illegal instruction").
#define SJ_TAP_LOAD(S)                                                                                  \
  do {                                                                                                  \
    const int s_ = (S);                                                                                 \
    if (s_ < total) {                                                                                   \
      const int item_ = blockIdx.x + (s_ >> 4) * gridDim.x, th_ = s_ & 15;                              \
      const int b_ = item_ / TILES_PER_IMG, tr_ = item_ % TILES_PER_IMG;                                \
      const int c0_ = ((tr_ % TILES_X) * TW - 1) * PIX_W - LEAD_W, c1_ = (tr_ / TILES_X) * TH - 1, c2_ = b_ * 8 + (th_ >> 1); \
      uint8_t* dst_ = smem + (s_ % NST) * STAGE;                                                        \
      mbar_expect_tx(&full[s_ % NST], 2 * (TH + 2) * SROW);                                             \
      if (th_ & 1) {                                                                                    \
        tma_load_3d(dst_ + SUB_A, &mapF, &full[s_ % NST], c0_, c1_, c2_);                               \
        tma_load_3d(dst_ + SUB_B, &mapF, &full[s_ % NST], c0_ + B_WORD0, c1_, c2_);                     \
      } else {                                                                                          \
        tma_load_3d(dst_ + SUB_A, &mapO, &full[s_ % NST], c0_, c1_, c2_);                               \
        tma_load_3d(dst_ + SUB_B, &mapO, &full[s_ % NST], c0_ + B_WORD0, c1_, c2_);                     \
      }                                                                                                 \
    }                                                                                                   \
  } while (0)
  if (tid == 0)
    for (int s0 = 0; s0 < NST; ++s0) SJ_TAP_LOAD(s0);

  const float b00 = p.bias[0], b01 = p.bias[1], b10 = p.bias[2], b11 = p.bias[3];
  // this thread's pixel (ty, tx) inside a stage: halo-row word 9 * tx, in box A or box B
  const int my_off = (tx < SPLIT_TX ? SUB_A + (LEAD_W + tx * PIX_W) * 4 : SUB_B + (LEAD_W + tx * PIX_W - B_WORD0) * 4) + ty * SROW;
  int s = 0;
  for (int it = 0; it < n_my; ++it) {
    float v[32];
#pragma unroll
    for (int th = 0; th < 16; ++th, ++s) {  // s % NST == th % NST (16 is a multiple of NST): stage offsets are constants
      mbar_wait(&full[th % NST], (uint32_t)(s / NST) & 1);
      const uint8_t* st = smem + (th % NST) * STAGE + my_off;
      float a0 = (th & 1) ? b10 : b00, a1 = (th & 1) ? b11 : b01;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const __half2 h = *reinterpret_cast<const __half2*>(st + (tap / 3) * SROW + (tap % 3) * PIX_W * 4 + tap * 4);
        const float2 f = __half22float2(h);
        a0 += f.x;
        a1 += f.y;
      }
      v[2 * th] = a0;
      v[2 * th + 1] = a1;
      // Stage hand-back: the generic-proxy reads above are ordered before the async-proxy (TMA) write that reuses the
      // stage by a proxy fence + the block barrier, and the stage refilled here is the one read in the PREVIOUS sub-item
      // (three loads in flight, one sub-item of slack).  The first version refilled the stage just read, without the fence:
      // bit-exact in isolation, but whole-model runs showed rare corrupted pixel groups at batch 16.
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0 && s >= 1) SJ_TAP_LOAD(s - 1 + NST);
    }
    const int item = blockIdx.x + it * gridDim.x;
    const int b = item / TILES_PER_IMG, tr = item % TILES_PER_IMG;
    const int y = (tr / TILES_X) * TH + ty, x = (tr % TILES_X) * TW + tx;
    float* outf = reinterpret_cast<float*>(p.out);
    if (p.out_layout == 2) {  // submission bytes, 32 per pixel
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
#undef SJ_TAP_LOAD
}

}  // namespace

// z_occ, z_flow: fp16 [B*8,256,256,18] (tc_upconv4h); bias fp32 [2][2]; out per out_layout (see out_conv)
void head_tapsum(Ctx& c, const void* z_occ, const void* z_flow, const float* bias, int B, int out_layout, void* out) {
  if (!c.ok() || c.dry) return;
  HeadSumP p{};
  p.bias = bias; p.out = out;
  p.num_items = B * TILES_PER_IMG; p.out_layout = out_layout;
  CUtensorMap mapO, mapF;
  uint64_t dz[3] = {(uint64_t)ROW_W, 256, (uint64_t)B * 8};
  uint64_t sz[2] = {(uint64_t)ROW_W * 4, (uint64_t)256 * ROW_W * 4};
  uint32_t bz[3] = {BOX_W, TH + 2, 1};
  if (!encode_tmap(&mapO, z_occ, 3, dz, sz, bz, 0, 4) || !encode_tmap(&mapF, z_flow, 3, dz, sz, bz, 0, 4)) {
    snprintf(tls().cuda_err, sizeof(tls().cuda_err), "cuTensorMapEncodeTiled failed (head_tapsum)");
    c.fail(SJ_ECUDA);
    return;
  }
  const size_t smem = NST * STAGE + 128;
  if (!SJ_SMEM_LIMIT_OK(head_tapsum_kernel, (int)(NST * STAGE + 128))) { c.fail(SJ_ECUDA); return; }
  const int grid = p.num_items < 3 * num_sms() ? p.num_items : 3 * num_sms();
  SJ_LAUNCH(c, "head_tapsum", head_tapsum_kernel, grid, NTHREADS, smem, mapO, mapF, p);
}

}  // namespace sj
