// Second half of the fused decoder head (K10): the 9-tap shifted sum over the projected columns written by
// tc_upconv4h_kernel, for both heads, plus the [B,8,256,256,4] -> [B,256,256,32] transpose (modules.py:767-770, :838)
// and optionally the submission quantisation (inference.py:124-136, :160-182):
//   out[b, y, x, t*4 + head*2 + o] = bias[head][o] + sum_{dy,dx} Z_head[b*8+t, y+dy-1, x+dx-1, (dy*3+dx)*2 + o]
// with zero outside the image (SAME padding of the 3x3 convolution).
//
// HBM-bound by construction: 2 x 36 B read per (waypoint, pixel), 4 B (fp32) or 1 B (quantised) written per output
// element; no tensor cores (9 adds per output).  Work item = (sample, 8 x 32 pixel tile); its 16 (waypoint, head) halo
// tiles stream through a 4-stage cp.async ring (rows are contiguous 18-channel fp16 pixels, copied as 16-byte pieces
// with zero fill outside the image); each thread owns one pixel, reads its nine fp16 pairs per sub-item conflict-free
// (pixel stride 9 words) and finally writes its 32 output channels as one 128-byte line.
#include <cuda_fp16.h>

#include "kernels.h"

namespace sj {
namespace {

constexpr int TW = 32, TH = 8, NTHREADS = TW * TH;
constexpr int ZCH = 18, PIX_B = ZCH * 2;               // 36 bytes per pixel
constexpr int ROW_B = 256 * PIX_B;                     // 9216 bytes per image row
constexpr int LEAD = 12;                               // the halo pixel x0-1 starts 12 bytes into the 16-byte aligned copy
constexpr int ROW_CHUNKS = ((TW + 2) * PIX_B + LEAD + 15) / 16;  // 78
constexpr int SROW = ROW_CHUNKS * 16;                  // 1248 bytes per staged row
constexpr int STAGE = (TH + 2) * SROW;                 // 12480
constexpr int NST = 4;
static_assert(16 % NST == 0, "stage index must follow the sub-item index");
constexpr int TILES_X = 256 / TW, TILES_PER_IMG = TILES_X * (256 / TH);

struct HeadSumP {
  const uint8_t* z[2];  // fp16 [B*8,256,256,18]: occupancy head, flow head
  const float* bias;    // [2][2]
  void* out;
  int num_items, out_layout;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 4) head_tapsum_kernel(const HeadSumP p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int tid = threadIdx.x, ty = tid / TW, tx = tid % TW;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
  const int n_my = p.num_items > (int)blockIdx.x ? (p.num_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int total = n_my * 16;
  pdl_wait();
  pdl_trigger();

  // Loader: the 16-byte pieces of a halo tile are dealt to the threads once (piece tid + 256 k -> row, column), so that a
  // sub-item costs each thread four address adds and four cp.async; validity (image border) changes only with the item.
  constexpr int NPIECE = (TH + 2) * ROW_CHUNKS, KMAX = (NPIECE + NTHREADS - 1) / NTHREADS;
  int p_soff[KMAX], p_row[KMAX], p_col[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int c = tid + k * NTHREADS;
    p_row[k] = c < NPIECE ? c / ROW_CHUNKS : -100000;  // never valid
    p_col[k] = (c % ROW_CHUNKS) * 16;
    p_soff[k] = (c / ROW_CHUNKS) * SROW + p_col[k];
  }
  int l_item = -1;
  int l_goff[KMAX];  // byte offset inside one [256,256,18] image
  unsigned l_ok = 0;
  long long l_img0 = 0;
  auto load = [&](int s) {
    if (s < total) {
      const int item = blockIdx.x + (s >> 4) * gridDim.x, th = s & 15;
      if (item != l_item) {  // new tile: origin, validity of this thread's pieces
        l_item = item;
        const int b = item / TILES_PER_IMG, tr = item % TILES_PER_IMG;
        const int y0 = (tr / TILES_X) * TH, x0 = (tr % TILES_X) * TW;
        const int col0 = x0 * PIX_B - PIX_B - LEAD;  // 16-byte aligned (x0 is a multiple of 32)
        l_img0 = (long long)b * 8 * 256 * ROW_B;
        l_ok = 0;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          const int y = y0 - 1 + p_row[k], off = col0 + p_col[k];
          const bool ok = y >= 0 && y < 256 && off >= 0 && off < ROW_B;
          l_ok |= (ok ? 1u : 0u) << k;
          l_goff[k] = y * ROW_B + off;
        }
      }
      const uint8_t* img = p.z[th & 1] + l_img0 + (long long)(th >> 1) * 256 * ROW_B;
      const uint32_t dst0 = smem_base + (s % NST) * STAGE;
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (p_row[k] >= 0) {
          const bool ok = (l_ok >> k) & 1;
          cp_async16(dst0 + p_soff[k], ok ? img + l_goff[k] : p.z[0], ok ? 16 : 0);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  for (int s = 0; s < NST - 1; ++s) load(s);
  const float b00 = p.bias[0], b01 = p.bias[1], b10 = p.bias[2], b11 = p.bias[3];
  int s = 0;
  for (int it = 0; it < n_my; ++it) {
    float v[32];
#pragma unroll
    for (int th = 0; th < 16; ++th, ++s) {  // s % NST == th % NST (16 is a multiple of NST): stage offsets are constants
      asm volatile("cp.async.wait_group %0;" ::"n"(NST - 2) : "memory");
      __syncthreads();
      load(s + NST - 1);  // refills the stage every thread finished reading before the barrier above
      const uint8_t* st = smem + (th % NST) * STAGE + LEAD;
      float a0 = (th & 1) ? b10 : b00, a1 = (th & 1) ? b11 : b01;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const __half2 h = *reinterpret_cast<const __half2*>(st + (ty + tap / 3) * SROW + (tx + tap % 3) * PIX_B + tap * 4);
        const float2 f = __half22float2(h);
        a0 += f.x;
        a1 += f.y;
      }
      v[2 * th] = a0;
      v[2 * th + 1] = a1;
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
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace

// z_occ, z_flow: fp16 [B*8,256,256,18] (tc_upconv4h); bias fp32 [2][2]; out per out_layout (see out_conv)
void head_tapsum(Ctx& c, const void* z_occ, const void* z_flow, const float* bias, int B, int out_layout, void* out) {
  if (!c.ok() || c.dry) return;
  HeadSumP p{};
  p.z[0] = (const uint8_t*)z_occ; p.z[1] = (const uint8_t*)z_flow; p.bias = bias; p.out = out;
  p.num_items = B * TILES_PER_IMG; p.out_layout = out_layout;
  const size_t smem = NST * STAGE;
  if (!SJ_SMEM_LIMIT_OK(head_tapsum_kernel, (int)(NST * STAGE))) { c.fail(SJ_ECUDA); return; }
  const int grid = p.num_items < 4 * num_sms() ? p.num_items : 4 * num_sms();
  SJ_LAUNCH(c, "head_tapsum", head_tapsum_kernel, grid, NTHREADS, smem, p);
}

}  // namespace sj
