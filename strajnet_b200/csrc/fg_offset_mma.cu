// FG-MSA offset network on warp-level tensor cores (bf16 path): grouped 3x3 conv (8 groups x 48 -> 48) + bias ->
// LayerNorm(384, eps 1e-3) -> tanh-GELU -> per-group 48 -> 2 projection -> 8*tanh -> off, pos
// (FG_MSA.py:84-92, :114-117, :134).  1.4 GFLOP per batch-16 step, but the CUDA-core kernel re-reads the 663 KB of
// conv weights for every 8 pixels; here one block owns two image rows (32 pixels = two m16 tiles), warp g owns
// group g (N = 48, K = 9 taps x 48 channels) and runs it as 324 mma.sync.m16n8k16 with the activation halo
// (4 x 18 pixels x 384 channels) staged once in shared memory and read with ldmatrix; the weights come as bf16
// B fragments packed on the host side (one coalesced 16-byte load per lane for four MMAs), or, without that copy, as
// fp32 converted on the fly (each fragment serves both pixel tiles).
#include "kernels.h"
#include "mma_sync.cuh"

namespace sj {
namespace {

constexpr int PS = 384 + 8;  // smem pixel stride in elements (784 B: conflict-free ldmatrix rows)
constexpr int HALO_PIX = 4 * 18;

__global__ void __launch_bounds__(256) fg_offset_mma_kernel(const bf16* __restrict__ q, int ldq, SjFgmsaW w,
                                                            float* __restrict__ off, float* __restrict__ pos) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* qs = reinterpret_cast<bf16*>(smem_raw);                       // [4][18][PS]
  float* red = reinterpret_cast<float*>(qs + HALO_PIX * PS);          // [8 warps][32 pixels]
  float* stat = red + 8 * 32;                                         // [32 pixels][2]
  const int b = blockIdx.x >> 3, i0 = (blockIdx.x & 7) * 2;           // image rows i0, i0+1
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g8 = lane >> 2, t = lane & 3;

  // ---- halo tile: rows i0-1 .. i0+2, columns -1 .. 16, zero outside the image (SAME padding) ----
  for (int i = threadIdx.x; i < HALO_PIX * 48; i += 256) {
    const int pix = i / 48, c = (i % 48) * 8;
    const int yy = i0 - 1 + pix / 18, xx = pix % 18 - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (yy >= 0 && yy < 16 && xx >= 0 && xx < 16)
      v = *reinterpret_cast<const uint4*>(q + ((long long)b * 256 + yy * 16 + xx) * ldq + c);
    *reinterpret_cast<uint4*>(qs + pix * PS + c) = v;
  }
  __syncthreads();

  // ---- grouped conv: acc[m-tile][n-tile][4], warp = group ----
  const int grp = warp;
  float acc[2][6][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 6; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
  const uint32_t qs_base = (uint32_t)__cvta_generic_to_shared(qs);
  // ldmatrix lane addressing for the A operand: tiles (pixels +0, k +0), (pixels +8, k +0), (pixels +0, k +8), (pixels +8, k +8)
  const int a_pix = (lane & 7) + ((lane >> 3) & 1) * 8, a_k = (lane >> 4) * 8;
  if (w.conv0_w_tc) {
    // bf16 weights already in B-fragment order (weights.py: fg_conv_fragments): one 16-byte load per lane feeds two
    // n-tiles x two pixel tiles = four MMAs
    const uint4* wf = reinterpret_cast<const uint4*>(w.conv0_w_tc) + (long long)grp * (9 * 3 * 3 * 32) + lane;
    // 27 (tap, k-step) iterations, fully unrolled with the next iteration's fragments in flight under this one's MMAs
    // (each load is an L2 round trip; serialised they were most of the kernel)
    uint4 bw[2][3];
#pragma unroll
    for (int pr = 0; pr < 3; ++pr) bw[0][pr] = __ldg(wf + pr * 32);
#pragma unroll
    for (int it = 0; it < 27; ++it) {
      const int tap = it / 3, ks = it % 3, dy = tap / 3, dx = tap % 3;
      if (it + 1 < 27) {
#pragma unroll
        for (int pr = 0; pr < 3; ++pr) bw[(it + 1) & 1][pr] = __ldg(wf + ((it + 1) * 3 + pr) * 32);
      }
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
        ldsm_x4(a[mt], qs_base + (uint32_t)((((mt + dy) * 18 + a_pix + dx) * PS + grp * 48 + ks * 16 + a_k) * 2));
#pragma unroll
      for (int pr = 0; pr < 3; ++pr) {
        const uint4 b = bw[it & 1][pr];
        mma_bf16(acc[0][2 * pr], a[0], b.x, b.y);
        mma_bf16(acc[1][2 * pr], a[1], b.x, b.y);
        mma_bf16(acc[0][2 * pr + 1], a[0], b.z, b.w);
        mma_bf16(acc[1][2 * pr + 1], a[1], b.z, b.w);
      }
    }
  } else {
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3, dx = tap % 3;
    const float* wt = w.conv0_w + (long long)tap * 48 * 384 + grp * 48;  // [cc][n], n contiguous (stride 384)
#pragma unroll
    for (int ks = 0; ks < 3; ++ks) {
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
        ldsm_x4(a[mt], qs_base + (uint32_t)((((mt + dy) * 18 + a_pix + dx) * PS + grp * 48 + ks * 16 + a_k) * 2));
#pragma unroll
      for (int nt = 0; nt < 6; ++nt) {
        const float* wk = wt + (long long)(ks * 16 + 2 * t) * 384 + nt * 8 + g8;
        const uint32_t b0 = pack_bf16(wk[0], wk[384]), b1 = pack_bf16(wk[8 * 384], wk[9 * 384]);
        mma_bf16(acc[0][nt], a[0], b0, b1);
        mma_bf16(acc[1][nt], a[1], b0, b1);
      }
    }
  }
  }
  // thread holds, for pixels (mt, g8) and (mt, g8 + 8), channels grp*48 + nt*8 + 2t + {0,1}
  float s4[4] = {0.f, 0.f, 0.f, 0.f};  // pixel index p = mt*2 + half -> pixel mt*16 + g8 + 8*half
#pragma unroll
  for (int nt = 0; nt < 6; ++nt) {
    const int n = grp * 48 + nt * 8 + 2 * t;
    const float b0 = w.conv0_b[n], b1 = w.conv0_b[n + 1];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      acc[mt][nt][0] += b0; acc[mt][nt][1] += b1; acc[mt][nt][2] += b0; acc[mt][nt][3] += b1;
      s4[mt * 2] += acc[mt][nt][0] + acc[mt][nt][1];
      s4[mt * 2 + 1] += acc[mt][nt][2] + acc[mt][nt][3];
    }
  }
  // ---- LayerNorm over the 384 channels of a pixel (8 warps x 4 lanes each hold a share): two passes ----
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s4[j] += __shfl_xor_sync(0xffffffffu, s4[j], 1);
    s4[j] += __shfl_xor_sync(0xffffffffu, s4[j], 2);
  }
  if (t == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) red[warp * 32 + (j >> 1) * 16 + g8 + 8 * (j & 1)] = s4[j];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    float m = 0.f;
    for (int k = 0; k < 8; ++k) m += red[k * 32 + threadIdx.x];
    stat[2 * threadIdx.x] = m / 384.f;
  }
  __syncthreads();
  float mu[4], q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) mu[j] = stat[2 * ((j >> 1) * 16 + g8 + 8 * (j & 1))];
#pragma unroll
  for (int nt = 0; nt < 6; ++nt)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float d;
      d = acc[mt][nt][0] - mu[mt * 2]; q4[mt * 2] = fmaf(d, d, q4[mt * 2]);
      d = acc[mt][nt][1] - mu[mt * 2]; q4[mt * 2] = fmaf(d, d, q4[mt * 2]);
      d = acc[mt][nt][2] - mu[mt * 2 + 1]; q4[mt * 2 + 1] = fmaf(d, d, q4[mt * 2 + 1]);
      d = acc[mt][nt][3] - mu[mt * 2 + 1]; q4[mt * 2 + 1] = fmaf(d, d, q4[mt * 2 + 1]);
    }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    q4[j] += __shfl_xor_sync(0xffffffffu, q4[j], 1);
    q4[j] += __shfl_xor_sync(0xffffffffu, q4[j], 2);
  }
  if (t == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) red[warp * 32 + (j >> 1) * 16 + g8 + 8 * (j & 1)] = q4[j];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = 0.f;
    for (int k = 0; k < 8; ++k) v += red[k * 32 + threadIdx.x];
    stat[2 * threadIdx.x + 1] = rsqrtf(v / 384.f + 1e-3f);
  }
  __syncthreads();
  // ---- GELU(LN) and the per-group 48 -> 2 projection ----
  float o4[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) o4[j][0] = o4[j][1] = 0.f;
#pragma unroll
  for (int nt = 0; nt < 6; ++nt) {
    const int cc = nt * 8 + 2 * t, n = grp * 48 + cc;
    const float ga0 = w.conv_norm.g[n], ga1 = w.conv_norm.g[n + 1], be0 = w.conv_norm.b[n], be1 = w.conv_norm.b[n + 1];
    const float p00 = w.offproj_w[cc * 2], p01 = w.offproj_w[cc * 2 + 1];
    const float p10 = w.offproj_w[cc * 2 + 2], p11 = w.offproj_w[cc * 2 + 3];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int j = mt * 2 + half;
        const float rs = stat[2 * (mt * 16 + g8 + 8 * half) + 1];
        const float u0 = gelu_tanh((acc[mt][nt][2 * half] - mu[j]) * rs * ga0 + be0);
        const float u1 = gelu_tanh((acc[mt][nt][2 * half + 1] - mu[j]) * rs * ga1 + be1);
        o4[j][0] = fmaf(u0, p00, fmaf(u1, p10, o4[j][0]));
        o4[j][1] = fmaf(u0, p01, fmaf(u1, p11, o4[j][1]));
      }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      o4[j][o] += __shfl_xor_sync(0xffffffffu, o4[j][o], 1);
      o4[j][o] += __shfl_xor_sync(0xffffffffu, o4[j][o], 2);
    }
  if (t == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + (j >> 1), jx = g8 + 8 * (j & 1), pix = i * 16 + jx;
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        const float a = tanhf(o4[j][o]) * 8.0f;  // offset_range = (Hk/2, Wk/2) = (8, 8), FG_MSA.py:115-117
        const long long idx = (((long long)b * 8 + grp) * 256 + pix) * 2 + o;
        off[idx] = a;
        pos[idx] = a + (o == 0 ? (float)jx : (float)i);  // tf.meshgrid 'xy': ref[i,j] = (j, i)
      }
    }
  }
}

}  // namespace

bool fg_offset_mma(Ctx& c, const void* q, int ldq, const SjFgmsaW* w, int B, float* off, float* pos) {
  if (c.dtype != SJ_BF16 || ldq % 8 || (reinterpret_cast<uintptr_t>(q) & 15)) return false;
  const size_t smem = (size_t)HALO_PIX * PS * 2 + (8 * 32 + 64) * 4;
  if (!SJ_SMEM_LIMIT_OK((fg_offset_mma_kernel), (int)smem)) {
    c.fail(SJ_ECUDA);
    return true;
  }
  SJ_LAUNCH(c, "fg_offset_mma", fg_offset_mma_kernel, B * 8, 256, smem, (const bf16*)q, ldq, *w, off, pos);
  return true;
}

}  // namespace sj
