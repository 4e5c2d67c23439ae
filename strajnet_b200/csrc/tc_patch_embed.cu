// Fused patch embedding on the tensor cores (K3, bf16 mode):
//   y[token] = LN_f( LN_0(conv4x4s4(img_0)[token] + b_0) [+ LN_1(conv4x4s4(img_1)[token'] + b_1)] ),  E = 96, eps 1e-5
// (PatchEmbed.call modules.py:437-446, the vec + map sum and all_patch_norm :580-587 / :602, patch_embed_flow + flow_norm
// :576-577), plus the LayerNorm statistics of y for norm1 of the first Swin block.
//
// As im2col -> tcgen05 GEMM -> combine this stage was 8 launches and 168 us of the batch-16 step (80 us now) for 0.3 GFLOP: the
// im2col matrix (25 MB for the occupancy raster) and the conv outputs made a round trip through HBM each.  Here one CTA
// turns 128 consecutive tokens into finished tokens:
//   * all 8 warps read the tile's image rows with coalesced 16-byte loads (the 4 x Cin x es source elements of a token
//     and a kernel row are contiguous, and so are the tokens of a token row; a thread keeps one unit of all 8 (token row,
//     kernel row) segments in flight), convert to bf16 and scatter them into the K-major SWIZZLE_128B A tile(s) in shared
//     memory ("im2col in shared memory"; element stride es = 2 picks plane 0 of the [.., 11, 2] raster);
//   * one elected lane issues the K = 16*Cin (zero-padded to 64) MMAs against the resident projection weights
//     (36 KB + 12 KB), one TMEM accumulator per input;
//   * the accumulators are staged through shared memory (over the dead A tile) so that a warp owns a token: bias, the
//     per-input LayerNorms, the sum, the final LayerNorm and the token's statistics are shuffle reductions over 3 channels
//     per lane with the per-channel constants in registers, and the stores are row-contiguous.  In 512-input mode the map
//     covers only the centre 64 x 64 tokens (modules.py:582-585): tokens outside get no map term.
// Two CTAs per SM (112 KB of shared memory, 256 TMEM columns each) overlap one tile's loads with the other's math.
#include <cstdio>

#include "kernels.h"
#include "tc_common.cuh"

namespace sj {
namespace {

using namespace tc;

constexpr int E = 96, BM = 128, NTHREADS = 256;
constexpr int A_CHUNK = BM * 128;        // 16 KB: 128 tokens x 64 K elements (SWIZZLE_128B)
constexpr int W_CHUNK = E * 128;         // 12 KB: 96 output channels x 64 K elements
constexpr int MAXC0 = 3, MAXC1 = 1;      // K chunks of input 0 (Cin <= 12) and input 1 (Cin <= 4)
constexpr int OFF_A0 = 0;
constexpr int OFF_A1 = OFF_A0 + MAXC0 * A_CHUNK;
constexpr int OFF_W0 = OFF_A1 + MAXC1 * A_CHUNK;
constexpr int OFF_W1 = OFF_W0 + MAXC0 * W_CHUNK;
constexpr int OFF_BAR = OFF_W1 + MAXC1 * W_CHUNK;
constexpr int SMEM_BYTES = OFF_BAR + 64;
static_assert(OFF_W0 % 1024 == 0 && OFF_W1 % 1024 == 0, "tile alignment");
// two CTAs per SM: 2 x (112 KB + 64 B + 1 KB reserved) <= 228 KB.  There is no room for an alignment slack, so the kernel
// has no static shared memory (the dynamic window then starts at offset 0 of the CTA's shared memory, 1024-aligned) and
// traps if that ever stops being true.
static_assert(2 * (SMEM_BYTES + 1024) <= 228 * 1024, "two CTAs per SM");

struct PeIn {
  const void* img;
  int itype, S, Cin, es, P, pad, nkc;  // P = S / 4 tokens per row; pad: this input covers tokens [pad, pad + P) of the grid
  int run, span_u;                     // source elements per (token, kernel row); 16-byte units per (token row, kernel row)
  long long row_bytes;                 // bytes of one image row
  uint32_t m_run;                      // magic multiplier: e / (run / 4) (fp32) or e / run (bytes)
  const float* bias;
  const float* g;
  const float* b;
};
struct PeFusedP {
  PeIn in[2];
  int n_in, B, P, num_tiles;
  const float* gf;
  const float* bf;
  bf16* y;
  float* st_mean;
  float* st_rstd;
};

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// byte address of element (row r, K index k) of a [128 x 64-element chunks] SWIZZLE_128B K-major tile
__device__ __forceinline__ uint8_t* a_addr(uint8_t* tile, int r, int k) {
  const int ch = k >> 6, w = k & 63;
  return tile + ch * A_CHUNK + r * 128 + (((w >> 3) ^ (r & 7)) << 4) + (w & 7) * 2;
}
__device__ __forceinline__ uint32_t bf2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <int ITYPE>
__device__ __forceinline__ float cvt_byte(uint32_t byte) {
  return ITYPE == IN_U8 ? (byte != 0 ? 1.0f : 0.0f) : (float)(int8_t)byte / 256.0f;
}

// One input's share of a tile.  The source is nseg = 4 x token rows contiguous segments (segment = 4 * token row + kernel
// row, `row_bytes` apart) of span_u 16-byte units each.  A thread owns unit positions e = tid, tid + 256, ...: where a
// position lands in the tile (token, offset inside the token's run) is the same in every segment, so it is worked out
// once and the thread then loads that position of ALL segments (up to 8 loads in flight) before it converts and scatters
// them; segment-dependent terms (row block, kernel row) are compile-time in the unrolled loops.
template <int ITYPE, int ES>
__device__ __forceinline__ void load_tile(uint8_t* tileA, const uint8_t* img_b, long long row_bytes, int nseg, int span_u,
                                          uint32_t m_run, int run, int Cin, int pad, int P, int tid) {
  constexpr int G = ITYPE == IN_F32 ? 1 : 4;  // 4-element groups per 16-byte unit
  for (int e = tid; e < span_u; e += NTHREADS) {
    int rowoff[G], kq[G];  // byte offset of the token's tile row within its token row block; K offset within the kernel row
    {
      int tk, j;
      if (ITYPE == IN_F32) {
        tk = __umulhi(e, m_run);  // m_run divides by run / 4 here
        j = (e - tk * (run >> 2)) << 2;
      } else {
        tk = __umulhi(e << 4, m_run);
        j = (e << 4) - tk * run;
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        rowoff[g] = (pad + tk) * 128;
        kq[g] = ES == 2 ? j >> 1 : j;
        j += 4;
        if (j >= run) { j = 0; ++tk; }
      }
    }
    const uint8_t* src = img_b + (long long)e * 16;
    uint4 v[8];
#pragma unroll
    for (int sg = 0; sg < 8; ++sg)
      if (sg < nseg) v[sg] = __ldg(reinterpret_cast<const uint4*>(src + sg * row_bytes));
#pragma unroll
    for (int sg = 0; sg < 8; ++sg) {
      if (sg < nseg) {
        const int rbase = (sg >> 2) * P * 128, kb = (sg & 3) * 4 * Cin;
        const uint32_t wv[4] = {v[sg].x, v[sg].y, v[sg].z, v[sg].w};
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const int k = kb + kq[g], ro = rbase + rowoff[g];
          uint8_t* dst = tileA + (k >> 6) * A_CHUNK + ro + ((((k >> 3) ^ (ro >> 7)) & 7) << 4) + (k & 7) * 2;
          float x0, x1, x2, x3;
          if (ITYPE == IN_F32) {
            x0 = __uint_as_float(wv[0]); x1 = __uint_as_float(wv[1]); x2 = __uint_as_float(wv[2]); x3 = __uint_as_float(wv[3]);
          } else {
            const uint32_t w = wv[g];
            x0 = cvt_byte<ITYPE>(w & 0xff); x1 = cvt_byte<ITYPE>((w >> 8) & 0xff);
            x2 = cvt_byte<ITYPE>((w >> 16) & 0xff); x3 = cvt_byte<ITYPE>(w >> 24);
          }
          // element stride 1 keeps all four (K indices k .. k + 3), element stride 2 keeps x0 and x2 (K indices k, k + 1)
          if (ES == 2)
            *reinterpret_cast<uint32_t*>(dst) = bf2(x0, x2);
          else
            *reinterpret_cast<uint2*>(dst) = make_uint2(bf2(x0, x1), bf2(x2, x3));
        }
      }
    }
  }
}

// ---- epilogue helpers ----
// accumulator (lane = token, 96 fp32 columns at `tm`) -> staging tile: warp w moves TMEM lane quarter w % 4, columns
// 48 * (w / 4) .. + 47.  Row r's 16-byte chunk c4 is stored at chunk c4 ^ (r & 7): conflict-free both ways.
__device__ __forceinline__ void stage_acc(float* stg, uint32_t tm, int warp, int lane) {
  const int q = warp & 3, h = warp >> 2, r = q * 32 + lane;
  const uint32_t t_addr = tm + ((uint32_t)(q * 32) << 16) + h * 48;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float v[16];
    tmem_ld16(t_addr + 16 * j, v);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int c4 = h * 12 + j * 4 + m;
      *reinterpret_cast<float4*>(stg + r * E + ((c4 ^ (r & 7)) << 2)) = make_float4(v[4 * m], v[4 * m + 1], v[4 * m + 2], v[4 * m + 3]);
    }
  }
}
// lane l's channels of token row r: 2l, 2l+1, 64+l
__device__ __forceinline__ void fetch_row(const float* stg, int r, int lane, float (&x)[3]) {
  const float2 a = *reinterpret_cast<const float2*>(stg + r * E + (((lane >> 1) ^ (r & 7)) << 2) + (lane & 1) * 2);
  x[0] = a.x;
  x[1] = a.y;
  x[2] = stg[r * E + (((16 + (lane >> 2)) ^ (r & 7)) << 2) + (lane & 3)];
}
// Sums over the warp of 8 independent values per lane in 9 shuffles (recursive halving): returns the total of value
// index 4 * bit4(lane) + 2 * bit3(lane) + bit2(lane)
__device__ __forceinline__ float reduce_scatter8(const float (&v)[8], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  float a4[4], a2[2];
#pragma unroll
  for (int j = 0; j < 4; ++j) a4[j] = (b4 ? v[j + 4] : v[j]) + __shfl_xor_sync(0xffffffffu, b4 ? v[j] : v[j + 4], 16);
#pragma unroll
  for (int j = 0; j < 2; ++j) a2[j] = (b3 ? a4[j + 2] : a4[j]) + __shfl_xor_sync(0xffffffffu, b3 ? a4[j] : a4[j + 2], 8);
  float a = (b2 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? a2[0] : a2[1], 4);
  a += __shfl_xor_sync(0xffffffffu, a, 2);
  a += __shfl_xor_sync(0xffffffffu, a, 1);
  return a;
}
// ... and every lane gets all 8 totals (8 more shuffles)
__device__ __forceinline__ void allreduce8(float (&v)[8], int lane) {
  const float a = reduce_scatter8(v, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __shfl_sync(0xffffffffu, a, ((i >> 2) & 1) << 4 | ((i >> 1) & 1) << 3 | (i & 1) << 2);
}
// LayerNorm (biased variance, two passes, eps 1e-5) of 8 token rows held 3 channels per lane
__device__ __forceinline__ void ln8(float (&x)[8][3], const float (&g)[3], const float (&b)[3], int lane) {
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = x[i][0] + x[i][1] + x[i][2];
  allreduce8(s, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float mu = s[i] * (1.0f / E);
#pragma unroll
    for (int c = 0; c < 3; ++c) x[i][c] -= mu;
    s[i] = fmaf(x[i][0], x[i][0], fmaf(x[i][1], x[i][1], x[i][2] * x[i][2]));
  }
  allreduce8(s, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float rs = rsqrtf(s[i] * (1.0f / E) + 1e-5f);
#pragma unroll
    for (int c = 0; c < 3; ++c) x[i][c] = fmaf(x[i][c] * rs, g[c], b[c]);
  }
}

// T0 / ES0: input type and element stride of input 0; T1: input type of input 1 (element stride 1), -1 = no second input
template <int T0, int ES0, int T1>
__global__ void __launch_bounds__(NTHREADS, 2)
tc_patch_embed_kernel(const __grid_constant__ CUtensorMap mapW0, const __grid_constant__ CUtensorMap mapW1,
                      const PeFusedP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* wfull = bars;
  uint64_t* dfull = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid % 32;

  if (tid == 0) {
    prefetch_tmap(&mapW0);
    if (T1 >= 0) prefetch_tmap(&mapW1);
    mbar_init(wfull, 1);
    mbar_init(dfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  // the K padding of the A tiles stays zero for the whole kernel
  for (int i = tid; i < (MAXC0 + MAXC1) * A_CHUNK / 16; i += NTHREADS)
    reinterpret_cast<uint4*>(smem + OFF_A0)[i] = make_uint4(0, 0, 0, 0);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    mbar_expect_tx(wfull, (p.in[0].nkc + (T1 >= 0 ? p.in[1].nkc : 0)) * W_CHUNK);
    for (int c = 0; c < p.in[0].nkc; ++c) tma_load_2d(smem + OFF_W0 + c * W_CHUNK, &mapW0, wfull, c * 64, 0);
    if (T1 >= 0)
      for (int c = 0; c < p.in[1].nkc; ++c) tma_load_2d(smem + OFF_W1 + c * W_CHUNK, &mapW1, wfull, c * 64, 0);
  }
  pdl_wait();
  pdl_trigger();

  // per-channel constants of this lane's channels (2l, 2l+1, 64+l) in the warp-per-token phase
  float k_bias0[3], k_g0[3], k_b0[3], k_bias1[3], k_g1[3], k_b1[3], k_gf[3], k_bf[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int ch = c < 2 ? 2 * lane + c : 64 + lane;
    k_bias0[c] = p.in[0].bias[ch]; k_g0[c] = p.in[0].g[ch]; k_b0[c] = p.in[0].b[ch];
    k_bias1[c] = k_g1[c] = k_b1[c] = 0.f;
    if (T1 >= 0) { k_bias1[c] = p.in[1].bias[ch]; k_g1[c] = p.in[1].g[ch]; k_b1[c] = p.in[1].b[ch]; }
    k_gf[c] = p.gf[ch]; k_bf[c] = p.bf[ch];
  }
  float* stg = reinterpret_cast<float*>(smem + OFF_A0);  // [128][96] fp32 staging tile = exactly the three A0 chunks
  static_assert(BM * E * 4 == MAXC0 * A_CHUNK, "staging tile");

  const int rows_per_tile = BM / p.P > 0 ? BM / p.P : 1;  // token rows per tile (P = 64: 2, P = 128: 1)
  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
    const long long tok0 = (long long)tile * BM;
    const int b = (int)(tok0 / ((long long)p.P * p.P));
    const int pi0 = (int)((tok0 % ((long long)p.P * p.P)) / p.P);
    const bool has1 = T1 >= 0 && !(p.in[1].pad != 0 && (pi0 - p.in[1].pad < 0 || pi0 - p.in[1].pad >= p.in[1].P));
    // ---- im2col in shared memory ----
    // the K padding of A0 (16-byte chunks [2 Cin, 8 nkc) of every row) held staging data: zero it again
    {
      const int z0 = 2 * p.in[0].Cin, z1 = 8 * p.in[0].nkc;
      if (tid < BM)
        for (int kc = z0; kc < z1; ++kc) *reinterpret_cast<uint4*>(a_addr(smem + OFF_A0, tid, kc * 8)) = make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {  // unrolled: p.in[] must be indexed statically (kernel parameter space)
      if (i >= (T1 < 0 ? 1 : 2)) break;
      const PeIn& in = p.in[i];
      if (in.pad != 0 && (pi0 - in.pad < 0 || pi0 - in.pad >= in.P)) continue;  // this token row has no input-i tokens
      uint8_t* tileA = smem + (i == 0 ? OFF_A0 : OFF_A1);
      const uint8_t* img_b = reinterpret_cast<const uint8_t*>(in.img) + ((long long)b * in.S + 4 * (pi0 - in.pad)) * in.row_bytes;
      if (i == 0)
        load_tile<T0, ES0>(tileA, img_b, in.row_bytes, rows_per_tile * 4, in.span_u, in.m_run, in.run, in.Cin, in.pad, p.P, tid);
      else
        load_tile<(T1 < 0 ? 0 : T1), 1>(tileA, img_b, in.row_bytes, rows_per_tile * 4, in.span_u, in.m_run, in.run, in.Cin, in.pad, p.P, tid);
    }
    fence_async_smem();
    __syncthreads();
    // ---- projections ----
    if (warp == 0) {
      if (it == 0) mbar_wait(wfull, 0);
      tc_fence_after();
      if (elect_one()) {
        constexpr uint32_t HI = desc_hi(128, 1024);
        const uint32_t idesc = make_idesc_bf16(BM, E);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (i >= (T1 < 0 ? 1 : 2) || (i == 1 && !has1)) break;
          const uint32_t a_lo = desc_lo(smem_u32(smem + (i == 0 ? OFF_A0 : OFF_A1)));
          const uint32_t w_lo = desc_lo(smem_u32(smem + (i == 0 ? OFF_W0 : OFF_W1)));
          for (int c = 0; c < p.in[i].nkc; ++c)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_w(tmem + i * 128, a_lo + ((c * A_CHUNK) >> 4) + 2 * k, HI, w_lo + ((c * W_CHUNK) >> 4) + 2 * k, HI, idesc,
                          (c | k) != 0);
        }
        umma_commit(dfull);
      }
      __syncwarp();
    }
    // ---- bias, LayerNorms, sum, final LayerNorm, statistics ----
    // The accumulators go through a staging tile (fp32 [128][96] over the dead A0 tile, 16-byte chunks XOR-swizzled by
    // row) so that a WARP owns a token: lane l holds channels 2l, 2l+1 and 64+l, the per-channel constants live in
    // registers for the whole kernel, the reductions are shuffles and the stores are row-contiguous.  (Thread = token,
    // 96 channels in registers, took ~800 dependent constant loads per token on 4 of the 8 warps: 20 us per tile.)
    mbar_wait(dfull, it & 1);
    tc_fence_after();
    float x[2][8][3];  // this warp's 16 token rows, in two groups of 8
    stage_acc(stg, tmem, warp, lane);
    tc_fence_before();
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      fetch_row(stg, warp * 16 + i, lane, x[i >> 3][i & 7]);
#pragma unroll
      for (int c = 0; c < 3; ++c) x[i >> 3][i & 7][c] += k_bias0[c];
    }
#pragma unroll
    for (int g = 0; g < 2; ++g) ln8(x[g], k_g0, k_b0, lane);
    if (has1) {
      __syncthreads();  // every warp has fetched its rows of the first accumulator
      tc_fence_after();
      stage_acc(stg, tmem + 128, warp, lane);
      tc_fence_before();
      __syncthreads();
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float y[8][3];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          fetch_row(stg, warp * 16 + 8 * g + i, lane, y[i]);
#pragma unroll
          for (int c = 0; c < 3; ++c) y[i][c] += k_bias1[c];
        }
        ln8(y, k_g1, k_b1, lane);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // does this token lie inside input 1's window of the token grid (its token row does: has1)?
          const int r = warp * 16 + 8 * g + i, pj_1 = r % p.P - p.in[1].pad;
          if (pj_1 >= 0 && pj_1 < p.in[1].P) {
#pragma unroll
            for (int c = 0; c < 3; ++c) x[g][i][c] += y[i][c];
          }
        }
      }
    }
    __syncthreads();  // the staging tile is dead: the next tile's loads may overwrite it (the accumulators were drained before)
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      ln8(x[g], k_gf, k_bf, lane);
      float ss[8], qq[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long tok = tok0 + warp * 16 + 8 * g + i;
        bf16* dst = p.y + tok * E;
        const __nv_bfloat162 h01 = __floats2bfloat162_rn(x[g][i][0], x[g][i][1]);
        const bf16 h2 = __float2bfloat16_rn(x[g][i][2]);
        *reinterpret_cast<__nv_bfloat162*>(dst + 2 * lane) = h01;
        dst[64 + lane] = h2;
        const float r0 = __low2float(h01), r1 = __high2float(h01), r2 = __bfloat162float(h2);  // statistics of the stored values
        ss[i] = r0 + r1 + r2;
        qq[i] = fmaf(r0, r0, fmaf(r1, r1, r2 * r2));
      }
      if (p.st_mean) {
        const float s1 = reduce_scatter8(ss, lane), q1 = reduce_scatter8(qq, lane);
        if ((lane & 3) == 0) {
          const long long tok = tok0 + warp * 16 + 8 * g + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
          const float mu = s1 * (1.0f / E);
          p.st_mean[tok] = mu;
          p.st_rstd[tok] = rsqrtf(fmaxf(q1 * (1.0f / E) - mu * mu, 0.f) + 1e-5f);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

}  // namespace

bool tc_patch_embed_supported(int B, int P, int Cin0, int Cin1, int P1, int pad1) {
  if (P != 64 && P != 128) return false;
  if (16 * Cin0 > 64 * MAXC0 || (Cin1 > 0 && 16 * Cin1 > 64 * MAXC1)) return false;
  if (Cin1 > 0 && (P1 + 2 * pad1 != P || (pad1 != 0 && P != BM))) return false;  // an offset window needs one token row per tile
  return ((long long)B * P * P) % BM == 0;
}

// img0 [B,S0,S0,Cin0(,es0)] (+ img1 [B,S1,S1,Cin1]) -> y bf16 [B,P*P,96] with P = S0/4 (+ norm1 statistics of y);
// pw*.proj.w_tc = [96][Kpad] bf16 (K = ky*4*Cin + kx*Cin + c, zero-padded to a multiple of 64)
void tc_patch_embed(Ctx& c, const void* img0, int itype0, int S0, int Cin0, int es0, const SjPatchEmbedW& pw0,
                    const void* img1, int itype1, int S1, int Cin1, const SjPatchEmbedW* pw1, int pad1, const SjNorm& nf,
                    int B, void* y, float* st_mean, float* st_rstd) {
  if (!c.ok() || c.dry) return;
  const int P = S0 / 4, n_in = img1 ? 2 : 1;
  if (!pw0.proj.w_tc || !pw0.proj.b || (n_in > 1 && (!pw1 || !pw1->proj.w_tc || !pw1->proj.b)) ||
      !tc_patch_embed_supported(B, P, Cin0, n_in > 1 ? Cin1 : 0, S1 / 4, pad1) || (es0 != 1 && es0 != 2) || S0 % 16 != 0 ||
      (n_in > 1 && S1 % 16 != 0)) {  // S % 16: every image row starts 16-byte aligned for the byte rasters too
    c.fail(SJ_EUNSUPPORTED);
    return;
  }
  PeFusedP p{};
  p.n_in = n_in; p.B = B; p.P = P; p.num_tiles = (int)((long long)B * P * P / BM);
  auto fill = [](PeIn& in, const void* img, int itype, int S, int Cin, int es, int pad, const SjPatchEmbedW& pw) {
    in.img = img; in.itype = itype; in.S = S; in.Cin = Cin; in.es = es; in.P = S / 4; in.pad = pad; in.nkc = (16 * Cin + 63) / 64;
    in.run = 4 * Cin * es;
    const int esz = itype == IN_F32 ? 4 : 1, upe = 16 / esz;  // elements per 16-byte unit
    in.span_u = in.P * in.run / upe;
    in.row_bytes = (long long)S * Cin * es * esz;
    auto magic = [](uint32_t d) { return (uint32_t)((1ull << 32) / d + 1); };  // exact for u * d < 2^32, d >= 2
    in.m_run = magic(itype == IN_F32 ? in.run / 4 : in.run);
    in.bias = pw.proj.b; in.g = pw.norm.g; in.b = pw.norm.b;
  };
  fill(p.in[0], img0, itype0, S0, Cin0, es0, 0, pw0);
  if (n_in > 1) fill(p.in[1], img1, itype1, S1, Cin1, 1, pad1, *pw1);
  p.gf = nf.g; p.bf = nf.b; p.y = (bf16*)y; p.st_mean = st_mean; p.st_rstd = st_rstd;
  CUtensorMap mapW0, mapW1;
  auto wmap = [&](CUtensorMap* m, const void* w, int nkc) {
    uint64_t d[2] = {(uint64_t)nkc * 64, (uint64_t)E};
    uint64_t s[1] = {(uint64_t)nkc * 64 * 2};
    uint32_t bx[2] = {64, (uint32_t)E};
    return encode_tmap(m, w, 2, d, s, bx, 128);
  };
  if (!wmap(&mapW0, pw0.proj.w_tc, p.in[0].nkc) || !wmap(&mapW1, n_in > 1 ? pw1->proj.w_tc : pw0.proj.w_tc, n_in > 1 ? p.in[1].nkc : p.in[0].nkc)) {
    snprintf(tls().cuda_err, sizeof(tls().cuda_err), "cuTensorMapEncodeTiled failed (tc_patch_embed)");
    c.fail(SJ_ECUDA);
    return;
  }
  // two persistent CTAs per SM.  (One tile per CTA, 512 CTAs at batch 16, so that later waves stagger the load and epilogue
  // phases: 83 us for the two launches against 80 us; an L2 prefetch of the next tile's rows during the epilogue: no change,
  // the loads run at L2 -> SM bandwidth either way.)
  const int grid = p.num_tiles < 2 * num_sms() ? p.num_tiles : 2 * num_sms();
  bool launched = false;
  const int t1 = n_in > 1 ? itype1 : -1;
  // (the shared-memory opt-in is remembered per call site: one site per instance)
#define SJ_PE_CASE(T0, ES0, T1)                                                                                  \
  if (!launched && itype0 == T0 && es0 == ES0 && t1 == T1) {                                                    \
    launched = true;                                                                                             \
    if (!SJ_SMEM_LIMIT_OK((tc_patch_embed_kernel<T0, ES0, T1>), 227 * 1024)) { c.fail(SJ_ECUDA); return; }       \
    SJ_LAUNCH(c, "tc_patch_embed", (tc_patch_embed_kernel<T0, ES0, T1>), grid, NTHREADS, SMEM_BYTES, mapW0, mapW1, p); \
  }
  SJ_PE_CASE(IN_F32, 1, -1)            // flow
  SJ_PE_CASE(IN_F32, 2, IN_F32)        // [.., 11, 2] fp32 raster + fp32 map
  SJ_PE_CASE(IN_F32, 1, IN_F32)        // vehicle plane alone
  SJ_PE_CASE(IN_U8, 2, IN_I8_DIV256)   // the record's own bytes
  SJ_PE_CASE(IN_U8, 1, IN_I8_DIV256)
  SJ_PE_CASE(IN_F32, 2, IN_I8_DIV256)
  SJ_PE_CASE(IN_F32, 1, IN_I8_DIV256)
  SJ_PE_CASE(IN_U8, 2, IN_F32)
  SJ_PE_CASE(IN_U8, 1, IN_F32)
#undef SJ_PE_CASE
  if (!launched) c.fail(SJ_EUNSUPPORTED);
}

}  // namespace sj
