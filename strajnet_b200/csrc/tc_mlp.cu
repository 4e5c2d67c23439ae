// Fused Swin MLP half on the tensor cores (C = 96, hidden = 384, bf16):
//   y = x + fc2( GELU( fc1( norm2(x) ) ) )                                   (modules.py:260-261 with Mlp.call :41-45)
// as ONE kernel.  As two tcgen05 GEMMs (tc_gemm.cu) this half costs 53 us per block at batch 16, nearly all of it
// epilogue and launch latency around a 50 MB hidden tensor that is written to HBM and read straight back; here the
// hidden activations of a 128-token tile never leave the SM (41 us per block; with an empty epilogue the kernel still
// takes 31 us: at batch 16 there are only 3.5 tiles per SM and each tile is a chain of ~6 MMA <-> epilogue hand-offs,
// and the 144 KB of resident weights leave room for one tile per SM, so the chain latency is what is left):
//   * weights resident in shared memory: norm2-folded W1 [384 x 96] (72 KB) and W2 [96 x 384] (72 KB);
//   * fc1 runs as four N = 96 quarters into four TMEM accumulators (LayerNorm folded: raw tokens feed the MMA, the
//     epilogue applies rstd * (acc - mean * colsum) + bias');
//   * the 8 epilogue warps apply the fold + tanh-GELU and write each quarter as a bf16 A operand (128 x 96, 24 KB,
//     32-unit SWIZZLE_64B chunks) into one of TWO shared-memory buffers; fc2 consumes it as K = 96 and accumulates the
//     four quarters into a fifth TMEM accumulator (96 columns).  With two buffers the fc2 MMAs of quarter q run under
//     the GELU of quarter q+1, and fc1 of the next tile under the output epilogue: the epilogue warps never wait;
//   * final epilogue: + bias + residual, bf16 store, and (optionally) the eps-1e-5 LayerNorm statistics of the output
//     rows for the next block's norm1.
// Warp 0 = TMA producer, warp 1 = tcgen05.mma issue (uniform loop, elect.sync), warps 2..9 = epilogue (two warps per
// TMEM lane quarter, each half of the columns).
#include <cstdio>

#include "kernels.h"
#include "tc_common.cuh"

namespace sj {
namespace {

using namespace tc;

constexpr int C = 96, HID = 384, HQ = HID / 4, NTHREADS = 320;
constexpr int X_CHUNK = 128 * 64;           // 8 KB: 128 tokens x 32 channels (SWIZZLE_64B)
constexpr int W1_CHUNK = HID * 64;          // 24 KB: 384 rows x 32 channels
constexpr int W2_CHUNK = C * 64;            // 6 KB: 96 rows x 32 hidden units
constexpr int H_CHUNK = 128 * 64;           // 8 KB: 128 tokens x 32 hidden units
constexpr int H_BUF = 3 * H_CHUNK;          // one quarter of the hidden tile
constexpr int OFF_X = 0;
constexpr int OFF_W1 = OFF_X + 3 * X_CHUNK;
constexpr int OFF_W2 = OFF_W1 + 3 * W1_CHUNK;
constexpr int OFF_H = OFF_W2 + 12 * W2_CHUNK;
constexpr int OFF_VEC = OFF_H + 2 * H_BUF;              // float colsum[384], bias1[384], bias2[96]
constexpr int OFF_STAT = OFF_VEC + (2 * HID + C) * 4;   // float2 [128]
constexpr int OFF_BAR = OFF_STAT + 128 * 8;
constexpr int SMEM_BYTES = OFF_BAR + 256;
static_assert(OFF_W2 % 1024 == 0 && OFF_H % 1024 == 0, "tile alignment");
static_assert(SMEM_BYTES + 1024 <= 227 * 1024, "shared memory budget");
constexpr uint32_t TM_Q = 0, TM_O = 384;    // TMEM columns: four fc1 quarters (96 each), fc2 output

enum { B_XFULL = 0, B_XEMPTY, B_WFULL, B_Q0, B_Q1, B_Q2, B_Q3, B_HFULL0, B_HFULL1, B_HEMPTY0, B_HEMPTY1, B_OFULL, B_OEMPTY, B_COUNT };

struct MlpP {
  int M, num_tiles;
  int ctas;            // CTAs serving this group (a launch may carry two independent groups: different x / weights)
  const bf16* x;       // [M, 96] input of the MLP half (= residual)
  bf16* out;           // [M, 96]
  const float* mean;   // norm2 statistics of x
  const float* rstd;
  const float* colsum; // [384] column sums of the folded fc1 weights
  const float* bias1;  // [384] beta . W1 + b1
  const float* bias2;  // [96]
  float* st_mean;      // optional: LayerNorm statistics (eps 1e-5) of the output rows
  float* st_rstd;
};

struct MlpPPair { MlpP g[2]; };
struct PairMaps { CUtensorMap m[2][3]; };

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(NTHREADS, 1)
tc_mlp96_kernel(const __grid_constant__ PairMaps maps, const MlpPPair pp) {
  // group of this CTA: CTAs [0, pp.g[0].ctas) run group 0, the rest group 1 (same shapes, other tensors: the flow / raster
  // branches of the encoder in lock step).  Both parameter sets sit in one array so that the choice is a constant-bank
  // offset, not a select per field.
  const int grp = blockIdx.x >= (unsigned)pp.g[0].ctas ? 1 : 0;
  const MlpP& p = pp.g[grp];
  const CUtensorMap* mapX = &maps.m[grp][0];
  const CUtensorMap* mapW1 = &maps.m[grp][1];
  const CUtensorMap* mapW2 = &maps.m[grp][2];
  const int cta = (int)blockIdx.x - (grp ? pp.g[0].ctas : 0), nctas = p.ctas;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + B_COUNT);
  float* vec = reinterpret_cast<float*>(smem + OFF_VEC);
  float2* stat_s = reinterpret_cast<float2*>(smem + OFF_STAT);

  const int warp = uniform_warp_idx(), lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(mapX);
    prefetch_tmap(mapW1);
    prefetch_tmap(mapW2);
    const int counts[B_COUNT] = {1, 1, 1, 1, 1, 1, 1, 8, 8, 1, 1, 1, 8};
    for (int i = 0; i < B_COUNT; ++i) mbar_init(&bar[i], counts[i]);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < HID; i += NTHREADS) {
    vec[i] = p.colsum[i];
    vec[HID + i] = p.bias1[i];
  }
  for (int i = threadIdx.x; i < C; i += NTHREADS) vec[2 * HID + i] = p.bias2[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp != 0) {  // the producer lane waits after it has issued the (constant) resident weights
    pdl_wait();
    pdl_trigger();
  }

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&bar[B_WFULL], 3 * W1_CHUNK + 12 * W2_CHUNK);
      for (int c = 0; c < 3; ++c)
        for (int hf = 0; hf < 2; ++hf)
          tma_load_2d(smem + OFF_W1 + c * W1_CHUNK + hf * 192 * 64, mapW1, &bar[B_WFULL], c * 32, hf * 192);
      for (int c = 0; c < 12; ++c) tma_load_2d(smem + OFF_W2 + c * W2_CHUNK, mapW2, &bar[B_WFULL], c * 32, 0);
      pdl_wait();
      uint32_t it = 0;
      for (int tile = cta; tile < p.num_tiles; tile += nctas, ++it) {
        mbar_wait(&bar[B_XEMPTY], (it & 1) ^ 1);
        mbar_expect_tx(&bar[B_XFULL], 3 * X_CHUNK);
        for (int c = 0; c < 3; ++c) tma_load_2d(smem + OFF_X + c * X_CHUNK, mapX, &bar[B_XFULL], c * 32, tile * 128);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t HI64 = desc_hi(64, 512);
    const uint32_t idq = make_idesc_bf16(128, HQ);  // N = 96 for fc1 quarters and for fc2
    const uint32_t x_lo = desc_lo(smem_u32(smem + OFF_X)), w1_lo = desc_lo(smem_u32(smem + OFF_W1));
    const uint32_t w2_lo = desc_lo(smem_u32(smem + OFF_W2)), h_lo = desc_lo(smem_u32(smem + OFF_H));
    mbar_wait(&bar[B_WFULL], 0);
    uint32_t it = 0;
    for (int tile = cta; tile < p.num_tiles; tile += nctas, ++it) {
      const uint32_t ph = it & 1;
      mbar_wait(&bar[B_XFULL], ph);
      tc_fence_after();
      if (elect_one()) {
        // fc1, four quarters of 96 hidden units; K = 96 = three 32-channel chunks x two K steps
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
          for (int s = 0; s < 6; ++s) {
            const uint32_t c = s >> 1, k = s & 1;
            umma_bf16_w(tmem + TM_Q + q * HQ, x_lo + (c * X_CHUNK >> 4) + 2 * k, HI64,
                        w1_lo + ((c * W1_CHUNK + q * HQ * 64) >> 4) + 2 * k, HI64, idq, s != 0);
          }
          umma_commit(&bar[B_Q0 + q]);
        }
        umma_commit(&bar[B_XEMPTY]);
      }
      __syncwarp();
      // fc2: quarter q comes through hidden buffer q & 1 (each buffer is filled twice per tile)
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        const uint32_t use = 2 * it + (q >> 1);  // how often buffer q & 1 has been used before
        mbar_wait(&bar[B_HFULL0 + (q & 1)], use & 1);
        if (q == 0) mbar_wait(&bar[B_OEMPTY], ph ^ 1);  // output accumulator drained by the previous tile
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int s = 0; s < 6; ++s) {
            const uint32_t c = s >> 1, k = s & 1;
            umma_bf16_w(tmem + TM_O, h_lo + (((q & 1) * H_BUF + c * H_CHUNK) >> 4) + 2 * k, HI64,
                        w2_lo + ((3 * q + c) * W2_CHUNK >> 4) + 2 * k, HI64, idq, (q | s) != 0);
          }
          umma_commit(&bar[B_HEMPTY0 + (q & 1)]);
          if (q == 3) umma_commit(&bar[B_OFULL]);
        }
        __syncwarp();
      }
    }
  } else {
    const int quarter = warp % 4, half = (warp - 2) / 4;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float* cs = vec;
    const float* b1 = vec + HID;
    const float* b2 = vec + 2 * HID;
    uint32_t it = 0;
    for (int tile = cta; tile < p.num_tiles; tile += nctas, ++it) {
      const uint32_t ph = it & 1;
      const long long row = (long long)tile * 128 + r;
      const bool valid = row < p.M;
      const float mean = valid ? p.mean[row] : 0.f, rstd = valid ? p.rstd[row] : 1.f;
      // residual: this thread's 48 output columns, fetched long before they are needed
      uint4 res[6];
      if (valid) {
#pragma unroll
        for (int i = 0; i < 6; ++i) res[i] = *reinterpret_cast<const uint4*>(p.x + row * C + half * 48 + 8 * i);
      }
      // ---- hidden quarters: LN fold + GELU -> bf16 A operand of fc2 (this thread: 48 of the 96 units) ----
      // Software pipeline over 8 parts (quarter q = part / 2; part 0: units +0..31, part 1: units +32..47): the TMEM load
      // of part i+1 is in flight while part i goes through the fold / GELU / shared-memory stores.
      uint32_t ta[32], tb[32];
      auto issue = [&](int part, uint32_t (&t)[32]) {
        const int q = part >> 1;
        const uint32_t tq = tmem + lane_addr + TM_Q + q * HQ + half * 48;
        if (!(part & 1)) {
          mbar_wait(&bar[B_Q0 + q], ph);
          tc_fence_after();
          tmem_ld32_issue(tq, t);
        } else {
          tmem_ld16_issue(tq + 32, t);
        }
      };
      auto process = [&](int part, uint32_t (&t)[32]) {
        const int q = part >> 1, sub = part & 1, cnt = sub ? 16 : 32;
        const uint32_t use = 2 * it + (q >> 1);
        if (!sub) mbar_wait(&bar[B_HEMPTY0 + (q & 1)], (use & 1) ^ 1);  // fc2 has finished reading the previous tenant
        uint8_t* hbuf = smem + OFF_H + (q & 1) * H_BUF;
        const int nl = half * 48 + sub * 32;  // quarter-local first hidden unit of this part
        const int ng = q * HQ + nl;           // global hidden unit
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          if (i < cnt) {
            const float4 c4 = *reinterpret_cast<const float4*>(cs + ng + i);
            const float4 d4 = *reinterpret_cast<const float4*>(b1 + ng + i);
            v[i] = act_fast(rstd * (__uint_as_float(t[i]) - mean * c4.x) + d4.x, ACT_GELU);
            v[i + 1] = act_fast(rstd * (__uint_as_float(t[i + 1]) - mean * c4.y) + d4.y, ACT_GELU);
            v[i + 2] = act_fast(rstd * (__uint_as_float(t[i + 2]) - mean * c4.z) + d4.z, ACT_GELU);
            v[i + 3] = act_fast(rstd * (__uint_as_float(t[i + 3]) - mean * c4.w) + d4.w, ACT_GELU);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (8 * j < cnt) {
            const int n = nl + 8 * j;  // multiple of 8: one 16-byte unit of a 64-byte swizzled row
            uint8_t* dst = hbuf + (n >> 5) * H_CHUNK + r * 64 + ((((n & 31) >> 3) ^ ((r >> 1) & 3)) << 4);
            st8_bf16(reinterpret_cast<bf16*>(dst), v + 8 * j);
          }
        }
        if (sub) {
          fence_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar[B_HFULL0 + (q & 1)]);
        }
      };
      issue(0, ta);
#pragma unroll
      for (int part = 0; part < 8; part += 2) {
        tmem_ld_wait(ta);
        issue(part + 1, tb);
        process(part, ta);
        tmem_ld_wait(tb);
        if (part + 2 < 8) issue(part + 2, ta);
        process(part + 1, tb);
      }
      // ---- output: + bias + residual, statistics, store ----
      mbar_wait(&bar[B_OFULL], ph);
      tc_fence_after();
      float st_s = 0.f, st_q = 0.f;
      {
        float v[32], u[16];
        tmem_ld32(tmem + lane_addr + TM_O + half * 48, v);
        tmem_ld16(tmem + lane_addr + TM_O + half * 48 + 32, u);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar[B_OEMPTY]);  // accumulator drained: the next tile's fc2 may start
        if (valid) {
          bf16* dst = p.out + row * C + half * 48;
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            float* s = i < 4 ? v + 8 * i : u + 8 * (i - 4);
            const uint32_t w[4] = {res[i].x, res[i].y, res[i].z, res[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              s[2 * j] += b2[half * 48 + 8 * i + 2 * j] + __uint_as_float(w[j] << 16);
              s[2 * j + 1] += b2[half * 48 + 8 * i + 2 * j + 1] + __uint_as_float(w[j] & 0xffff0000u);
            }
            if (p.st_mean) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float q = __bfloat162float(__float2bfloat16_rn(s[j]));
                st_s += q;
                st_q = fmaf(q, q, st_q);
              }
            }
            st8_bf16(dst + 8 * i, s);
          }
        }
      }
      if (p.st_mean) {  // the two warps of a lane quarter each hold half of the row
        if (half) stat_s[r] = make_float2(st_s, st_q);
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
        if (!half && valid) {
          const float2 o2 = stat_s[r];
          const float mu = (st_s + o2.x) * (1.0f / C);
          p.st_mean[row] = mu;
          p.st_rstd[row] = rsqrtf(fmaxf((st_q + o2.y) * (1.0f / C) - mu * mu, 0.f) + 1e-5f);
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");  // stat_s is reused by the next tile
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace

bool tc_mlp96_supported(int Cc, int hidden, const SjSwinBlockW& w) {
  return Cc == C && hidden == HID && w.fc1.w_tc && w.fc1.tc_colsum && w.fc1.tc_bias && w.fc2.w_tc && w.fc2.b;
}

// out = x + fc2(GELU(fc1(LN(x)))) for x, out bf16 [M, 96]; mean/rstd = norm2 statistics of x.  groups = 2: two independent
// problems of the same size in ONE launch (half of the CTAs each)
static void tc_mlp96_launch(Ctx& c, int groups, const void* const* x, void* const* out, const float* const* mean,
                            const float* const* rstd, const SjSwinBlockW* const* w, int M, float* const* st_mean,
                            float* const* st_rstd) {
  if (!c.ok() || c.dry) return;
  PairMaps pm;
  CUtensorMap (&maps)[2][3] = pm.m;
  MlpPPair ppair = {};
  MlpP (&pp)[2] = ppair.g;
  const int tiles = cdiv(M, 128);
  int per = tiles < num_sms() / groups ? tiles : num_sms() / groups;
  for (int g = 0; g < groups; ++g) {
    if (!tc_mlp96_supported(C, HID, *w[g]) || !mean[g] || !rstd[g]) { c.fail(SJ_EINVAL); return; }
    uint64_t dx[2] = {(uint64_t)C, (uint64_t)M};
    uint64_t sx[1] = {(uint64_t)C * 2};
    uint32_t bx[2] = {32, 128};
    uint64_t d1[2] = {(uint64_t)C, (uint64_t)HID};
    uint32_t b1[2] = {32, 192};
    uint64_t d2[2] = {(uint64_t)HID, (uint64_t)C};
    uint64_t s2[1] = {(uint64_t)HID * 2};
    uint32_t b2[2] = {32, (uint32_t)C};
    if (!encode_tmap(&maps[g][0], x[g], 2, dx, sx, bx, 64) || !encode_tmap(&maps[g][1], w[g]->fc1.w_tc, 2, d1, sx, b1, 64) ||
        !encode_tmap(&maps[g][2], w[g]->fc2.w_tc, 2, d2, s2, b2, 64)) {
      snprintf(tls().cuda_err, sizeof(tls().cuda_err), "cuTensorMapEncodeTiled failed (tc_mlp96)");
      c.fail(SJ_ECUDA);
      return;
    }
    MlpP& p = pp[g];
    p.M = M; p.num_tiles = tiles; p.ctas = per;
    p.x = (const bf16*)x[g]; p.out = (bf16*)out[g]; p.mean = mean[g]; p.rstd = rstd[g];
    p.colsum = w[g]->fc1.tc_colsum; p.bias1 = w[g]->fc1.tc_bias; p.bias2 = w[g]->fc2.b;
    p.st_mean = st_mean ? st_mean[g] : nullptr; p.st_rstd = st_rstd ? st_rstd[g] : nullptr;
  }
  if (groups == 1) {
    pp[1] = pp[0];
    for (int i = 0; i < 3; ++i) maps[1][i] = maps[0][i];
  }
  if (!SJ_SMEM_LIMIT_OK((tc_mlp96_kernel), 227 * 1024)) {
    c.fail(SJ_ECUDA);
    return;
  }
  SJ_LAUNCH(c, "tc_mlp96", tc_mlp96_kernel, groups * per, NTHREADS, 1024 + SMEM_BYTES, pm, ppair);
}

void tc_mlp96(Ctx& c, const void* x, void* out, const float* mean, const float* rstd, const SjSwinBlockW& w, int M,
              float* st_mean, float* st_rstd) {
  const SjSwinBlockW* wp = &w;
  tc_mlp96_launch(c, 1, &x, &out, &mean, &rstd, &wp, M, &st_mean, &st_rstd);
}

void tc_mlp96_pair(Ctx& c, const void* const x[2], void* const out[2], const float* const mean[2], const float* const rstd[2],
                   const SjSwinBlockW* const w[2], int M, float* const st_mean[2], float* const st_rstd[2]) {
  tc_mlp96_launch(c, 2, x, out, mean, rstd, w, M, st_mean, st_rstd);
}

}  // namespace sj
