// Dense bf16 GEMM on the 5th-gen tensor cores: C = [R +] act(LNfold(A . W^T) + bias).
//
//   * persistent: grid = #SMs, each CTA strides over (group, m-tile, n-tile) work items;
//   * warp-specialised: warp 0 = TMA producer, warp 1 = TMEM allocator + tcgen05.mma issuer,
//     warps 2.. = epilogue (TMEM lane quarter = warp % 4; EW = 8 or 16 epilogue warps, i.e. 2 or 4 warps per quarter,
//     each taking a share of the columns: the epilogue is a chain of TMEM / global-memory round trips and more
//     resident warps hide it);
//   * operands staged by TMA (cp.async.bulk.tensor, SWIZZLE_128B, 64-element K blocks) through a
//     4-stage mbarrier ring; accumulators live in TMEM (fp32, 128 lanes x BN columns), double
//     buffered so the epilogue of tile i overlaps the main loop of tile i+1;
//   * epilogue: tcgen05.ld -> LayerNorm fold / bias / GELU|ELU / residual -> bf16 -> global.
//
// LayerNorm fold: LN(x).W + b == rstd*(x.(g*W) - mean*colsum(g*W)) + (beta.W + b), so the raw
// activations go through TMA untouched and the per-row affine is applied in the epilogue.
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "kernels.h"
#include "tc_common.cuh"

namespace sj {
namespace {

using namespace tc;

constexpr int BM = 128, BK = 64, STAGES = 4;
constexpr int nthreads(int ew) { return 64 + 32 * ew; }  // TMA warp, MMA warp, EW epilogue warps
constexpr int A_STAGE_BYTES_MAX = 256 * BK * 2;

struct KParams {
  int M, N, K, BN, groups, m_tiles, n_tiles, k_blocks;
  int a_mode, a_inner;  // a_mode 1: rows = [outer][G][inner]
  int mWo, mHoTile;     // a_mode 2 (PatchMerging): merged width, merged rows per m-tile
  int mC;
  const float* bias;
  int bias_gstride;
  int act;
  const bf16* R;
  int ldr;
  bf16* C;
  int ldc;
  RowMap cm;
  const float* ln_mean;
  const float* ln_rstd;
  const float* ln_s;
  int ln_gstride;
  int vec_smem;    // bias / column-sum vectors of all groups staged in shared memory (floats per vector, 0 = read from global)
  float* st_mean;  // optional: LayerNorm statistics of the OUTPUT rows (needs BN == N), written at the C row index
  float* st_rstd;
  float st_eps;
  int coal;       // output rows of a tile are consecutive in C: store through the per-warp transpose (see the epilogue)
  int stg_off;    // byte offset of the staging buffers from the barrier block
  int a_rows;     // rows per A stage (128; 256 for the shifted-view probe)
  int dbg_shift;  // probe: the MMA reads A rows [shift, shift+128) of the stage
  int dbg_bo;     // probe: set the descriptor base_offset field from the start address
};

// ACT / LN are compile-time so the fully unrolled epilogue stays small enough for the instruction caches
// RPF (residual prefetch): the residual rows of tile i+1 are fetched into registers while tile i is in its epilogue
// (each 32-column chunk is re-loaded for the next tile right after it has been consumed), so the epilogue never sits
// on a global-memory round trip.  Needs 32-column chunks only and at most 96 columns per thread.
template <int ACT, bool LN, bool RPF = false, int EW = 8>
__global__ void __launch_bounds__(nthreads(EW), 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const KParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_stage_bytes = p.BN * BK * 2;
  const int A_STAGE_BYTES = p.a_rows * BK * 2;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + STAGES * b_stage_bytes);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  constexpr int NTHREADS = nthreads(EW), NSUB = EW / 4;  // NSUB warps share each TMEM lane quarter
  float2* stat_s = reinterpret_cast<float2*>(bars + 32);  // [2 accumulator stages][4 column shares][128 rows] partial (sum, sum of squares)
  float* vec_s = reinterpret_cast<float*>(stat_s + 2 * 4 * BM);  // [bias: vec_smem floats][column sums: vec_smem floats]

  const int warp = uniform_warp_idx(), lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], EW);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above (barriers, TMEM, tensor-map prefetch) overlaps the previous kernel's tail
  pdl_trigger();

  const int tiles_per_group = p.m_tiles * p.n_tiles;
  const int num_tiles = tiles_per_group * p.groups;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int g = tile / tiles_per_group, rem = tile % tiles_per_group;
        const int mt = rem / p.n_tiles, nt = rem % p.n_tiles;
        const int m0 = mt * BM, n0 = nt * p.BN;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], A_STAGE_BYTES + b_stage_bytes);
          void* da = smem_a + stage * A_STAGE_BYTES;
          int k0 = kb * BK;
          if (p.a_mode == 0) {
            tma_load_2d(da, &mapA, &full_bar[stage], k0, m0);
          } else if (p.a_mode == 1) {
            tma_load_4d(da, &mapA, &full_bar[stage], k0, m0 % p.a_inner, g, m0 / p.a_inner);
          } else {
            // PatchMerging gather: k = q*C + c, q -> (dy, dx) = (q & 1, q >> 1); view {C, 2(dx), W/2, 2(dy), B*H/2}
            // C need not be a multiple of 64: each quadrant is covered by ceil(C/64) blocks; the channels a box
            // reads past C are zero-filled by TMA, which cancels the weight columns of the next quadrant.
            const int cpb = (p.mC + BK - 1) / BK;
            const int q = kb / cpb, c0 = (kb % cpb) * BK;
            const int r0 = (m0 / p.mWo);  // first merged row (b*Ho + i) of this tile
            tma_load_5d(da, &mapA, &full_bar[stage], c0, q >> 1, 0, q & 1, r0);
            k0 = q * p.mC + c0;
          }
          tma_load_3d(smem_b + stage * b_stage_bytes, &mapB, &full_bar[stage], k0, n0, g);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // the whole warp runs the loop (uniform control flow keeps the descriptors in uniform registers); one elected lane
    // issues the MMAs and commits
    const uint32_t idesc = make_idesc_bf16(BM, p.BN);
    int stage = 0, as = 0;
    uint32_t phase = 0, aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * p.BN;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem_a + stage * A_STAGE_BYTES) + p.dbg_shift * 128;
        uint64_t da = make_smem_desc(a_addr, 128);
        if (p.dbg_bo) da |= (uint64_t)((a_addr >> 7) & 7) << 49;  // matrix base offset, bits [49,52)
        const uint64_t db = make_smem_desc(smem_u32(smem_b + stage * b_stage_bytes), 128);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)  // +32 bytes per K=16 step inside the 128B swizzle row
            umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (kb == p.k_blocks - 1) umma_commit(&tfull_bar[as]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else {
    // NSUB warps per TMEM lane quarter; the BN/16 column units are dealt out as evenly as possible, low shares first
    const int quarter = warp % 4, sub = (warp - 2) / 4;
    const int n16 = p.BN / 16, ubase = n16 / NSUB, urem = n16 % NSUB;
    const int c_begin = 16 * (sub * ubase + min(sub, urem)), c_end = c_begin + 16 * (ubase + (sub < urem ? 1 : 0));
    if (p.vec_smem) {
      // per-column epilogue vectors: staged once, then read as shared-memory broadcasts (global loads of them miss L1
      // behind the streaming residual / output traffic and show up as long-scoreboard stalls)
      const int et = threadIdx.x - 64;
      for (int i = et; i < p.vec_smem; i += NTHREADS - 64) {
        vec_s[i] = p.bias ? p.bias[i] : 0.f;
        if (LN) vec_s[p.vec_smem + i] = p.ln_s[i];
      }
      asm volatile("bar.sync 5, %0;" ::"n"(EW * 32) : "memory");
    }
    // Coalesced stores: thread r owns output row r, whose neighbour rows are ldc * 2 bytes away, so a direct 16-byte store
    // per thread costs 32 L1 wavefronts per instruction (measured: the LSU wavefront pipe was the busiest unit of the
    // store-heavy GEMMs).  When the rows of a tile are consecutive in C, each 32-column chunk goes through a 2 KB per-warp
    // transpose instead and four consecutive lanes write one row's 64 bytes (8 wavefronts per instruction).
    uint8_t* stg = reinterpret_cast<uint8_t*>(bars) + p.stg_off + (warp - 2) * 2048;
    int as = 0;
    uint32_t aphase = 0;
    // output row of this thread in a tile: validity, C/R row index (after the row maps), group, first column
    auto coords = [&](int tile, int& g, int& n0, int& m, bool& valid, long long& crow) {
      g = tile / tiles_per_group;
      const int rem = tile % tiles_per_group;
      const int mt = rem / p.n_tiles, nt = rem % p.n_tiles;
      m = mt * BM + quarter * 32 + lane;
      n0 = nt * p.BN;
      valid = m < p.M;
      crow = 0;
      if (valid) {
        crow = m;
        if (p.cm.inner > 0) crow = (long long)(m / p.cm.inner) * p.cm.outer + (m % p.cm.inner);
        crow += (long long)g * p.cm.gstride;
        if (p.cm.map) crow = (crow / p.cm.map_len) * p.cm.map_len + p.cm.map[crow % p.cm.map_len];
      }
    };
    constexpr int RCH = 3;  // RPF: at most three 32-column chunks per thread
    uint4 rn[RPF ? 4 * RCH : 1];
    auto prefetch = [&](int ci, long long crow_, int n0_) {
      if constexpr (RPF) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          rn[4 * ci + i] = *reinterpret_cast<const uint4*>(p.R + crow_ * p.ldr + n0_ + c_begin + 32 * ci + 8 * i);
      }
    };
    if constexpr (RPF) {
      if ((int)blockIdx.x < num_tiles) {
        int g, n0, m;
        bool valid;
        long long crow;
        coords(blockIdx.x, g, n0, m, valid, crow);
#pragma unroll
        for (int ci = 0; ci < RCH; ++ci)
          if (valid && c_begin + 32 * ci < c_end) prefetch(ci, crow, n0);
      }
    }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int g, n0, m;
      bool valid;
      long long crow;
      coords(tile, g, n0, m, valid, crow);
      bool nvalid = false;
      long long ncrow = 0;
      int nn0 = 0;
      if constexpr (RPF) {
        if (tile + (int)gridDim.x < num_tiles) {
          int ng, nm;
          coords(tile + gridDim.x, ng, nn0, nm, nvalid, ncrow);
        }
      }
      float mean = 0.f, rstd = 1.f;
      if (valid && LN) {
        // LN statistics are indexed like the A rows
        long long arow = m;
        if (p.a_mode == 1) arow = ((long long)(m / p.a_inner) * p.groups + g) * p.a_inner + (m % p.a_inner);
        mean = p.ln_mean[arow];
        rstd = p.ln_rstd[arow];
      }
      const float* bias = p.bias ? (p.vec_smem ? vec_s : p.bias) + (long long)g * p.bias_gstride : nullptr;
      const float* lns = LN ? (p.vec_smem ? vec_s + p.vec_smem : p.ln_s) + (long long)g * p.ln_gstride : nullptr;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * p.BN;
      float st_s = 0.f, st_q = 0.f;  // statistics of this thread's share of the (bf16-rounded) output row
      // residual rows are fetched BEFORE the TMEM load of the same columns so the global-memory latency overlaps
      // the tcgen05.ld round trip; everything is fully unrolled so v[] / rr[] stay in registers
      auto chunk = [&](auto cnt_tag, int c, auto ci_tag) {
        constexpr int CNT = decltype(cnt_tag)::value;
        constexpr int CI = decltype(ci_tag)::value;  // RPF: which chunk of the prefetched residual registers
        const int n = n0 + c;
        float v[CNT];
        uint4 rr[CNT / 8];
        if constexpr (!RPF) {
          if (valid && p.R) {
#pragma unroll
            for (int i = 0; i < CNT / 8; ++i) rr[i] = *reinterpret_cast<const uint4*>(p.R + crow * p.ldr + n + 8 * i);
          }
        }
        if constexpr (CNT == 32) tmem_ld32(t_addr + c, v);
        else tmem_ld16(t_addr + c, v);
        if constexpr (RPF) {
#pragma unroll
          for (int i = 0; i < CNT / 8; ++i) rr[i] = rn[4 * CI + i];
          if (nvalid) prefetch(CI, ncrow, nn0);  // this chunk's registers are free again: fetch the next tile's
        }
        if (!valid && !p.coal) return;
#pragma unroll
        for (int i = 0; i < CNT; i += 4) {  // 16-byte loads of the per-column vectors (L1-resident)
          float4 l4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = l4;
          if (LN) l4 = *reinterpret_cast<const float4*>(lns + n + i);
          if (bias) b4 = *reinterpret_cast<const float4*>(bias + n + i);
          if (LN) {
            v[i] = rstd * (v[i] - mean * l4.x); v[i + 1] = rstd * (v[i + 1] - mean * l4.y);
            v[i + 2] = rstd * (v[i + 2] - mean * l4.z); v[i + 3] = rstd * (v[i + 3] - mean * l4.w);
          }
          v[i] = act_fast(v[i] + b4.x, ACT); v[i + 1] = act_fast(v[i + 1] + b4.y, ACT);
          v[i + 2] = act_fast(v[i + 2] + b4.z, ACT); v[i + 3] = act_fast(v[i + 3] + b4.w, ACT);
        }
#pragma unroll
        for (int i = 0; i < CNT; i += 8) {
          if (p.R && valid) {
            const uint32_t w[4] = {rr[i / 8].x, rr[i / 8].y, rr[i / 8].z, rr[i / 8].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              v[i + 2 * j] += __uint_as_float(w[j] << 16);
              v[i + 2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
            }
          }
          if (p.st_mean) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float r = __bfloat162float(__float2bfloat16_rn(v[i + j]));
              st_s += r;
              st_q = fmaf(r, r, st_q);
            }
          }
          if (p.coal) st8_bf16(reinterpret_cast<bf16*>(stg + lane * 64 + (((i / 8) ^ ((lane >> 1) & 3)) << 4)), v + i);
          else st8_bf16(p.C + crow * p.ldc + n + i, v + i);
        }
        if (p.coal) {
          constexpr int PER = CNT / 8;  // 16-byte pieces per row in this chunk
          __syncwarp();
          const long long crow0 = __shfl_sync(0xffffffffu, crow, 0);  // lane 0 of a warp with any valid row is valid
          const int rows_ok = min(32, p.M - (m - lane));               // valid rows of this warp
#pragma unroll
          for (int i = 0; i < PER; ++i) {
            const int gidx = 32 * i + lane, row = gidx / PER, pc = gidx % PER;
            if (row < rows_ok) {
              const uint4 o = *reinterpret_cast<const uint4*>(stg + row * 64 + ((pc ^ ((row >> 1) & 3)) << 4));
              *reinterpret_cast<uint4*>(p.C + (crow0 + row) * p.ldc + n + pc * 8) = o;
            }
          }
          __syncwarp();
        }
      };
      if constexpr (RPF) {
        if (c_begin < c_end) chunk(std::integral_constant<int, 32>{}, c_begin, std::integral_constant<int, 0>{});
        if (c_begin + 32 < c_end) chunk(std::integral_constant<int, 32>{}, c_begin + 32, std::integral_constant<int, 1>{});
        if (c_begin + 64 < c_end) chunk(std::integral_constant<int, 32>{}, c_begin + 64, std::integral_constant<int, 2>{});
      } else {
        int c = c_begin;
        for (; c + 32 <= c_end; c += 32) chunk(std::integral_constant<int, 32>{}, c, std::integral_constant<int, 0>{});
        if (c < c_end) chunk(std::integral_constant<int, 16>{}, c, std::integral_constant<int, 0>{});
      }
      if (p.st_mean) {
        // the warps of a TMEM lane quarter each hold a share of the row: combine through shared memory (fixed order)
        const int rloc = quarter * 32 + lane;
        if (sub) stat_s[(as * 4 + sub) * BM + rloc] = make_float2(st_s, st_q);
        asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "n"(NSUB * 32) : "memory");
        if (!sub && valid) {
          float2 o2 = stat_s[(as * 4 + 1) * BM + rloc];
#pragma unroll
          for (int q = 2; q < NSUB; ++q) {
            const float2 t = stat_s[(as * 4 + q) * BM + rloc];
            o2.x += t.x;
            o2.y += t.y;
          }
          const float mu = (st_s + o2.x) / (float)p.N;
          const float var = fmaxf((st_q + o2.y) / (float)p.N - mu * mu, 0.f);
          p.st_mean[crow] = mu;
          p.st_rstd[crow] = rsqrtf(var + p.st_eps);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int pick_bn(int N) {
  for (int bn = 256; bn >= 16; bn -= 16)
    if (N % bn == 0) return bn;
  return 0;
}

}  // namespace

// ---- tensor-map encoder (driver entry point fetched at run time; no link-time libcuda dependency) -----
bool encode_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box, int swizzle_bytes, int elem_bytes) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || !ptr)
      return false;
    fn = reinterpret_cast<EncodeFn>(ptr);
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                  reinterpret_cast<const cuuint64_t*>(dims), reinterpret_cast<const cuuint64_t*>(strides_bytes),
                  reinterpret_cast<const cuuint32_t*>(box), estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

static int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

bool tc_gemm_stats_ok(int N) { return N % 16 == 0 && pick_bn(N) == N; }

bool tc_gemm_supported(const TcGemmP& a) {
  if (!a.Bw || a.K < 64 || a.K % 32 || a.N % 16 || pick_bn(a.N) == 0) return false;
  if (a.a_mode == 0 && (a.lda % 8)) return false;
  if (a.a_mode == 1 && (a.a_inner % BM)) return false;
  if (a.a_mode == 2) {
    int Wo = a.mW / 2;
    if (a.mC % 32 || Wo <= 0 || BM % Wo || (a.mH / 2) % (BM / Wo) || a.K != 4 * a.mC) return false;
  }
  if (a.ldc % 8 || (a.R && a.ldr % 8)) return false;
  return true;
}

void tc_gemm(Ctx& c, const TcGemmP& a) {
  if (!c.ok() || c.dry) return;
  if (!tc_gemm_supported(a)) { c.fail(SJ_EUNSUPPORTED); return; }
  KParams p{};
  p.M = a.M; p.N = a.N; p.K = a.K; p.groups = a.groups < 1 ? 1 : a.groups;
  p.BN = pick_bn(a.N);
  p.m_tiles = cdiv(a.M, BM);
  if (!a.st_mean) {
    // n-tile width: minimise  waves x (c0 + BN)  with waves = ceil(tiles / #SMs) and c0 = 160 columns' worth of per-tile
    // fixed cost (A tile load, pipeline fill, epilogue drain).  Fitted to tools/gemm_bn_sweep.py on B200 (batch-16 shapes,
    // M = 4096: N = 384 -> 96 (one wave of 128 tiles; the old "largest BN with >= #SMs tiles" rule took 64 = two waves, 6.9
    // -> 5.7 us, 14.2 -> 9.8 us at K = 1536), N = 1152 -> 128 (9.8 -> 8.4 us), N = 1536 -> 192 (14.4 -> 11.8 us); M = 16384
    // shapes keep their width).  Ties go to the wider tile.
    long long best = -1;
    for (int bn = 256; bn >= 32; bn -= 16) {
      if (a.N % bn) continue;
      const long long tiles = (long long)p.m_tiles * (a.N / bn) * p.groups;
      const long long cost = ((tiles + num_sms() - 1) / num_sms()) * (160 + bn);
      if (best < 0 || cost < best) { best = cost; p.BN = bn; }
    }
    // tuning knob (tools/gemm_bn_sweep.py): force the n-tile width; read per call so that one process can sweep it
    if (const char* e = getenv("SJ_TCG_BN")) {
      const int bn = atoi(e);
      if (bn >= 16 && bn <= 256 && bn % 16 == 0 && a.N % bn == 0) p.BN = bn;
    }
  }
  p.n_tiles = a.N / p.BN;
  p.k_blocks = a.a_mode == 2 ? 4 * cdiv(a.mC, BK) : cdiv(a.K, BK);
  p.a_mode = a.a_mode; p.a_inner = a.a_inner;
  p.bias = a.bias; p.bias_gstride = a.bias_gstride; p.act = a.act;
  p.R = (const bf16*)a.R; p.ldr = a.ldr; p.C = (bf16*)a.C; p.ldc = a.ldc; p.cm = a.cm;
  p.ln_mean = a.ln_mean; p.ln_rstd = a.ln_rstd; p.ln_s = a.ln_s; p.ln_gstride = a.ln_gstride;
  if (a.st_mean) {
    if (p.n_tiles != 1 || !a.st_rstd) { c.fail(SJ_EUNSUPPORTED); return; }
    p.st_mean = a.st_mean; p.st_rstd = a.st_rstd; p.st_eps = a.st_eps;
  }
  p.a_rows = a.dbg_shift > 0 ? 256 : BM; p.dbg_shift = a.dbg_shift; p.dbg_bo = a.dbg_bo;

  CUtensorMap mapA, mapB;
  bool ok = true;
  if (a.a_mode == 0) {
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.M + (a.dbg_shift > 0 ? 256 : 0)};  // the probe reads past M
    uint64_t str[1] = {(uint64_t)a.lda * 2};
    uint32_t box[2] = {BK, (uint32_t)p.a_rows};
    ok = encode_tmap(&mapA, a.A, 2, dims, str, box, 128);
  } else if (a.a_mode == 1) {
    // rows ordered [outer][G][inner]; M counts rows per group = outer*inner
    uint64_t outer = (uint64_t)a.M / a.a_inner;
    uint64_t dims[4] = {(uint64_t)a.K, (uint64_t)a.a_inner, (uint64_t)p.groups, outer};
    uint64_t str[3] = {(uint64_t)a.lda * 2, (uint64_t)a.lda * 2 * a.a_inner, (uint64_t)a.lda * 2 * a.a_inner * p.groups};
    uint32_t box[4] = {BK, BM, 1, 1};
    ok = encode_tmap(&mapA, a.A, 4, dims, str, box, 128);
  } else {
    const int C = a.mC, H = a.mH, W = a.mW, Wo = W / 2;
    const uint64_t rows = (uint64_t)a.M / Wo;  // B * H/2
    p.mWo = Wo; p.mHoTile = BM / Wo; p.mC = C;
    uint64_t dims[5] = {(uint64_t)C, 2, (uint64_t)Wo, 2, rows};
    uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)C * 4, (uint64_t)W * C * 2, (uint64_t)W * C * 4};
    uint32_t box[5] = {BK, 1, (uint32_t)Wo, 1, (uint32_t)(BM / Wo)};
    ok = encode_tmap(&mapA, a.A, 5, dims, str, box, 128);
    (void)H;
  }
  {
    uint64_t dims[3] = {(uint64_t)a.K, (uint64_t)a.N, (uint64_t)p.groups};
    uint64_t str[2] = {(uint64_t)a.K * 2, (uint64_t)a.K * 2 * a.N};
    uint32_t box[3] = {BK, (uint32_t)p.BN, 1};
    ok = ok && encode_tmap(&mapB, a.Bw, 3, dims, str, box, 128);
  }
  if (!ok) {
    snprintf(tls().cuda_err, sizeof(tls().cuda_err), "cuTensorMapEncodeTiled failed (tc_gemm M=%d N=%d K=%d)", a.M, a.N, a.K);
    c.fail(SJ_ECUDA);
    return;
  }
  // Two experiment knobs, both measured neutral on B200 at batch 16 and therefore off by default:
  // SJ_TCG_EW=16: 16 epilogue warps (4 per TMEM lane quarter; needs >= 16 columns per warp);
  // SJ_TCG_RPF=1: the residual-prefetch epilogue (145+ registers: 8 warps only).
  static const bool ew16_on = getenv("SJ_TCG_EW") != nullptr && atoi(getenv("SJ_TCG_EW")) == 16;
  static const bool rpf_on = getenv("SJ_TCG_RPF") != nullptr;
  const bool ew16 = ew16_on && p.BN >= 64;
  const size_t ring = (size_t)STAGES * (p.a_rows * BK * 2 + p.BN * BK * 2);
  size_t tail = 256 + 2 * 4 * BM * sizeof(float2);  // barriers, output-statistics partials
  // rows of a tile consecutive in C <=> no scatter map and the [outer][G][inner] interleave (if any) is tile-aligned
  static const bool coal_off = getenv("SJ_TCG_NO_COALESCE") != nullptr;
  const size_t staging = (size_t)(ew16 ? 16 : 8) * 2048;
  p.coal = !coal_off && !a.cm.map && (a.cm.inner == 0 || a.cm.inner % BM == 0) && a.ldc % 8 == 0 &&
           1024 + ring + tail + staging <= 227 * 1024;
  {
    // stage the per-column vectors when they are dense ([groups][N]) and fit beside the operand ring
    const int vec = p.groups * a.N;
    const bool dense = (p.groups == 1 || ((!a.bias || a.bias_gstride == a.N) && (!a.ln_mean || a.ln_gstride == a.N)));
    const size_t need = (size_t)vec * 4 * (a.ln_mean ? 2 : 1);
    if ((a.bias || a.ln_mean) && dense && 1024 + ring + tail + need + (p.coal ? staging : 0) <= 225 * 1024 && need <= 24 * 1024) {
      p.vec_smem = vec;
      tail += need;
    }
  }
  tail = (tail + 15) & ~size_t(15);
  p.stg_off = (int)tail;
  const size_t smem = 1024 + ring + tail + (p.coal ? staging : 0);
  const int tiles = p.m_tiles * p.n_tiles * p.groups;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  // at least ~115 KB so that two CTAs (each allocating all 512 TMEM columns) can never share an SM
  const size_t smem_launch = smem < 120 * 1024 ? 120 * 1024 : smem;
  const bool ln = a.ln_mean != nullptr;
#define SJ_TCG2(ACT_, LN_, RPF_, EW_)                                                                                 \
  do {                                                                                                                \
    if (!SJ_SMEM_LIMIT_OK((tc_gemm_kernel<ACT_, LN_, RPF_, EW_>), 227 * 1024)) {                                                            \
      c.fail(SJ_ECUDA);                                                                                               \
      return;                                                                                                         \
    }                                                                                                                 \
    SJ_LAUNCH(c, "tc_gemm", (tc_gemm_kernel<ACT_, LN_, RPF_, EW_>), grid, nthreads(EW_), smem_launch, mapA, mapB, p); \
  } while (0)
#define SJ_TCG(ACT_, LN_, RPF_)                    \
  do {                                             \
    if (ew16 && !(RPF_)) SJ_TCG2(ACT_, LN_, false, 16); \
    else SJ_TCG2(ACT_, LN_, RPF_, 8);              \
  } while (0)
  // residual prefetch: both column halves of a tile must be whole 32-column chunks, at most three per thread
  const int csplit = ((p.BN / 16 + 1) / 2) * 16, w0 = csplit, w1 = p.BN - csplit;
  const bool rpf = rpf_on && a.R && !ln && w0 % 32 == 0 && w1 % 32 == 0 && w0 <= 96 && w1 <= 96 && w1 > 0;
  if (a.act == ACT_GELU) { if (ln) SJ_TCG(ACT_GELU, true, false); else SJ_TCG(ACT_GELU, false, false); }
  else if (a.act == ACT_ELU) {
    if (ln) SJ_TCG(ACT_ELU, true, false);
    else if (rpf) SJ_TCG(ACT_ELU, false, true);
    else SJ_TCG(ACT_ELU, false, false);
  } else {
    if (ln) SJ_TCG(ACT_NONE, true, false);
    else if (rpf) SJ_TCG(ACT_NONE, false, true);
    else SJ_TCG(ACT_NONE, false, false);
  }
#undef SJ_TCG2
#undef SJ_TCG
}

// ---- dispatcher ------------------------------------------------------------------------------------
void gemm(Ctx& c, const GemmP& g) {
  if (!c.ok()) return;
  if (c.dtype == SJ_BF16 && g.W_tc && g.amode != A_CONV3 && !g.am.map) {
    const bool ln = g.ln_mean != nullptr;
    // a folded tensor-core copy can only serve the LayerNorm'ed use of the layer, and vice versa
    if (ln == (g.tc_colsum != nullptr)) {
      TcGemmP t;
      bool ok = true;
      t.A = g.A; t.lda = g.lda;
      if (g.amode == A_MERGE) {
        t.a_mode = 2; t.mH = g.H; t.mW = g.Wd; t.mC = g.Cin;
      } else if (g.am.inner > 0) {
        // rows [outer][G][inner] visited per group: outer stride must be G*inner and group stride inner
        ok = g.am.outer == g.groups * g.am.inner && g.am.gstride == g.am.inner;
        t.a_mode = 1; t.a_inner = g.am.inner;
      } else {
        ok = g.am.gstride == 0;
        t.a_mode = 0;
      }
      t.Bw = g.W_tc; t.M = g.M; t.N = g.N; t.K = g.K; t.groups = g.groups;
      t.bias = ln ? g.tc_bias : g.bias; t.bias_gstride = g.bias_gstride; t.act = g.act;
      t.R = g.R; t.ldr = g.ldr; t.C = g.C; t.ldc = g.ldc; t.cm = g.cm;
      t.st_mean = g.st_mean; t.st_rstd = g.st_rstd; t.st_eps = g.st_eps;
      if (ln) {
        t.ln_mean = g.ln_mean; t.ln_rstd = g.ln_rstd; t.ln_s = g.tc_colsum;
        t.ln_gstride = g.groups > 1 ? g.N : 0;
        if (g.groups > 1) t.bias_gstride = g.N;
      }
      if (ok && (g.groups == 1 || g.w_gstride == (long long)g.K * g.N) && tc_gemm_supported(t)) {
        tc_gemm(c, t);
        return;
      }
    }
  }
  if (g.st_mean) { c.fail(SJ_EUNSUPPORTED); return; }  // fused output statistics exist on the tensor-core path only
  gemm_simt(c, g);
}

}  // namespace sj
