// Hardware probe (not product code): tcgen05.mma issue cost on B200 as a function of N, operand swizzle, operand source
// (A from shared memory vs from TMEM), descriptor alignment (shifted views of a TMA patch) and cta_group, plus a
// functional check of the TS form (A operand written to TMEM by tcgen05.st as packed bf16x2).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I strajnet_b200/csrc -o tools/bin/mma_probe tools/mma_probe.cu
//   tools/bin/mma_probe > gpurun_out/mma_probe.txt
//
// Method: one warp issues n back-to-back MMAs into the same accumulator, commits, waits; cycles per MMA =
// (T(n = 576) - T(n = 64)) / 512 so that the fixed launch / drain latency cancels.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tc_common.cuh"

using namespace sj;
using namespace sj::tc;

namespace sj {  // tc_common.cuh declares these; the probe links nothing else
bool encode_tmap(CUtensorMap*, const void*, int, const uint64_t*, const uint64_t*, const uint32_t*, int) { return false; }
}

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                        uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d),
      "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma2_bf16_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}


// One step = a compile-time list of MMAs, fully unrolled (descriptors fold to "base + constant", as in the product
// kernels); the issuing warp runs the loop uniformly and one elected lane issues.
//   PAT 0: one MMA of N columns per step, operands advance through the K16 slices of a row
//   PAT 1: tc_upconv4's 10-MMA step (N = 192, 3 x 96, 6 x 48 over 9 shifted views), 64-byte rows
//   PAT 2: the same work as 16 N = 48 MMAs
//   PAT 3: the same work as 4 N = 192 MMAs (one per tap, phases stacked)
//   PAT 4: PAT 0 alternating between two accumulators
struct Op { int ro, dx, slot, ring, cnt; };
template <int PAT, int N, int RB, int TS, int CTA2>
__device__ __forceinline__ void issue_step(uint32_t tmem, uint32_t a_lo, uint32_t b_lo, int s) {
  constexpr uint32_t M = CTA2 ? 256 : 128;
  auto mma = [&](uint32_t d, uint32_t a, uint32_t ahi, uint32_t b, uint32_t bhi, uint32_t idesc) {
    if constexpr (CTA2) umma2_bf16_w(d, a, ahi, b, bhi, idesc, 1u);
    else if constexpr (TS) umma_ts(d, tmem + 384 + (a & 31), b, bhi, idesc, 1u);
    else umma_bf16_w(d, a, ahi, b, bhi, idesc, 1u);
  };
  if constexpr (PAT == 0 || PAT == 4) {
    constexpr uint32_t HI = desc_hi(RB, 8 * RB);
    const uint32_t idesc = make_idesc_bf16(M, N);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t k = 2 * (i % (RB / 32));
      mma(tmem + (PAT == 4 ? (i & 1) * 256 : 0), a_lo + (TS ? 4 * k : k), HI, b_lo + k, HI, idesc);
    }
  } else {
    constexpr uint32_t A_HI = desc_hi(64, 640), B_HI = desc_hi(64, 512);
    constexpr int B_TILE = 3072 / (CTA2 ? 2 : 1);  // cta_group::2: each CTA holds half of the N rows
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if constexpr (PAT == 1) {
        constexpr Op OPS[10] = {{1, 1, 0, 0, 4}, {0, 1, 4, 0, 2}, {1, 2, 6, 1, 2}, {2, 1, 8, 2, 2}, {1, 0, 10, 0, 1},
                                {1, 0, 11, 3, 1}, {0, 0, 12, 0, 1}, {0, 2, 13, 1, 1}, {2, 2, 14, 2, 1}, {2, 0, 15, 3, 1}};
#pragma unroll
        for (int o = 0; o < 10; ++o)
          mma(tmem + OPS[o].ring * 48, a_lo + (((OPS[o].ro * 10 + OPS[o].dx) * 64) >> 4) + 2 * k, A_HI,
              b_lo + ((OPS[o].slot * B_TILE) >> 4) + 2 * k, B_HI, make_idesc_bf16(M, 48 * OPS[o].cnt));
      } else if constexpr (PAT == 2) {
#pragma unroll
        for (int o = 0; o < 16; ++o)
          mma(tmem + (o % 4) * 48, a_lo + ((((o / 4) % 3 * 10 + o % 3) * 64) >> 4) + 2 * k, A_HI,
              b_lo + ((o * B_TILE) >> 4) + 2 * k, B_HI, make_idesc_bf16(M, 48));
      } else {
#pragma unroll
        for (int o = 0; o < 4; ++o)
          mma(tmem, a_lo + ((((o / 2) * 10 + o % 2) * 64) >> 4) + 2 * k, A_HI, b_lo + ((o * 4 * B_TILE) >> 4) + 2 * k, B_HI,
              make_idesc_bf16(M, 192));
      }
    }
  }
}
template <int PAT>
constexpr int mmas_per_step() { return PAT == 0 || PAT == 4 ? 8 : (PAT == 1 ? 20 : (PAT == 2 ? 32 : 8)); }

__device__ int g_random_data = 0;
__device__ int g_extras = 0;  // bit 0: tcgen05.commit after every step (20 MMAs of PAT 1); bit 1: + tcgen05.fence::after_thread_sync;
                              // bit 2: + an mbarrier wait on that commit every 3rd step (the kernel's per-tile hand-off)  // 1: operands = random bf16 bit patterns in [-2, 2] (tensor-core power as in real kernels)
template <int PAT, int N, int RB, int TS, int CTA2>
__device__ __forceinline__ void probe_body(int steps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (i + 1) * 2654435761u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    // random: sign + mantissa random, exponent in {0x3e, 0x3f} (|x| in [0.125, 2)); constant-ish: ~0.0078 with 3 low bits varying
    reinterpret_cast<uint32_t*>(smem)[i] = g_random_data ? ((h & 0x80ff80ffu) | 0x3e003e00u | ((h >> 3) & 0x01000100u))
                                                         : (0x3c003c00u ^ ((i * 2654435761u) & 0x00700070u));
  }
  const int warp = uniform_warp_idx();
  uint32_t rank = 0;
  if constexpr (CTA2) rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&bar2, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    if constexpr (CTA2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc(&tmem_slot, 512);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const uint32_t a_lo = desc_lo(smem_u32(smem)), b_lo = desc_lo(smem_u32(smem + 64 * 1024));
    const long long t0 = clock64();
    if (rank == 0) {
      if (elect_one()) {
        const int extras = g_extras;
        uint32_t xph = 0;
#pragma unroll 1
        for (int s = 0; s < steps; ++s) {
          issue_step<PAT, N, RB, TS, CTA2>(tmem, a_lo, b_lo, s);
          if (extras & 1) {
            if constexpr (!CTA2) umma_commit(&bar2);
            if (extras & 2) tc_fence_after();
            if ((extras & 4) && s % 3 == 2) {
              // wait until everything issued so far has completed (phase parity of the commit just issued)
              // each commit completes one phase of bar2 (count 1): the commit of step s completes phase s
              uint32_t done = 0;
              while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&bar2)), "r"((uint32_t)(s & 1)) : "memory");
            }
          }
        }
        (void)xph;
        if constexpr (CTA2)
          asm volatile(
              "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                  smem_u32(&bar)),
              "h"((uint16_t)3)
              : "memory");
        else
          umma_commit(&bar);
      }
      __syncwarp();
    }
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if constexpr (CTA2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    else tmem_dealloc(tmem, 512);
  }
}
template <int PAT, int N, int RB, int TS>
__global__ void __launch_bounds__(128, 1) probe1_kernel(int steps, long long* out) {
  probe_body<PAT, N, RB, TS, 0>(steps, out);
}
template <int PAT, int N, int RB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe2_kernel(int steps, long long* out) {
  probe_body<PAT, N, RB, 0, 1>(steps, out);
}

// ---- functional check of the TS form ---------------------------------------------------------------------------
// D[128 x 32] = A[128 x 48] . B[32 x 48]^T with A written to TMEM by the 128 row threads (tcgen05.st 32x32b, bf16x2 packed:
// column c of lane m holds A[m][2c] (low half) and A[m][2c+1]), B in shared memory as a SWIZZLE_128B K-major tile.
__global__ void __launch_bounds__(128, 1) ts_check_kernel(const bf16* A, const bf16* B, float* D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int i = threadIdx.x; i < 32 * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  // B tile: row n (0..31), 48 bf16 = 6 chunks of 16 B; swizzle: chunk ^= (row % 8)
  for (int i = threadIdx.x; i < 32 * 6; i += blockDim.x) {
    const int n = i / 6, ch = i % 6;
    const uint4 v = *reinterpret_cast<const uint4*>(B + n * 48 + ch * 8);
    *reinterpret_cast<uint4*>(smem + (n / 8) * 1024 + (n % 8) * 128 + ((ch ^ (n % 8)) * 16)) = v;
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // A -> TMEM columns [64, 88): thread = row
  {
    const int m = threadIdx.x;
    uint32_t r[24];
    for (int c = 0; c < 24; ++c) r[c] = *reinterpret_cast<const uint32_t*>(A + m * 48 + 2 * c);
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 64;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
        "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
        : "memory");
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16};" ::"r"(taddr + 8),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    const uint32_t b_hi = desc_hi(128, 1024), b_lo = desc_lo(smem_u32(smem));
    const uint32_t idesc = make_idesc_bf16(128, 32);
    if (elect_one()) {
      for (int k = 0; k < 3; ++k) umma_ts(tmem, tmem + 64 + 8 * k, b_lo + 2 * k, b_hi, idesc, k != 0);
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
    for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * 32 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

#define CK(x)                                                                             \
  do {                                                                                    \
    cudaError_t e_ = (x);                                                                 \
    if (e_ != cudaSuccess) {                                                              \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
      exit(1);                                                                            \
    }                                                                                     \
  } while (0)


static long long* d_out;
constexpr int SMEM = 161 * 1024;
// cycles per MMA: (T(steps = 72) - T(steps = 8)) / (64 * MMAs per step), averaged over the CTAs (pairs) of the grid
template <typename K>
static double run(K kernel, int per_step, bool two_cta, int grid) {
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  double t[2];
  const int steps[2] = {8, 72};
  for (int pass = 0; pass < 2; ++pass) {
    std::vector<long long> h(grid);
    double best = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
      kernel<<<grid, 128, SMEM>>>(steps[pass], d_out);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h.data(), d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
      double m = 0;
      int cnt = 0;
      for (int i = 0; i < grid; i += (two_cta ? 2 : 1), ++cnt) m += (double)h[i];
      m /= cnt;
      if (m < best) best = m;
    }
    t[pass] = best;
  }
  return (t[1] - t[0]) / (64.0 * per_step);
}

template <int N>
static void row_n() {
  printf("N=%-3d floor %5.1f | SS sw128 %6.1f (148 CTAs %6.1f) | SS sw64 %6.1f | TS %6.1f | 2 accumulators %6.1f | cta_group::2 sw128 %6.1f (74 pairs %6.1f) sw64 %6.1f\n",
         N, N / 2.0, run(probe1_kernel<0, N, 128, 0>, 8, false, 1), run(probe1_kernel<0, N, 128, 0>, 8, false, 148),
         run(probe1_kernel<0, N, 64, 0>, 8, false, 1), run(probe1_kernel<0, N, 128, 1>, 8, false, 1),
         run(probe1_kernel<4, N, 128, 0>, 8, false, 1), run(probe2_kernel<0, N, 128>, 8, true, 2),
         run(probe2_kernel<0, N, 128>, 8, true, 148), run(probe2_kernel<0, N, 64>, 8, true, 2));
  fflush(stdout);
}

int main(int argc, char** argv) {
  const char* sec = argc > 1 ? argv[1] : "all";
  auto on = [&](const char* n) { return !strcmp(sec, "all") || !strcmp(sec, n); };
  CK(cudaMalloc(&d_out, 4096 * sizeof(long long)));
  if (on("n")) {
    printf("# cycles per tcgen05.mma (M = 128 per CTA, K = 16, bf16, back to back into one accumulator); floor = N/2\n");
    row_n<16>(); row_n<32>(); row_n<48>(); row_n<64>(); row_n<96>(); row_n<128>(); row_n<192>(); row_n<256>();
  }
  if (on("pattern")) {
    printf("# one 16-channel step pair of tc_upconv4 (sum of N = 768 per K16 slice, floor 384 cycles per slice)\n");
    printf("10 stacked MMAs (product kernel): %7.1f cycles per K16 slice (148 CTAs %7.1f)\n",
           10 * run(probe1_kernel<1, 0, 64, 0>, 20, false, 1), 10 * run(probe1_kernel<1, 0, 64, 0>, 20, false, 148));
    printf("16 MMAs of N = 48:                %7.1f\n", 16 * run(probe1_kernel<2, 0, 64, 0>, 32, false, 1));
    printf("4 MMAs of N = 192:                %7.1f\n", 4 * run(probe1_kernel<3, 0, 64, 0>, 8, false, 1));
    printf("cta_group::2, 10 stacked MMAs:    %7.1f (74 pairs %7.1f)\n", 10 * run(probe2_kernel<1, 0, 64>, 20, true, 2),
           10 * run(probe2_kernel<1, 0, 64>, 20, true, 148));
    printf("cta_group::2, 16 MMAs of N = 48:  %7.1f\n", 16 * run(probe2_kernel<2, 0, 64>, 32, true, 2));
    printf("cta_group::2, 4 MMAs of N = 192:  %7.1f\n", 4 * run(probe2_kernel<3, 0, 64>, 8, true, 2));
  }
  if (on("commit")) {
    printf("# cost of tcgen05.commit / fence / a full hand-off inside the 10-MMA pattern stream (cycles per K16 slice, floor 384, plain 515)\n");
    for (int ex : {0, 1, 3, 5, 7}) {
      CK(cudaMemcpyToSymbol(g_extras, &ex, sizeof(int)));
      printf("extras=%d (%s%s%s): %7.1f\n", ex, ex & 1 ? "commit per 20 MMAs" : "none", ex & 2 ? " + fence" : "",
             ex & 4 ? " + wait for completion every 60 MMAs" : "", 10 * run(probe1_kernel<1, 0, 64, 0>, 20, false, 1));
    }
    int zero = 0;
    CK(cudaMemcpyToSymbol(g_extras, &zero, sizeof(int)));
  }
  if (on("power")) {
    // steady-state cost under load: the 10-MMA pattern on all 148 SMs for ~10 ms per launch, 12 launches back to back,
    // with near-constant operands (low toggle rate) and with random operands (what a real layer feeds the tensor cores)
    auto kernel = probe1_kernel<1, 0, 64, 0>;
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    for (int rnd = 0; rnd < 2; ++rnd) {
      CK(cudaMemcpyToSymbol(g_random_data, &rnd, sizeof(int)));
      const int steps = 20000;
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0));
      CK(cudaEventCreate(&e1));
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        for (int l = 0; l < 12; ++l) kernel<<<148, 128, SMEM>>>(steps, d_out);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        std::vector<long long> h(148);
        CK(cudaMemcpy(h.data(), d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
        double cyc = 0;
        for (int i = 0; i < 148; ++i) cyc += (double)h[i] / 148;
        printf("%s operands: %.1f cycles per K16 slice (10 MMAs), last launch %.0f cycles, 12 launches in %.2f ms -> SM clock ~%.0f MHz\n",
               rnd ? "random  " : "constant", cyc / (steps * 2.0), cyc, ms, cyc * 12 / (ms * 1e3));
      }
    }
    int zero = 0;
    CK(cudaMemcpyToSymbol(g_random_data, &zero, sizeof(int)));
  }
  // ---- TS functional check
  if (on("tscheck")) {
    std::vector<bf16> hA(128 * 48), hB(32 * 48);
    std::vector<float> fA(128 * 48), fB(32 * 48), hD(128 * 32);
    unsigned s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 32768.0f - 1.0f; };
    for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16(rnd()); fA[i] = __bfloat162float(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16(rnd()); fB[i] = __bfloat162float(hB[i]); }
    bf16 *dA, *dB;
    float* dD;
    CK(cudaMalloc(&dA, hA.size() * 2));
    CK(cudaMalloc(&dB, hB.size() * 2));
    CK(cudaMalloc(&dD, hD.size() * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    ts_check_kernel<<<1, 128, 8 * 1024>>>(dA, dB, dD);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 32; ++n) {
        double ref = 0;
        for (int k = 0; k < 48; ++k) ref += (double)fA[m * 48 + k] * fB[n * 48 + k];
        maxerr = fmax(maxerr, fabs(ref - hD[m * 32 + n]));
      }
    printf("TS functional check (A from TMEM via tcgen05.st, K = 48, N = 32): max abs err %.3e %s\n", maxerr,
           maxerr < 1e-4 ? "OK" : "MISMATCH");
  }
  return 0;
}
