// Last decoder stage fused with the head's channel contraction (K8 + the dense half of K10):
//   x4 = ELU(conv3x3_SAME(nearest_up2(x3)) + b)          96 -> 48 @128^2 -> 256^2       (modules.py:746-749, :732-737)
//   Z[q, tap*2 + o] = sum_c x4[q, c] * Wout[tap, c, o]    the 48 -> 2 heads (modules.py:767-770) re-associated as a pointwise
//                                                         projection to 18 columns followed by a 9-tap shifted sum
// x4 (755 MB per launch at batch 16) never leaves the SM: only Z (fp16, 36 B per output pixel instead of 96) is written,
// and head_tapsum (tc_headsum.cu) finishes  out[p, o] = b[o] + sum_tap Z[p + off(tap), tap*2 + o].
//
// The up-convolution is tc_upconv4.cu's scheme (four sub-pixel phases per CTA, 9 shifted UMMA views of one TMA-staged
// patch, phases stacked along N, 144 KB of folded weights resident).  What is new:
//   * the projection runs on the tensor cores with its A operand in TENSOR MEMORY (tcgen05.mma "TS" form): the epilogue
//     threads read their accumulator row, apply bias + ELU, pack to bf16 pairs and store them back over the SAME TMEM
//     columns (tcgen05.st); three K = 16 MMAs per phase (N = 32, 18 real columns) then write Z into the 128 TMEM columns
//     the double-buffered accumulators leave free.  Measured on B200 (tools/mma_probe.cu, profiles/r02_mma_probe.txt) a
//     TS-form MMA costs N/2 cycles flat (16 here) where the SS form costs max(N/2, 32 + N/4): the whole projection is 192
//     tensor cycles per tile next to ~3100 for the convolution, and needs no shared memory for x4;
//   * ONE warp issues both MMA streams, in an order that keeps the pipe full: conv chunk 0, chunk 1 of tile j, then the
//     projection of tile j-1 (its operand was produced by the epilogue warps while chunks 0-1 ran), then chunk 2;
//   * the input patch streams through a 4-slot ring of 32-channel chunks (48 KB instead of 2 x 36 KB) to make room;
//   * Z rows leave through a per-warp shared-memory transpose so that every global store instruction writes 512
//     contiguous bytes (the un-fused kernel's 16-byte pieces at 192-byte stride were half-sector writes).
// Warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warps 2..9 = epilogue (2..5: output rows 2y, 6..9: rows 2y+1).
#include <cuda_fp16.h>

#include <cstdio>

#include "kernels.h"
#include "tc_common.cuh"

namespace sj {
namespace {

using namespace tc;

constexpr int CIN = 96, COUT = 48, KC = 32, NCH = CIN / KC;
constexpr int TH = 16, TW = 8, PH = TH + 2, PW = TW + 2;
constexpr int NTHREADS = 320;
constexpr int A_SUB = (PH * PW * KC * 2 + 1023) & ~1023;  // one 32-channel chunk of the patch: 11520 -> 12288 B
constexpr int NSLOT = 4;
constexpr int B_TILE = COUT * KC * 2;                     // 3072 B
constexpr int B_BYTES = NCH * 16 * B_TILE;                // 144 KB
constexpr int W2_BYTES = 32 * 128;                        // head weights: 32 rows (18 real) x 64 channels (48 real), SWIZZLE_128B
constexpr int ZCH = 18;                                   // Z columns per pixel
constexpr int STG_ROW = TW * 2 * ZCH * 2;                 // 576 B: one output row of a tile (16 pixels x 18 fp16)
constexpr int STG_WARP = 4 * STG_ROW;                     // 2304 B per epilogue warp
constexpr int OFF_B = NSLOT * A_SUB;
constexpr int OFF_W2 = OFF_B + B_BYTES;
constexpr int OFF_STG = OFF_W2 + W2_BYTES;
constexpr int OFF_BAR = OFF_STG + 8 * STG_WARP;
constexpr int SMEM_BYTES = OFF_BAR + 1024;
static_assert(OFF_W2 % 1024 == 0, "head weight tile alignment (128-byte swizzle)");
static_assert(SMEM_BYTES + 1024 <= 227 * 1024, "shared memory");
constexpr int ACC_COLS = 4 * COUT;                        // 192 TMEM columns per accumulator stage
constexpr int Z_COL0 = 2 * ACC_COLS;                      // Z: 4 phases x 32 columns at [384, 512)

struct Slot { int ro, dx, ring; };
__host__ __device__ constexpr int ring_py(int r) { return r >> 1; }
__host__ __device__ constexpr int ring_px(int r) { return (r == 1 || r == 2) ? 1 : 0; }

struct Up4hP {
  int NB, H, W, tiles_x, tiles_y, num_tiles;
  const float* bias;
  __half* z;  // [NB, 2H, 2W, 18]
};

// D[tmem] (+)= A[tmem] . B[smem]^T: A = 128 lanes x 16 bf16 (8 packed 32-bit columns)
__device__ __forceinline__ void umma_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::
          "r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float (&v)[2]) {
  uint32_t r0, r1;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  v[0] = __uint_as_float(r0);
  v[1] = __uint_as_float(r1);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  __half2 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

__global__ void __launch_bounds__(NTHREADS, 1)
tc_upconv4h_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                   const __grid_constant__ CUtensorMap mapW2, const Up4hP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem + OFF_B;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* afull = bars;                // [NSLOT]
  uint64_t* aempty = bars + NSLOT;       // [NSLOT]
  uint64_t* bfull = bars + 2 * NSLOT;    // weights resident
  uint64_t* tfull = bfull + 1;           // [2] conv accumulator of a tile complete
  uint64_t* tempty = tfull + 2;          // [2] ... and consumed (projection MMAs of that tile complete)
  uint64_t* a2full = tempty + 2;         // [2] bf16 x4 rows of a tile are in TMEM (8 epilogue warps)
  uint64_t* zfull = a2full + 2;          // Z of a tile complete
  uint64_t* zempty = zfull + 1;          // ... and read out (8 epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(zempty + 1);
  float* bias_s = reinterpret_cast<float*>(bars + 32);  // [COUT]

  const int warp = uniform_warp_idx(), lane = threadIdx.x % 32;
  for (int i = threadIdx.x; i < COUT; i += NTHREADS) bias_s[i] = p.bias[i];
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    prefetch_tmap(&mapW2);
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(&afull[s], 1);
      mbar_init(&aempty[s], 1);
    }
    mbar_init(bfull, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 1);
      mbar_init(&a2full[a], 8);
    }
    mbar_init(zfull, 1);
    mbar_init(zempty, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  if (warp != 0) {  // the producer lane waits after it has issued the (constant) resident weights
    pdl_wait();
    pdl_trigger();
  }

  constexpr Slot SLOTS[16] = {{1, 1, 0}, {1, 1, 1}, {1, 1, 2}, {1, 1, 3}, {0, 1, 0}, {0, 1, 1}, {1, 2, 1}, {1, 2, 2},
                              {2, 1, 2}, {2, 1, 3}, {1, 0, 0}, {1, 0, 3}, {0, 0, 0}, {0, 2, 1}, {2, 2, 2}, {2, 0, 3}};

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bfull, B_BYTES + W2_BYTES);
      for (int ch = 0; ch < NCH; ++ch)
        for (int s = 0; s < 16; ++s) {
          const int py = ring_py(SLOTS[s].ring), px = ring_px(SLOTS[s].ring);
          const int a = SLOTS[s].ro - py, b = SLOTS[s].dx - px;
          tma_load_2d(smem_b + (ch * 16 + s) * B_TILE, &mapB, bfull, (a * 2 + b) * CIN + ch * KC, (py * 2 + px) * COUT);
        }
      tma_load_2d(smem + OFF_W2, &mapW2, bfull, 0, 0);
      pdl_wait();
      int slot = 0;
      uint32_t sph = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int n = t / tiles_per_img, tr = t % tiles_per_img;
        const int y0 = (tr / p.tiles_x) * TH, x0 = (tr % p.tiles_x) * TW;
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
          mbar_wait(&aempty[slot], sph ^ 1);
          mbar_expect_tx(&afull[slot], PH * PW * KC * 2);
          tma_load_4d(smem + slot * A_SUB, &mapA, &afull[slot], ch * KC, x0 - 1, y0 - 1, n);
          if (++slot == NSLOT) { slot = 0; sph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    struct Op { int ro, dx, slot, ring, cnt; };
    constexpr Op OPS[10] = {{1, 1, 0, 0, 4}, {0, 1, 4, 0, 2}, {1, 2, 6, 1, 2}, {2, 1, 8, 2, 2}, {1, 0, 10, 0, 1},
                            {1, 0, 11, 3, 1}, {0, 0, 12, 0, 1}, {0, 2, 13, 1, 1}, {2, 2, 14, 2, 1}, {2, 0, 15, 3, 1}};
    constexpr uint32_t A_HI = desc_hi(KC * 2, PW * KC * 2), B_HI = desc_hi(KC * 2, 8 * KC * 2), W2_HI = desc_hi(128, 1024);
    const uint32_t idesc1 = make_idesc_bf16(128, COUT), idesc2 = make_idesc_bf16(128, 2 * COUT),
                   idesc4 = make_idesc_bf16(128, 4 * COUT), idesc_z = make_idesc_bf16(128, 32);
    const uint32_t b_lo = desc_lo(smem_u32(smem_b)), w2_lo = desc_lo(smem_u32(smem + OFF_W2));
    // projection of local tile j (accumulator stage j & 1): A = the bf16 rows the epilogue stored over the accumulator
    auto project = [&](int j) {
      const int acc = j & 1;
      mbar_wait(&a2full[acc], (uint32_t)(j >> 1) & 1);
      mbar_wait(zempty, ((uint32_t)j & 1) ^ 1);  // Z of tile j-1 has been read out
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ring = 0; ring < 4; ++ring)
#pragma unroll
          for (int k = 0; k < 3; ++k)
            umma_ts_bf16(tmem_base + Z_COL0 + ring * 32, tmem_base + acc * ACC_COLS + ring * COUT + 8 * k, w2_lo + 2 * k,
                         W2_HI, idesc_z, k != 0);
        umma_commit(zfull);
        umma_commit(&tempty[acc]);
      }
      __syncwarp();
    };
    int slot = 0, j = 0;
    uint32_t sph = 0;
    mbar_wait(bfull, 0);
    tc_fence_after();
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++j) {
      const int acc = j & 1;
      mbar_wait(&tempty[acc], ((uint32_t)(j >> 1) & 1) ^ 1);
      const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
#pragma unroll 1
      for (int ch = 0; ch < NCH; ++ch) {
        mbar_wait(&afull[slot], sph);
        tc_fence_after();
        const uint32_t a_lo = desc_lo(smem_u32(smem + slot * A_SUB));
        const uint32_t bc_lo = b_lo + ((ch * 16 * B_TILE) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) {
#pragma unroll
            for (int o = 0; o < 10; ++o) {
              const uint32_t va = a_lo + (((OPS[o].ro * PW + OPS[o].dx) * KC * 2) >> 4) + 2 * k;
              const uint32_t vb = bc_lo + ((OPS[o].slot * B_TILE) >> 4) + 2 * k;
              const uint32_t idesc = OPS[o].cnt == 4 ? idesc4 : (OPS[o].cnt == 2 ? idesc2 : idesc1);
              // the centre view covers all 192 columns and is issued first: it alone initialises the accumulators
              umma_bf16_w(d_tmem + OPS[o].ring * COUT, va, A_HI, vb, B_HI, idesc, (o != 0 || (ch | k) != 0) ? 1u : 0u);
            }
          }
          umma_commit(&aempty[slot]);
          if (ch == NCH - 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++slot == NSLOT) { slot = 0; sph ^= 1; }
        if (ch == 1 && j > 0) project(j - 1);
      }
    }
    if (j > 0) project(j - 1);
  } else {
    const int quarter = warp % 4, py = (warp - 2) / 4, ew = warp - 2;
    const int r = quarter * 32 + lane, ty = r / TW, tx = r % TW;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    uint8_t* stg = smem + OFF_STG + ew * STG_WARP;
    float bias_r[COUT];
#pragma unroll
    for (int i = 0; i < COUT; ++i) bias_r[i] = bias_s[i];

    // Z of local tile jz (global tile tz): TMEM -> fp16 -> per-warp transpose -> 512-byte contiguous global stores
    auto read_out = [&](int jz, int tz) {
      mbar_wait(zfull, (uint32_t)jz & 1);
      tc_fence_after();
      uint32_t zq[2][9];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint32_t z_addr = tmem_base + lane_addr + Z_COL0 + (py * 2 + half) * 32;
        float v[16], w[2];
        tmem_ld16(z_addr, v);
        tmem_ld2(z_addr + 16, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) zq[half][i] = pack_f16(v[2 * i], v[2 * i + 1]);
        zq[half][8] = pack_f16(w[0], w[1]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(zempty);
      // ring order: row phase 0 holds (px0, px1) in halves (0, 1), row phase 1 holds (px1, px0)
      {
        uint2* d = reinterpret_cast<uint2*>(stg + (lane / TW) * STG_ROW + tx * (2 * ZCH * 2));
        uint32_t w18[18];  // px = 0 sits in half `py`, px = 1 in the other (selects, not indexed: stays in registers)
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          w18[i] = py ? zq[1][i] : zq[0][i];
          w18[9 + i] = py ? zq[0][i] : zq[1][i];
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) d[i] = make_uint2(w18[2 * i], w18[2 * i + 1]);
      }
      __syncwarp();
      const int n = tz / tiles_per_img, tr = tz % tiles_per_img;
      const int y0 = (tr / p.tiles_x) * TH, x0 = (tr % p.tiles_x) * TW;
      uint8_t* zg = reinterpret_cast<uint8_t*>(p.z);
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const int c = lane + 32 * i;
        if (c < 4 * (STG_ROW / 16)) {
          const int row = c / (STG_ROW / 16), col = c % (STG_ROW / 16);
          const uint4 v = *reinterpret_cast<const uint4*>(stg + row * STG_ROW + col * 16);
          const long long Y = 2 * (y0 + quarter * 4 + row) + py;
          *reinterpret_cast<uint4*>(zg + (((long long)n * (2 * p.H) + Y) * (2 * p.W) + 2 * x0) * (ZCH * 2) + col * 16) = v;
        }
      }
      __syncwarp();
    };

    int j = 0, t_prev = -1;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++j) {
      const int acc = j & 1;
      mbar_wait(&tfull[acc], (uint32_t)(j >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint32_t t_addr = tmem_base + lane_addr + acc * ACC_COLS + (py * 2 + half) * COUT;
        uint32_t q[24];
        {
          float v[32];
          tmem_ld32(t_addr, v);
#pragma unroll
          for (int i = 0; i < 32; i += 2)
            q[i >> 1] = pack_bf16(act_fast(v[i] + bias_r[i], ACT_ELU), act_fast(v[i + 1] + bias_r[i + 1], ACT_ELU));
        }
        {
          float v[16];
          tmem_ld16(t_addr + 32, v);
#pragma unroll
          for (int i = 0; i < 16; i += 2)
            q[16 + (i >> 1)] =
                pack_bf16(act_fast(v[i] + bias_r[32 + i], ACT_ELU), act_fast(v[i + 1] + bias_r[32 + i + 1], ACT_ELU));
        }
        // bf16 x4 row back over the accumulator's own columns: channels (2c, 2c+1) in column c of this phase block
        tmem_st16(t_addr, q);
        tmem_st8(t_addr + 16, q + 16);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a2full[acc]);
      if (j > 0) read_out(j - 1, t_prev);
      t_prev = t;
    }
    if (j > 0) read_out(j - 1, t_prev);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

bool tc_upconv4h_supported(int H, int W, int Cin, int Cout) {
  return Cin == CIN && Cout == COUT && H % TH == 0 && W % TW == 0;
}

// x bf16 [NB,H,W,96] -> z fp16 [NB,2H,2W,18]; w_tc = folded kernels [4 phases][48][4*96] bf16;
// head_w_tc = one head of SjDecoderW.out_w_tc: bf16 [32 rows = tap*2+o (18 real)][64 channels (48 real)]
void tc_upconv4h(Ctx& c, const void* x, void* z, const void* w_tc, const float* bias, const void* head_w_tc, int NB, int H,
                 int W) {
  if (!c.ok() || c.dry) return;
  if (!w_tc || !bias || !head_w_tc || H % TH || W % TW) { c.fail(SJ_EUNSUPPORTED); return; }
  Up4hP p{};
  p.NB = NB; p.H = H; p.W = W;
  p.tiles_x = W / TW; p.tiles_y = H / TH;
  p.num_tiles = NB * p.tiles_x * p.tiles_y;
  p.bias = bias;
  p.z = (__half*)z;
  CUtensorMap mapA, mapB, mapW2;
  uint64_t da[4] = {(uint64_t)CIN, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
  uint64_t sa[3] = {(uint64_t)CIN * 2, (uint64_t)W * CIN * 2, (uint64_t)H * W * CIN * 2};
  uint32_t ba[4] = {KC, PW, PH, 1};
  uint64_t db[2] = {(uint64_t)4 * CIN, (uint64_t)4 * COUT};
  uint64_t sb[1] = {(uint64_t)4 * CIN * 2};
  uint32_t bb[2] = {KC, COUT};
  uint64_t dw[2] = {64, 32};
  uint64_t sw[1] = {128};
  uint32_t bw[2] = {64, 32};
  if (!encode_tmap(&mapA, x, 4, da, sa, ba, KC * 2) || !encode_tmap(&mapB, w_tc, 2, db, sb, bb, KC * 2) ||
      !encode_tmap(&mapW2, head_w_tc, 2, dw, sw, bw, 128)) {
    snprintf(tls().cuda_err, sizeof(tls().cuda_err), "cuTensorMapEncodeTiled failed (tc_upconv4h)");
    c.fail(SJ_ECUDA);
    return;
  }
  const size_t smem = 1024 + SMEM_BYTES;
  if (!SJ_SMEM_LIMIT_OK(tc_upconv4h_kernel, 227 * 1024)) { c.fail(SJ_ECUDA); return; }
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  SJ_LAUNCH(c, "tc_upconv4h", tc_upconv4h_kernel, grid, NTHREADS, smem, mapA, mapB, mapW2, p);
}

}  // namespace sj
