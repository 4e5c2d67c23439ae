// Decoder up-convolution on the tensor cores (K8): nearest x2 upsample + 3x3 SAME conv + bias + ELU
// (modules.py:746-749), executed as four 2x2 sub-pixel convolutions on the LOW-RES input with
// pre-summed taps (weights.fold_upconv_subpixel): 2.25x fewer FLOPs, 4x fewer input bytes.
//
// Implicit GEMM without an im2col buffer and without re-reading the input per tap:
//   * work item = (image, 16x8 low-res pixel tile, output row phase py); both column phases px are
//     produced in two TMEM accumulators [128 pixels x Cout];
//   * per Cin chunk ONE 4-D TMA box {channels, 10, 17, 1} stages the tile plus its halo (rows y0-1+py..,
//     columns x0-1..x0+8) in shared memory; out-of-image rows/columns are zero-filled by TMA, which
//     is exactly the SAME padding of the reference conv;
//   * each of the 2x3 taps is a *shifted view* of that one patch: the UMMA descriptor starts
//     (a*10 + dx+1) pixel rows into the patch and uses stride_byte_offset = 10 pixel rows, so MMA row
//     r = 8*ty+tx reads patch pixel (ty+a, tx+dx+1).  The 128B/64B swizzle is a function of the
//     absolute shared-memory address (verified on B200 by tools/probe_shift.py), so views that are
//     not aligned to the swizzle atom read exactly what TMA wrote;
//   * folded weights [4 phases][Cout][4*Cin] either stream through an mbarrier ring (large layers) or,
//     when the 8 tiles of one row phase fit (the 96->48 layer), stay resident in shared memory
//     for the whole persistent CTA;
//   * warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warps 2..5 = epilogue (bias + ELU + bf16,
//     16-byte stores); accumulators double-buffered in TMEM when 4*Cout fits 512 columns.
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "tc_common.cuh"

namespace sj {
namespace {

using namespace tc;

constexpr int TH = 16, TW = 8, PH = TH + 1, PW = TW + 2, BM = 128;
constexpr int NTHREADS = 352;  // warp 0 TMA, warps 1 and 10 MMA issuers (one per column phase), warps 2..9 epilogue
constexpr int MAX_STAGES = 8;

struct ConvP {
  int NB, H, W, Cin, Cout;  // low-res input geometry
  int tiles_x, tiles_y, num_tiles;
  int acs;                  // TMEM column stride between accumulators
  int nacc;                 // accumulator stages (1 or 2)
  int na, nbs;              // A-patch slots, B stages (streaming mode)
  int stack;                // dx = 0 tap as one N = 2*Cout MMA (accumulators adjacent: acs == Cout)
  int a_slot;               // bytes per A slot (all chunks of a group, each 1024-aligned)
  int a_sub;                // bytes per chunk sub-patch (padded to 1024)
  const float* bias;
  bf16* out;                // [NB, 2H, 2W, Cout]
};

__device__ __forceinline__ uint64_t make_desc_sbo(uint32_t addr, uint32_t row_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(row_bytes == 128 ? 2 : 4) << 61;
  return d;
}

// KC: channels per swizzled smem row (64 -> SWIZZLE_128B, 32 -> SWIZZLE_64B); NCH: chunks per A slot;
// RESB: weights of this CTA's row phase resident in shared memory
// A128: the activation patch is staged in 64-channel SWIZZLE_128B chunks (fewer, wider TMA rows) even when the weights use
// 32-channel chunks; channels past Cin are zero-filled by TMA and their K steps are skipped
template <int KC, int NCH, bool RESB, bool A128>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_upconv_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const ConvP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int ROWB = KC * 2;                    // weight tile row bytes
  constexpr int AKC = A128 ? 64 : KC;             // channels per activation chunk
  constexpr int AROWB = AKC * 2;
  constexpr int ANCH = (NCH * KC + AKC - 1) / AKC;  // activation chunks per slot
  const int b_sub = p.Cout * ROWB;            // one weight tile: Cout rows x one chunk
  const int b_stage = NCH * b_sub;            // streaming: one (tap, column phase) weight tile per stage
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + p.na * p.a_slot;
  const int b_bytes = RESB ? 8 * NCH * b_sub : p.nbs * b_stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + ((b_bytes + 1023) & ~1023));
  uint64_t* afull = bars;                      // [na]
  uint64_t* aempty = bars + MAX_STAGES;        // [na]
  uint64_t* bfull = bars + 2 * MAX_STAGES;     // [nbs] (bfull[0] doubles as "weights resident")
  uint64_t* bempty = bars + 3 * MAX_STAGES;    // [nbs]
  uint64_t* tfull_bar = bars + 4 * MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* bias_s = reinterpret_cast<float*>(bars + 48);  // [Cout] (<= 256 floats), 16-byte aligned
  uint8_t* stg_all = reinterpret_cast<uint8_t*>(bars) + 2048;  // 8 epilogue warps x 2 KB output staging (see the epilogue)
  for (int i = threadIdx.x; i < p.Cout; i += NTHREADS) bias_s[i] = p.bias[i];

  const int warp = uniform_warp_idx(), lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(&afull[s], 1);
      mbar_init(&aempty[s], 2);  // one tcgen05.commit per MMA-issuing warp
      mbar_init(&bfull[s], 1);
      mbar_init(&bempty[s], 2);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 2);
      mbar_init(&tempty_bar[a], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp != 0) {  // the producer lane waits after it has issued the (constant) resident weights
    pdl_wait();
    pdl_trigger();
  }

  const int groups = p.Cin / (KC * NCH);       // A slots consumed per item
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  // each CTA serves one row phase; tiles are strided over the CTAs of that phase
  const int py = blockIdx.x & 1;
  const int cta_in_phase = blockIdx.x >> 1, ctas_per_phase = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      if (RESB) {  // the 8 (px, a, b) tiles of this row phase, loaded once
        mbar_expect_tx(&bfull[0], 8 * NCH * b_sub);
        // order [a][chunk][slot], slot = (px0,b0) (px0,b1) (px1,b0) (px1,b1): the two tiles of the dx = 0 tap
        // (slots 1, 2) are adjacent, so one MMA with N = 2*Cout feeds both column phases
        for (int a = 0; a < 2; ++a)
          for (int ch = 0; ch < NCH; ++ch)
            for (int px = 0; px < 2; ++px)
              for (int b = 0; b < 2; ++b)
                tma_load_2d(smem_b + (((a * NCH + ch) * 4) + px * 2 + b) * b_sub, &mapB, &bfull[0],
                            (a * 2 + b) * p.Cin + ch * KC, (py * 2 + px) * p.Cout);
      }
      pdl_wait();
      int as_ = 0, bs_ = 0;
      uint32_t aph = 0, bph = 0;
      for (int t = cta_in_phase; t < p.num_tiles; t += ctas_per_phase) {
        const int n = t / tiles_per_img, tr = t % tiles_per_img;
        const int y0 = (tr / p.tiles_x) * TH, x0 = (tr % p.tiles_x) * TW;
        for (int cg = 0; cg < groups; ++cg) {
          mbar_wait(&aempty[as_], aph ^ 1);
          mbar_expect_tx(&afull[as_], ANCH * PH * PW * AROWB);
#pragma unroll
          for (int ch = 0; ch < ANCH; ++ch)
            tma_load_4d(smem_a + as_ * p.a_slot + ch * p.a_sub, &mapA, &afull[as_], cg * NCH * KC + ch * AKC, x0 - 1,
                        y0 - 1 + py, n);
          if (++as_ == p.na) { as_ = 0; aph ^= 1; }
          if (!RESB) {
            for (int a = 0; a < 2; ++a)
              for (int di = 0; di < 3; ++di) {
                const int dxi = di == 0 ? 1 : (di == 1 ? 0 : 2);  // centre column tap first (see the MMA warp)
                const int dx = dxi - 1, nb = dx == 0 ? 2 : 1;
                for (int s = 0; s < nb; ++s) {  // the centre column tap feeds both column phases: two stages
                  const int px = dx < 0 ? 0 : (dx > 0 ? 1 : s);
                  const int b = dx - px + 1;  // column tap inside phase px: low-res offset = b - 1 + px
                  mbar_wait(&bempty[bs_], bph ^ 1);
                  mbar_expect_tx(&bfull[bs_], NCH * b_sub);
#pragma unroll
                  for (int ch = 0; ch < NCH; ++ch)
                    tma_load_2d(smem_b + bs_ * b_stage + ch * b_sub, &mapB, &bfull[bs_],
                                (a * 2 + b) * p.Cin + (cg * NCH + ch) * KC, (py * 2 + px) * p.Cout);
                  if (++bs_ == p.nbs) { bs_ = 0; bph ^= 1; }
                }
              }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // Two MMA-issuing warps, one per column phase px (independent TMEM accumulators).  Each warp runs its loop
    // uniformly (descriptors stay in uniform registers) and one elected lane issues; both commit to the same
    // mbarriers (arrival count 2).
    {
      const int px = warp == 1 ? 0 : 1;
      const uint32_t idesc = make_idesc_bf16(BM, p.Cout);
      constexpr uint32_t A_HI = desc_hi(AROWB, PW * AROWB), B_HI = desc_hi(ROWB, 8 * ROWB);
      constexpr uint32_t A_SUB16 = (((PH * PW * AROWB) + 1023) & ~1023) >> 4;  // == p.a_sub >> 4
      const uint32_t b_sub16 = (uint32_t)b_sub >> 4;
      const uint32_t b_base_lo = desc_lo(smem_u32(smem_b));
      int as_ = 0, bs_ = 0, acc = 0;
      uint32_t aph = 0, bph = 0, tph = 0;
      if (RESB) {
        mbar_wait(&bfull[0], 0);
        tc_fence_after();
      }
      for (int t = cta_in_phase; t < p.num_tiles; t += ctas_per_phase) {
        mbar_wait(&tempty_bar[acc], tph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (acc * 2 + px) * p.acs;
        uint32_t started = 0;
        for (int cg = 0; cg < groups; ++cg) {
          mbar_wait(&afull[as_], aph);
          tc_fence_after();
          const uint32_t a_lo = desc_lo(smem_u32(smem_a + as_ * p.a_slot));
#pragma unroll
          for (int a = 0; a < 2; ++a) {
#pragma unroll
            for (int di = 0; di < 3; ++di) {
              const int dxi = di == 0 ? 1 : (di == 1 ? 0 : 2);  // same stage order as the producer
              // column tap dx = dxi-1 feeds phase px iff b = dxi - px is 0 or 1.  Streamed weights: one stage per
              // (tap, column phase) tile -- the centre tap is two stages (px 0 then px 1), the others one; both MMA warps
              // walk every stage and release it, only the warp whose phase the tile belongs to issues MMAs
              const int b = dxi - px;
              const int nst = RESB ? 1 : (dxi == 1 ? 2 : 1);
#pragma unroll
              for (int st = 0; st < nst; ++st) {
                uint32_t bstage_lo = 0;
                if (!RESB) {
                  mbar_wait(&bfull[bs_], bph);
                  tc_fence_after();
                  bstage_lo = b_base_lo + (uint32_t)bs_ * NCH * b_sub16;
                }
                const int tile_px = dxi == 0 ? 0 : (dxi == 2 ? 1 : st);  // whose tile this stage holds
                const bool mine = (b == 0 || b == 1) && (RESB || tile_px == px);
                if (elect_one()) {
                  if (mine) {
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) {
                      // shifted view of the staged patch: MMA row 8*ty+tx -> patch pixel (ty + a, tx + dxi)
                      const uint32_t view = a_lo + (((a * PW + dxi) * AROWB) >> 4);
                      const uint32_t vb = RESB ? b_base_lo + (uint32_t)(((a * NCH + ch) * 4) + px * 2 + b) * b_sub16
                                               : bstage_lo + (uint32_t)ch * b_sub16;
#pragma unroll
                      for (int k = 0; k < KC / 16; ++k) {
                        const int kk = ch * (KC / 16) + k;  // K step (16 channels) within the slot
                        const uint32_t va = view + (kk / (AKC / 16)) * A_SUB16 + 2 * (kk % (AKC / 16));
                        umma_bf16_w(d_tmem, va, A_HI, vb + 2 * k, B_HI, idesc, started);
                        started = 1;
                      }
                    }
                  }
                  if (!RESB) umma_commit(&bempty[bs_]);
                }
                __syncwarp();
                if (mine) started = 1;
                if (!RESB) {
                  if (++bs_ == p.nbs) { bs_ = 0; bph ^= 1; }
                }
              }
            }
          }
          if (elect_one()) umma_commit(&aempty[as_]);
          __syncwarp();
          if (++as_ == p.na) { as_ = 0; aph ^= 1; }
        }
        if (elect_one()) umma_commit(&tfull_bar[acc]);
        __syncwarp();
        if (++acc == p.nacc) { acc = 0; tph ^= 1; }
      }
    }
  } else {
    const int quarter = warp % 4, px = (warp - 2) / 4;  // warps 2..5 -> px 0, warps 6..9 -> px 1
    const int r = quarter * 32 + lane, ty = r / TW, tx = r % TW;
    // Output rows leave through a per-warp shared-memory transpose, 32 channels at a time: a thread's pixel is 2 * Cout * 2
    // bytes away from its neighbour's in global memory, so direct 16-byte stores cost 32 L1 wavefronts per instruction;
    // after the transpose four consecutive lanes cover one pixel's 64 bytes (8 wavefronts per instruction).
    uint8_t* stg = stg_all + (warp - 2) * 2048;
    int acc = 0;
    uint32_t tph = 0;
    for (int t = cta_in_phase; t < p.num_tiles; t += ctas_per_phase) {
      const int n = t / tiles_per_img, tr = t % tiles_per_img;
      const int y0 = (tr / p.tiles_x) * TH, x0 = (tr % p.tiles_x) * TW;
      mbar_wait(&tfull_bar[acc], tph);
      tc_fence_after();
      {
        // first output pixel of this warp's four tile rows
        bf16* warp0 = p.out + (((long long)n * (2 * p.H) + 2 * (y0 + quarter * 4) + py) * (2 * p.W) + 2 * x0 + px) * p.Cout;
        const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (acc * 2 + px) * p.acs;
        auto flush = [&](int c, int cnt) {  // cnt = 32 or 16 channels staged at [lane][cnt * 2 bytes]
          __syncwarp();
          const int per = cnt / 8;  // 16-byte pieces per pixel
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int g = 32 * i + lane;
            if (g < 32 * per) {
              const int pix = g / per, pc = g % per;
              const uint4 v = *reinterpret_cast<const uint4*>(stg + pix * 64 + ((pc ^ ((pix >> 1) & 3)) << 4));
              bf16* d = warp0 + ((long long)(2 * (pix / TW)) * (2 * p.W) + 2 * (pix % TW)) * p.Cout + c + pc * 8;
              *reinterpret_cast<uint4*>(d) = v;
            }
          }
          __syncwarp();
        };
        int c = 0;
        for (; c + 32 <= p.Cout; c += 32) {
          float v[32];
          tmem_ld32(t_addr + c, v);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c + i);
            v[i] = act_fast(v[i] + b4.x, ACT_ELU); v[i + 1] = act_fast(v[i + 1] + b4.y, ACT_ELU);
            v[i + 2] = act_fast(v[i + 2] + b4.z, ACT_ELU); v[i + 3] = act_fast(v[i + 3] + b4.w, ACT_ELU);
          }
#pragma unroll
          for (int i = 0; i < 32; i += 8)
            st8_bf16(reinterpret_cast<bf16*>(stg + lane * 64 + (((i / 8) ^ ((lane >> 1) & 3)) << 4)), v + i);
          flush(c, 32);
        }
        if (c < p.Cout) {
          float v[16];
          tmem_ld16(t_addr + c, v);
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c + i);
            v[i] = act_fast(v[i] + b4.x, ACT_ELU); v[i + 1] = act_fast(v[i + 1] + b4.y, ACT_ELU);
            v[i + 2] = act_fast(v[i + 2] + b4.z, ACT_ELU); v[i + 3] = act_fast(v[i + 3] + b4.w, ACT_ELU);
          }
#pragma unroll
          for (int i = 0; i < 16; i += 8)
            st8_bf16(reinterpret_cast<bf16*>(stg + lane * 64 + (((i / 8) ^ ((lane >> 1) & 3)) << 4)), v + i);
          flush(c, 16);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == p.nacc) { acc = 0; tph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int KC, int NCH, bool RESB, bool A128>
void launch_upconv(Ctx& c, const void* x, const void* w_tc, ConvP p) {
  const int rowb = KC * 2;
  constexpr int AKC = A128 ? 64 : KC, ANCH = (NCH * KC + AKC - 1) / AKC;
  CUtensorMap mapA, mapB;
  uint64_t da[4] = {(uint64_t)p.Cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.NB};
  uint64_t sa[3] = {(uint64_t)p.Cin * 2, (uint64_t)p.W * p.Cin * 2, (uint64_t)p.H * p.W * p.Cin * 2};
  uint32_t ba[4] = {AKC, PW, PH, 1};
  uint64_t db[2] = {(uint64_t)4 * p.Cin, (uint64_t)4 * p.Cout};
  uint64_t sb[1] = {(uint64_t)4 * p.Cin * 2};
  uint32_t bb[2] = {KC, (uint32_t)p.Cout};
  if (!encode_tmap(&mapA, x, 4, da, sa, ba, AKC * 2) || !encode_tmap(&mapB, w_tc, 2, db, sb, bb, rowb)) {
    snprintf(tls().cuda_err, sizeof(tls().cuda_err), "cuTensorMapEncodeTiled failed (tc_upconv Cin=%d Cout=%d)", p.Cin, p.Cout);
    c.fail(SJ_ECUDA);
    return;
  }
  p.a_sub = (PH * PW * AKC * 2 + 1023) & ~1023;
  p.a_slot = ANCH * p.a_sub;
  // Streamed weights: the MMA warps were waiting on the weight ring (ncu: tensor pipe 48 % at 384 -> 192 with long-scoreboard
  // stalls, L2 throughput 18 %: ring depth, not bandwidth).  Two patch slots instead of three (a patch lasts eight weight
  // stages, longer than its own TMA round trip) and ONE (tap, column phase) tile per stage instead of a two-tile stage that
  // four of six taps half filled: 2 -> 6 stages at Cout = 192, 4 -> 8 at Cout = 128; 92 -> 70 us and 90 -> 72 us per launch.
  p.na = RESB ? 3 : 2;
  const int b_sub = p.Cout * rowb;
  const int budget = 222 * 1024 - 1024 - 2048 - 16384;  // alignment slack + barriers + bias + output staging
  int b_bytes;
  if (RESB) {
    b_bytes = 8 * NCH * b_sub;
    p.nbs = 1;
  } else {
    const int b_stage = NCH * b_sub;  // one (tap, column phase) tile per stage
    p.nbs = (budget - p.na * p.a_slot) / b_stage;
    if (p.nbs > MAX_STAGES) p.nbs = MAX_STAGES;
    if (p.nbs < 2) { c.fail(SJ_EUNSUPPORTED); return; }
    b_bytes = p.nbs * b_stage;
  }
  size_t smem = 1024 + (size_t)p.na * p.a_slot + ((b_bytes + 1023) & ~1023) + 2048 + 16384;  // barriers + bias + staging
  if (smem > 227 * 1024) { c.fail(SJ_EUNSUPPORTED); return; }
  if (smem < 120 * 1024) smem = 120 * 1024;  // one CTA per SM: each CTA owns all 512 TMEM columns
  if (!SJ_SMEM_LIMIT_OK((tc_upconv_kernel<KC, NCH, RESB, A128>), 227 * 1024)) {
    c.fail(SJ_ECUDA);
    return;
  }
  int grid = num_sms() & ~1;  // even: CTA parity = output row phase
  if (grid > 2 * p.num_tiles) grid = 2 * p.num_tiles;
  SJ_LAUNCH(c, "tc_upconv", (tc_upconv_kernel<KC, NCH, RESB, A128>), grid, NTHREADS, smem, mapA, mapB, p);
}

}  // namespace

bool tc_upconv_supported(int H, int W, int Cin, int Cout) {
  if (H % TH || W % TW || Cout % 16 || Cout < 16 || Cout > 256) return false;
  if (Cin == 96) return 8 * 3 * Cout * 64 <= 100 * 1024;  // resident-weight variant
  return Cin % 64 == 0;
}

// x bf16 [NB,H,W,Cin] -> y bf16 [NB,2H,2W,Cout]; w_tc = folded kernels [4][Cout][4*Cin] bf16
void tc_upconv(Ctx& c, const void* x, void* y, const void* w_tc, const float* bias, int NB, int H, int W, int Cin,
               int Cout) {
  if (!c.ok() || c.dry) return;
  if (!tc_upconv_supported(H, W, Cin, Cout) || !w_tc || !bias) { c.fail(SJ_EUNSUPPORTED); return; }
  ConvP p{};
  p.NB = NB; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.tiles_x = W / TW; p.tiles_y = H / TH;
  p.num_tiles = NB * p.tiles_x * p.tiles_y;
  p.stack = 0;  // (one MMA warp per column phase now; the stacked centre-tap MMA is no longer used)
  p.acs = (Cout + 63) / 64 * 64;
  p.nacc = 4 * p.acs <= 512 ? 2 : 1;
  p.bias = bias;
  p.out = (bf16*)y;
  if (Cin == 96) launch_upconv<32, 3, true, true>(c, x, w_tc, p);   // 96 -> 48: weights resident (72 KB per row phase)
  else launch_upconv<64, 1, false, false>(c, x, w_tc, p);
}

}  // namespace sj
