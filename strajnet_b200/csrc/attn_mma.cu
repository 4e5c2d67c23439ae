// Attention cores of the bf16 path for the small / odd-shaped attentions of the model (bf16 operands, fp32
// accumulation and softmax):
//   * window attention of Swin layers 1-2 (64 keys, head_dim 32; modules.py:109-131),
//   * tfa.layers.MultiHeadAttention of the trajectory stack (11 / 64 keys, head sizes 64 and 42;
//     trajNet.py:20,42,80,225),
//   * FG-MSA attention with the bilinearly sampled relative-position bias (256 keys, head_dim 48;
//     FG_MSA.py:147-176).
// These are 11..256-key problems worth ~1 % of the forward's FLOPs; a 128-row tcgen05 tile with TMEM round
// trips does not fit them, so they run flash-attention style on warp-level MMAs (mma.sync.m16n8k16): one warp
// owns 16 queries, K / V of the (batch, head) are staged once per block in shared memory and read with
// ldmatrix, the logit transform (scale, bias, masks) is applied on the accumulator fragments in registers and
// the probabilities feed the P.V MMA straight from registers.
#include "kernels.h"
#include "mma_sync.cuh"

namespace sj {
namespace {

constexpr float LOG2E = 1.4426950408889634f;

enum AttnMode { AM_TFA = 0, AM_FG = 1, AM_WIN = 2 };

struct AttnP {
  const bf16* q;
  const bf16* k;
  const bf16* v;
  bf16* out;
  int ldq, ldk, ldv, ldo;
  int batch, heads, D, Nq, Nk, nk_pad;
  int hpb, qtiles;  // heads and 16-query tiles per block
  int vec16;        // K/V rows can be staged with 16-byte loads
  float scale;      // logits = scale * (q . k)
  // AM_TFA
  const int* qmask;
  const int* kmask;
  int mask_div;
  // AM_FG
  const float* fg_pos;
  const float* fg_table;
  // AM_WIN
  const float* rpb_table;
  const float* mask;  // explicit [nW,64,64] mask (mask_mode 2)
  int mask_mode, H, W, shift, nW;
};

// DP: head dim padded to a multiple of 16; KCH: keys per inner chunk (16 or 64)
template <int DP, int KCH, int MODE>
__global__ void __launch_bounds__(256) attn_mma_kernel(const AttnP p) {
  pdl_wait();  // programmatic dependent launch: see common.cuh
  pdl_trigger();
  constexpr int RS = DP + 8;  // smem row stride in elements: 16-byte aligned rows, conflict-free ldmatrix
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* Ks = reinterpret_cast<bf16*>(smem_raw);               // [hpb][nk_pad][RS]
  bf16* Vs = Ks + (size_t)p.hpb * p.nk_pad * RS;              // [hpb][nk_pad][RS]
  float* fx = reinterpret_cast<float*>(Vs + (size_t)p.hpb * p.nk_pad * RS);
  // AM_TFA: int kvalid[nk_pad]; AM_FG: float Tp[33*33], Ps[2*Nk]; AM_WIN: float tbl[225], int rid[64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int hs = warp / p.qtiles, qt = warp % p.qtiles;
  const int b = blockIdx.z, h = blockIdx.y * p.hpb + hs;
  const int nthreads = blockDim.x;

  // ---- stage K / V of every head of this block (zero-filled past Nk and past D) ----
  for (int s = 0; s < p.hpb; ++s) {
    const int hh = blockIdx.y * p.hpb + s;
    const bf16* Kg = p.k + (long long)b * p.Nk * p.ldk + hh * p.D;
    const bf16* Vg = p.v + (long long)b * p.Nk * p.ldv + hh * p.D;
    bf16* ks = Ks + (size_t)s * p.nk_pad * RS;
    bf16* vs = Vs + (size_t)s * p.nk_pad * RS;
    if (p.vec16) {
      constexpr int CH = DP / 8;
      for (int i = threadIdx.x; i < p.nk_pad * CH; i += nthreads) {
        const int m = i / CH, c = (i % CH) * 8;
        uint4 kq = make_uint4(0, 0, 0, 0), vq = kq;
        if (m < p.Nk && c < p.D) {
          kq = *reinterpret_cast<const uint4*>(Kg + (long long)m * p.ldk + c);
          vq = *reinterpret_cast<const uint4*>(Vg + (long long)m * p.ldv + c);
        }
        *reinterpret_cast<uint4*>(ks + m * RS + c) = kq;
        *reinterpret_cast<uint4*>(vs + m * RS + c) = vq;
      }
    } else {
      constexpr int CH = DP / 2;
      for (int i = threadIdx.x; i < p.nk_pad * CH; i += nthreads) {
        const int m = i / CH, c = (i % CH) * 2;
        uint32_t kq = 0, vq = 0;
        if (m < p.Nk && c < p.D) {
          kq = *reinterpret_cast<const uint32_t*>(Kg + (long long)m * p.ldk + c);
          vq = *reinterpret_cast<const uint32_t*>(Vg + (long long)m * p.ldv + c);
        }
        *reinterpret_cast<uint32_t*>(ks + m * RS + c) = kq;
        *reinterpret_cast<uint32_t*>(vs + m * RS + c) = vq;
      }
    }
  }
  const int mb = (MODE == AM_TFA) ? b / p.mask_div : 0;
  int wloc = 0;
  if (MODE == AM_TFA) {
    int* kvalid = reinterpret_cast<int*>(fx);
    for (int i = threadIdx.x; i < p.nk_pad; i += nthreads)
      kvalid[i] = (i < p.Nk) ? (p.kmask ? p.kmask[(long long)mb * p.Nk + i] : 1) : 0;
  } else if (MODE == AM_FG) {
    // zero-padded 33x33 copy of this head's 31x31 table (hpb == 1), then the key positions
    float* Tp = fx;
    float* Ps = fx + 33 * 33;
    const int hh = blockIdx.y;
    for (int i = threadIdx.x; i < 33 * 33; i += nthreads) {
      const int r = i / 33 - 1, cc = i % 33 - 1;
      Tp[i] = (r >= 0 && r < 31 && cc >= 0 && cc < 31) ? p.fg_table[(r * 31 + cc) * p.heads + hh] : 0.f;
    }
    const float* pos = p.fg_pos + ((long long)b * p.heads + hh) * p.Nk * 2;
    for (int i = threadIdx.x; i < 2 * p.Nk; i += nthreads) Ps[i] = pos[i];
  } else {
    float* tbl = fx;
    int* rid = reinterpret_cast<int*>(fx + 228);
    const int hh = blockIdx.y;  // hpb == 1
    for (int i = threadIdx.x; i < 225; i += nthreads) tbl[i] = p.rpb_table[i * p.heads + hh];
    if (p.mask_mode == 1) {
      const int nws = (p.H / 8) * (p.W / 8);
      wloc = b % nws;
      for (int n = threadIdx.x; n < 64; n += nthreads) {
        const int y = (wloc / (p.W / 8)) * 8 + n / 8, x = (wloc % (p.W / 8)) * 8 + n % 8;
        rid[n] = shift_region_id(p.H, p.W, 8, p.shift, y, x);
      }
    } else if (p.mask_mode == 2) {
      wloc = b % p.nW;
    }
  }
  __syncthreads();
  if (h >= p.heads) return;

  // ---- Q fragments straight from global memory (4-byte loads; rows / dims out of range read as 0) ----
  const int r0 = qt * 16 + blockIdx.x * p.qtiles * 16 + g, r1 = r0 + 8;
  if (qt * 16 + blockIdx.x * p.qtiles * 16 >= p.Nq) return;
  uint32_t qa[DP / 16][4];
  {
    const bf16* Q0 = p.q + ((long long)b * p.Nq + r0) * p.ldq + h * p.D;
    const bf16* Q1 = p.q + ((long long)b * p.Nq + r1) * p.ldq + h * p.D;
#pragma unroll
    for (int ks = 0; ks < DP / 16; ++ks) {
      const int c0 = ks * 16 + 2 * t, c1 = c0 + 8;
      qa[ks][0] = (r0 < p.Nq && c0 < p.D) ? *reinterpret_cast<const uint32_t*>(Q0 + c0) : 0u;
      qa[ks][1] = (r1 < p.Nq && c0 < p.D) ? *reinterpret_cast<const uint32_t*>(Q1 + c0) : 0u;
      qa[ks][2] = (r0 < p.Nq && c1 < p.D) ? *reinterpret_cast<const uint32_t*>(Q0 + c1) : 0u;
      qa[ks][3] = (r1 < p.Nq && c1 < p.D) ? *reinterpret_cast<const uint32_t*>(Q1 + c1) : 0u;
    }
  }
  int qv0 = 1, qv1 = 1;
  if (MODE == AM_TFA && p.qmask) {
    qv0 = r0 < p.Nq ? p.qmask[(long long)mb * p.Nq + r0] : 0;
    qv1 = r1 < p.Nq ? p.qmask[(long long)mb * p.Nq + r1] : 0;
  }

  const uint32_t ks_base = (uint32_t)__cvta_generic_to_shared(Ks + (size_t)hs * p.nk_pad * RS);
  const uint32_t vs_base = (uint32_t)__cvta_generic_to_shared(Vs + (size_t)hs * p.nk_pad * RS);
  // ldmatrix lane addressing (see the fragment layouts of mma.m16n8k16):
  //   K (non-transposed): tiles (keys +0, dims +0), (keys +0, dims +8), (keys +8, dims +0), (keys +8, dims +8)
  const uint32_t k_lane = (uint32_t)((((lane >> 4) * 8 + (lane & 7)) * RS + ((lane >> 3) & 1) * 8) * 2);
  //   V (transposed):     tiles (keys +0, dims +0), (keys +8, dims +0), (keys +0, dims +8), (keys +8, dims +8)
  const uint32_t v_lane = (uint32_t)(((((lane >> 3) & 1) * 8 + (lane & 7)) * RS + (lane >> 4) * 8) * 2);

  float o[DP / 8][4];
#pragma unroll
  for (int i = 0; i < DP / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int kc = 0; kc < p.nk_pad; kc += KCH) {
    float s[KCH / 8][4];
#pragma unroll
    for (int i = 0; i < KCH / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < DP / 16; ++ks) {
#pragma unroll
      for (int np = 0; np < KCH / 16; ++np) {
        uint32_t bk[4];
        ldsm_x4(bk, ks_base + (uint32_t)(((kc + np * 16) * RS + ks * 16) * 2) + k_lane);
        mma_bf16(s[2 * np], qa[ks], bk[0], bk[1]);
        mma_bf16(s[2 * np + 1], qa[ks], bk[2], bk[3]);
      }
    }
    // ---- logits (log2 domain): scale, bias, masks ----
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < KCH / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kc + nt * 8 + 2 * t + (e & 1);
        const int row = (e < 2) ? r0 : r1;
        float lg = s[nt][e] * p.scale;
        if (MODE == AM_TFA) {
          const int* kvalid = reinterpret_cast<const int*>(fx);
          const int keep = ((e < 2) ? qv0 : qv1) & kvalid[key];
          if (!keep) lg = lg + (-10e9f);  // fp32 add as in tfa: absorbs the logit (Q7)
        } else if (MODE == AM_FG) {
          // zero-border bilinear sample of the 31x31 table (occu_metric.py:394-409, tfa_image.py:116-171)
          const float* Tp = fx;
          const float* Ps = fx + 33 * 33;
          const float iq = (float)(row >> 4), jq = (float)(row & 15);
          const float r = jq - Ps[2 * key] + 1.0f, cc = iq - Ps[2 * key + 1] + 1.0f;
          const float rf = fminf(fmaxf(floorf(r), 0.f), 31.f), cf = fminf(fmaxf(floorf(cc), 0.f), 31.f);
          const float ar = fminf(fmaxf(r - rf, 0.f), 1.f), ac = fminf(fmaxf(cc - cf, 0.f), 1.f);
          const float* t0 = Tp + (int)rf * 33 + (int)cf;
          const float tl = t0[0], tr = t0[1], bl = t0[33], br = t0[34];
          const float top = ac * (tr - tl) + tl, bot = ac * (br - bl) + bl;
          lg += ar * (bot - top) + top;
        } else {
          const float* tbl = fx;
          const int* rid = reinterpret_cast<const int*>(fx + 228);
          const int n = row & 63;
          lg += tbl[((n >> 3) - (key >> 3) + 7) * 15 + ((n & 7) - (key & 7) + 7)];
          if (p.mask_mode == 1) lg += (rid[key] != rid[n]) ? -100.0f : 0.0f;
          else if (p.mask_mode == 2) lg += p.mask[((long long)wloc * 64 + n) * 64 + key];
        }
        lg *= LOG2E;
        if (key >= p.Nk) lg = -INFINITY;  // padding keys
        s[nt][e] = lg;
        if (e < 2) mx0 = fmaxf(mx0, lg);
        else mx1 = fmaxf(mx1, lg);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float n0 = fmaxf(m0, mx0), n1 = fmaxf(m1, mx1);
    const float c0 = ex2f(m0 - n0), c1 = ex2f(m1 - n1);  // first chunk: exp2(-inf) = 0
    m0 = n0; m1 = n1;
    l0 *= c0; l1 *= c1;
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) {
      o[i][0] *= c0; o[i][1] *= c0;
      o[i][2] *= c1; o[i][3] *= c1;
    }
#pragma unroll
    for (int nt = 0; nt < KCH / 8; ++nt) {
      s[nt][0] = ex2f(s[nt][0] - n0); s[nt][1] = ex2f(s[nt][1] - n0);
      s[nt][2] = ex2f(s[nt][2] - n1); s[nt][3] = ex2f(s[nt][3] - n1);
      l0 += s[nt][0] + s[nt][1];
      l1 += s[nt][2] + s[nt][3];
    }
    // ---- O += P . V (P fragments reused as the A operand) ----
#pragma unroll
    for (int j = 0; j < KCH / 16; ++j) {
      uint32_t pa[4];
      pa[0] = pack_bf16(s[2 * j][0], s[2 * j][1]);
      pa[1] = pack_bf16(s[2 * j][2], s[2 * j][3]);
      pa[2] = pack_bf16(s[2 * j + 1][0], s[2 * j + 1][1]);
      pa[3] = pack_bf16(s[2 * j + 1][2], s[2 * j + 1][3]);
#pragma unroll
      for (int dp = 0; dp < DP / 16; ++dp) {
        uint32_t bv[4];
        ldsm_x4_t(bv, vs_base + (uint32_t)(((kc + j * 16) * RS + dp * 16) * 2) + v_lane);
        mma_bf16(o[2 * dp], pa, bv[0], bv[1]);
        mma_bf16(o[2 * dp + 1], pa, bv[2], bv[3]);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  bf16* O0 = p.out + ((long long)b * p.Nq + r0) * p.ldo + h * p.D;
  bf16* O1 = p.out + ((long long)b * p.Nq + r1) * p.ldo + h * p.D;
#pragma unroll
  for (int i = 0; i < DP / 8; ++i) {
    const int c = i * 8 + 2 * t;
    if (c < p.D) {
      if (r0 < p.Nq) *reinterpret_cast<uint32_t*>(O0 + c) = pack_bf16(o[i][0] * i0, o[i][1] * i0);
      if (r1 < p.Nq) *reinterpret_cast<uint32_t*>(O1 + c) = pack_bf16(o[i][2] * i1, o[i][3] * i1);
    }
  }
  if (h == p.heads - 1) {  // columns past heads*D are defined as zero (head size 42: 126 -> 128)
    for (int c = p.heads * p.D - h * p.D + 2 * t; c + h * p.D < p.ldo; c += 8) {
      if (r0 < p.Nq) *reinterpret_cast<uint32_t*>(O0 + c) = 0u;
      if (r1 < p.Nq) *reinterpret_cast<uint32_t*>(O1 + c) = 0u;
    }
  }
}

template <int DP, int KCH, int MODE>
void launch_attn(Ctx& c, AttnP p, int grid_x, int grid_y) {
  const size_t kv = (size_t)2 * p.hpb * p.nk_pad * (DP + 8) * 2;
  size_t extra = 0;
  if (MODE == AM_TFA) extra = (size_t)p.nk_pad * 4;
  else if (MODE == AM_FG) extra = (size_t)(33 * 33 + 2 * p.Nk) * 4;
  else extra = (228 + 64) * 4;
  const size_t smem = kv + extra;
  if (smem > 48 * 1024) {
    if (!SJ_SMEM_LIMIT_OK((attn_mma_kernel<DP, KCH, MODE>), (int)smem)) {
      c.fail(SJ_ECUDA);
      return;
    }
  }
  dim3 grid(grid_x, grid_y, p.batch);
  SJ_LAUNCH(c, "attn_mma", (attn_mma_kernel<DP, KCH, MODE>), grid, 32 * p.hpb * p.qtiles, smem, p);
}

bool aligned4(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 3) == 0; }
bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; }

}  // namespace

// bf16 variant of mha_core (kernels.h); returns false when the shape is not covered (caller falls back to the
// CUDA-core kernel of attention.cu)
bool attn_mma_mha(Ctx& c, const MhaP& m) {
  if (c.dtype != SJ_BF16 || (m.D & 1) || m.D > 64 || m.Nk > 256) return false;
  if ((m.ldq | m.ldk | m.ldv | m.ldo) & 1) return false;
  if (!aligned4(m.q) || !aligned4(m.k) || !aligned4(m.v) || !aligned4(m.out)) return false;
  AttnP p{};
  p.q = (const bf16*)m.q; p.k = (const bf16*)m.k; p.v = (const bf16*)m.v; p.out = (bf16*)m.out;
  p.ldq = m.ldq; p.ldk = m.ldk; p.ldv = m.ldv; p.ldo = m.ldo;
  p.batch = m.batch; p.heads = m.heads; p.D = m.D; p.Nq = m.Nq; p.Nk = m.Nk;
  p.qmask = m.qmask; p.kmask = m.kmask; p.mask_div = m.mask_div < 1 ? 1 : m.mask_div;
  p.fg_pos = m.fg_pos; p.fg_table = m.fg_table;
  p.vec16 = (m.D % 8 == 0) && (m.ldk % 8 == 0) && (m.ldv % 8 == 0) && aligned16(m.k) && aligned16(m.v);
  const bool fg = m.fg_pos != nullptr;
  if (fg) {
    if (m.D != 48 || m.Nq != 256 || m.Nk != 256) return false;
    p.scale = 0.14433756729740643f;  // 48 ** -0.5 (FG_MSA.py:31)
    p.nk_pad = 256; p.hpb = 1; p.qtiles = 4;
    launch_attn<48, 64, AM_FG>(c, p, 4, m.heads);
    return true;
  }
  p.scale = 1.0f / sqrtf((float)m.D);  // tfa: query /= sqrt(head_size)
  const int dp = (m.D + 15) / 16 * 16;
  if (m.Nq <= 16 && m.Nk <= 16 && m.heads <= 8) {  // per-actor step attention: all heads of one actor per block
    p.nk_pad = 16; p.hpb = m.heads; p.qtiles = 1;
    if (dp == 64) launch_attn<64, 16, AM_TFA>(c, p, 1, 1);
    else if (dp == 48) launch_attn<48, 16, AM_TFA>(c, p, 1, 1);
    else return false;
    return true;
  }
  p.nk_pad = (m.Nk + 63) / 64 * 64;
  p.hpb = 1;
  const int tiles = (m.Nq + 15) / 16;
  p.qtiles = tiles >= 8 ? 8 : (tiles >= 4 ? 4 : tiles);
  const int gx = (tiles + p.qtiles - 1) / p.qtiles;
  if (dp == 64) launch_attn<64, 64, AM_TFA>(c, p, gx, m.heads);
  else if (dp == 48) launch_attn<48, 64, AM_TFA>(c, p, gx, m.heads);
  else return false;
  return true;
}

// bf16 variant of window_attn_core (kernels.h)
bool attn_mma_window(Ctx& c, const void* qkv, void* out, const float* rpb_table, int n_windows_total, int C, int heads,
                     int mask_mode, int H, int W, int shift, const float* mask, int nW) {
  const int D = C / heads;
  if (c.dtype != SJ_BF16 || (D != 32 && D != 16) || C % 8 || !aligned16(qkv) || !aligned4(out)) return false;
  AttnP p{};
  p.q = (const bf16*)qkv; p.k = p.q + C; p.v = p.q + 2 * C; p.out = (bf16*)out;
  p.ldq = p.ldk = p.ldv = 3 * C; p.ldo = C;
  p.batch = n_windows_total; p.heads = heads; p.D = D; p.Nq = 64; p.Nk = 64; p.nk_pad = 64;
  p.hpb = 1; p.qtiles = 4; p.vec16 = 1;
  p.scale = D == 32 ? 0.17677669529663687f : 0.25f;  // head_dim ** -0.5 (modules.py:73)
  p.rpb_table = rpb_table; p.mask = mask; p.mask_mode = mask_mode; p.H = H; p.W = W; p.shift = shift; p.nW = nW;
  if (D == 32) launch_attn<32, 64, AM_WIN>(c, p, 1, heads);
  else launch_attn<16, 64, AM_WIN>(c, p, 1, heads);
  return true;
}

}  // namespace sj
