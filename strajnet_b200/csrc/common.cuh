// Shared device/host helpers for the strajnet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/strajnet_b200.h"

namespace sj {

typedef __nv_bfloat16 bf16;

// ---- per-thread library state ---------------------------------------------------------------
struct TlsState {
  long long launches = 0;
  long long tc_launches = 0;  // launches of tcgen05 kernels (names starting with "tc_")
  char cuda_err[256] = {0};
  // opt-in timing probe (sj_probe_start / sj_probe_stop): CUDA events around every launch whose
  // role starts with `probe_role`, recorded on the launching stream
  static constexpr int kMaxProbe = 1024;
  bool probe_on = false;
  char probe_role[64] = {0};
  int probe_n = 0;
  cudaEvent_t probe_ev[2 * kMaxProbe] = {};
  bool probe_ev_ready = false;
  // fork/join helper stream of the whole-model forward (actor branch of the trajectory stack runs beside the raster
  // encoder); created lazily per host thread and device, never destroyed
  cudaStream_t side_stream = nullptr;
  cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
  int side_device = -1;
};
TlsState& tls();

// ---- execution context: stream, activation dtype, workspace arena ---------------------------
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0, high = 0;
  bool overflow = false;
  void* alloc(size_t bytes) {
    size_t a = (off + 255) & ~size_t(255);
    off = a + bytes;
    if (off > high) high = off;
    if (base == nullptr || off > cap) {
      overflow = true;
      return nullptr;
    }
    return base + a;
  }
  size_t mark() const { return off; }
  void release(size_t m) { off = m; }
};

struct Ctx {
  cudaStream_t stream = nullptr;
  int dtype = SJ_F32;
  bool dry = false;  // dry run: only size the workspace, launch nothing
  const char* role = "";  // which step of the forward is being enqueued (for the timing probe)
  Arena ws;
  int status = SJ_OK;
  size_t esize() const { return dtype == SJ_BF16 ? 2 : 4; }
  void* alloc(size_t bytes) { return ws.alloc(bytes); }
  // element-count allocation in the activation dtype
  void* alloc_act(size_t n) { return ws.alloc(n * esize()); }
  bool ok() const { return status == SJ_OK; }
  void fail(int s) {
    if (status == SJ_OK) status = s;
  }
};

// records a launch, and the CUDA error if one is pending
void note_launch(Ctx& c, const char* what);
// timing probe hooks around a launch (no-ops unless a probe is armed and the role matches)
int probe_before(Ctx& c);
void probe_after(Ctx& c, int slot);

// Programmatic dependent launch (PDL): every kernel of the library executes pdl_wait() before its first access to
// memory another kernel of the forward may have written (and before its first global write), and pdl_trigger() near
// its top.  Launched with programmatic stream serialisation, kernel i+1 then gets its CTAs scheduled, barriers
// initialised, TMEM allocated and constant weights staged while kernel i drains, instead of after its last CTA has
// retired.  Both are no-ops for a kernel launched without the attribute, which is the default (opt in: sj_set_pdl /
// SJ_PDL_MASK).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// which launches get the attribute: bit 0 = tcgen05 kernels ("tc_*"), bit 1 = all others, bit 2 = front half of the forward only
int pdl_mask();

template <typename... Exp, typename... Act>
inline void launch_kernel(const char* what, const char* role, cudaStream_t stream, dim3 grid, dim3 block, size_t smem,
                          void (*kernel)(Exp...), Act&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  const bool is_tc = what[0] == 't' && what[1] == 'c' && what[2] == '_';
  // bit 2 of the mask restricts programmatic launches to the latency-bound front half of the forward (encoder, FG-MSA,
  // trajectory stack: ~75 launches of a few microseconds each); the decoder's long persistent kernels lose from early-
  // resident successors holding shared memory / TMEM slots
  const int mask = pdl_mask();
  bool front = true;
  if (mask & 4) front = role && (role[0] == 'e' || role[0] == 'f' || role[0] == 't');  // "enc", "fgmsa", "traj*"
  cfg.numAttrs = ((mask & (is_tc ? 1 : 2)) && front) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<Act&&>(args)...);
}

// Raises a kernel's dynamic shared-memory limit on first use per (call site, host thread, device) instead of on every
// launch; evaluates to true when the limit is in place.
#define SJ_SMEM_LIMIT_OK(kernel, bytes)                                                                        \
  ([&]() -> bool {                                                                                             \
    static thread_local int sj_dev_done_ = -1;                                                                 \
    static thread_local cudaError_t sj_err_ = cudaSuccess;                                                     \
    int sj_dev_ = 0;                                                                                           \
    if (cudaGetDevice(&sj_dev_) != cudaSuccess) return false;                                                  \
    if (sj_dev_ != sj_dev_done_) {                                                                             \
      sj_err_ = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes));        \
      sj_dev_done_ = sj_dev_;                                                                                  \
    }                                                                                                          \
    return sj_err_ == cudaSuccess;                                                                             \
  }())

#define SJ_LAUNCH(ctx, what, kernel, grid, block, smem, ...)                                   \
  do {                                                                                        \
    if (!(ctx).dry && (ctx).ok()) {                                                           \
      int sj_slot_ = sj::probe_before(ctx);                                                   \
      sj::launch_kernel(what, (ctx).role, (ctx).stream, dim3(grid), dim3(block), (smem), kernel, __VA_ARGS__);  \
      sj::probe_after((ctx), sj_slot_);                                                       \
      sj::note_launch((ctx), what);                                                           \
    }                                                                                         \
  } while (0)

// ---- typed element access --------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// 4 consecutive elements (pointer must be 4-element aligned)
template <typename T> __device__ __forceinline__ float4 ld4(const T* p);
template <> __device__ __forceinline__ float4 ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 ld4<bf16>(const bf16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T> __device__ __forceinline__ void st4(T* p, float4 v);
template <> __device__ __forceinline__ void st4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void st4<bf16>(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// ---- math --------------------------------------------------------------------------------------
// tanh-GELU exactly as modules.py:18-29
__device__ __forceinline__ float gelu_tanh(float x) {
  const float k = 0.7978845608028654f;  // sqrt(2/pi)
  float u = k * (x + 0.044715f * x * x * x);
  return x * (0.5f * (1.0f + tanhf(u)));
}
__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }
enum Act { ACT_NONE = 0, ACT_GELU = 1, ACT_ELU = 2 };
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_GELU) return gelu_tanh(v);
  if (act == ACT_ELU) return elu1(v);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Region id of the shifted-window mask image (modules.py:192-203) at position (y, x) of the
// *shifted* image: rows/cols split at H-ws and H-shift, id = 3*rowband + colband.
__host__ __device__ __forceinline__ int shift_region_id(int H, int W, int ws, int shift, int y, int x) {
  int by = y < H - ws ? 0 : (y < H - shift ? 1 : 2);
  int bx = x < W - ws ? 0 : (x < W - shift ? 1 : 2);
  return by * 3 + bx;
}

// Submission quantisation of one waypoint (inference.py:124-136, :160-182): occupancy logits -> sigmoid ->
// round(p*255) as uint8; flow -> clip(round(f), -128, 127) as int8 (np.round = round-half-even = rint).
__device__ __forceinline__ uint32_t quantize_waypoint(float obs, float occ, float fx, float fy) {
  const float po = 1.0f / (1.0f + expf(-obs)), pc = 1.0f / (1.0f + expf(-occ));
  const uint32_t qo = (uint32_t)__float2int_rn(po * 255.0f), qc = (uint32_t)__float2int_rn(pc * 255.0f);
  const int ix = max(-128, min(127, __float2int_rn(fx))), iy = max(-128, min(127, __float2int_rn(fy)));
  return qo | (qc << 8) | ((uint32_t)(ix & 0xff) << 16) | ((uint32_t)(iy & 0xff) << 24);
}
// raw model inputs as the reference's record decode produces them (inference.py:91-93)
enum InType { IN_F32 = 0, IN_U8 = 1, IN_I8_DIV256 = 2 };
__device__ __forceinline__ float load_input(const void* base, long long idx, int type) {
  if (type == IN_U8) return reinterpret_cast<const uint8_t*>(base)[idx] != 0 ? 1.0f : 0.0f;  // bool raster -> float
  if (type == IN_I8_DIV256) return (float)reinterpret_cast<const int8_t*>(base)[idx] / 256.0f;  // int8 / 256
  return reinterpret_cast<const float*>(base)[idx];
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace sj
