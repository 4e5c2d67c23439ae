// Validation-side ops of the reference, forward values only, as ONE pass over the grids:
//   * OGMFlow_loss.__call__ (loss.py:50-170): sigmoid cross-entropy (+ tfa focal) on the observed / occluded occupancy
//     logits, masked L1 flow loss, flow-warp consistency loss (zero-border bilinear `sample` of the flow-origin occupancy
//     at identity + flow, occu_metric.py:345-409), the use_gt gate (PR-AUC of the ground-truth warp);
//   * compute_occupancy_flow_metrics (occu_metric.py:26-140): PR-AUC (tf.keras.metrics.AUC, 100 thresholds,
//     interpolation), soft IoU, end-point error, flow-warped occupancy AUC / IoU (arguments swapped as in the reference).
// HBM-bound: every input element (logits [B,H,W,32], three [B,8,H,W] grids, one [B,8,H,W,2] flow field) is read once;
// the four bilinear taps of the warp hit L1/L2.  Per waypoint the kernel keeps 18 running sums (fp32 per thread, fp64
// across the block, written per block and summed in a fixed order by the finalize kernel: deterministic) and four
// 2 x 101-bin histograms of "number of AUC thresholds below the prediction" (integer counts: order-independent).
#include "kernels.h"

namespace sj {
namespace {

constexpr int NACC = 18, NAUC = 4, NBIN = 101, HIST = NAUC * 2 * NBIN, NT = 256;
enum Acc { A_OBS_SCE, A_OBS_FOCAL, A_OCC_SCE, A_OCC_FOCAL, A_FLOW_L1, A_EXISTS, A_WARP_MAIN, A_WARP_BCE, A_IOU_OBS,
           A_SUM_TO, A_SUM_PO, A_IOU_OCC, A_SUM_TC, A_SUM_PC, A_EPE, A_IOU_G, A_SUM_G, A_SUM_TA };

struct EvalP {
  const float* pred;     // [B,H,W,32]
  const float* gt_obs;   // [B,8,H,W]
  const float* gt_occ;   // [B,8,H,W]
  const float* gt_flow;  // [B,8,H,W,2]
  const float* origin;   // [B,8,H,W]
  int B, H, W, flags, nblk;
  double* partial;  // [8][nblk][NACC]
  int* hist;        // [8][NAUC][2][NBIN]
  float* fhist;     // [8][NAUC][2][NBIN] label MASS per bin (SJ_EVAL_AUC_FLOAT_LABELS), else unused
};

__device__ __forceinline__ float sce(float z, float x) {  // tf.nn.sigmoid_cross_entropy_with_logits
  return fmaxf(x, 0.f) - x * z + __logf(1.0f + __expf(-fabsf(x)));  // MUFU-grade: the sums are compared at 2e-4
}
__device__ __forceinline__ float bce_prob(float z, float p) {  // Keras backend binary_crossentropy, from_logits=False
  const float eps = 1e-7f;
  p = fminf(fmaxf(p, eps), 1.0f - eps);
  return -(z * __logf(p + eps) + (1.0f - z) * __logf(1.0f - p + eps));
}
__device__ __forceinline__ float focal(float z, float prob, float ce) {  // tfa SigmoidFocalCrossEntropy, alpha .25, gamma 2
  const float p_t = z * prob + (1.0f - z) * (1.0f - prob);
  const float a_t = z * 0.25f + (1.0f - z) * 0.75f;
  const float m = 1.0f - p_t;
  return a_t * (m * m) * ce;
}
__device__ __forceinline__ float sigmoidf(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// sample(image, warp, pixel_type=0) with a zero border: pad by 1, warp + 1, floor clamped to [0, size-2], alpha
// clamped to [0,1] (occu_metric.py:394-409, tfa_image.py:116-171); (wx, wy) = (x, y) source coordinates
__device__ __forceinline__ float sample_zero(const float* __restrict__ img, int H, int W, float wx, float wy) {
  const float qx = wx + 1.0f, qy = wy + 1.0f;
  const float fx = fminf(fmaxf(floorf(qx), 0.f), (float)W), fy = fminf(fmaxf(floorf(qy), 0.f), (float)H);
  const float ax = fminf(fmaxf(qx - fx, 0.f), 1.f), ay = fminf(fmaxf(qy - fy, 0.f), 1.f);
  const int ix = (int)fx - 1, iy = (int)fy - 1;  // back to unpadded coordinates: taps (iy, ix) .. (iy+1, ix+1)
  auto at = [&](int y, int x) { return (y >= 0 && y < H && x >= 0 && x < W) ? img[y * W + x] : 0.f; };
  const float tl = at(iy, ix), tr = at(iy, ix + 1), bl = at(iy + 1, ix), br = at(iy + 1, ix + 1);
  const float top = ax * (tr - tl) + tl, bot = ax * (br - bl) + bl;
  return ay * (bot - top) + top;
}

// number of AUC thresholds strictly below v (thr[0] = -1e-7, thr[i] = i/99, thr[99] = 1 + 1e-7, all float32)
// Branch-free: c0 = floor(99 v) is within one of the answer, so three table compares around it settle it exactly
// (thresholds 0 .. c0-2 are below v for sure, thresholds c0+2 .. are not).  NaN counts no threshold, like `pred > thr`.
__device__ __forceinline__ int auc_bin(float v, const float* thr) {
  const int c0 = max(1, min(98, (int)(v * 99.0f)));
  return (c0 - 1) + (thr[c0 - 1] < v) + (thr[c0] < v) + (thr[c0 + 1] < v);
}
// all 32 lanes call; lanes with key < 0 contribute nothing.  Occupancy grids are mostly empty, so a warp's 32 cells
// usually fall into ONE bin: that case costs one vote and one atomic; otherwise every lane adds its own count.
__device__ __forceinline__ void hist_add(int* h, int key, int lane) {
  const int first = __shfl_sync(0xffffffffu, key, 0);
  if (__all_sync(0xffffffffu, key == first)) {
    if (lane == 0 && key >= 0) atomicAdd(&h[key], 32);
  } else if (key >= 0) {
    atomicAdd(&h[key], 1);
  }
}

__global__ void __launch_bounds__(NT) eval_pass_kernel(const EvalP p) {
  pdl_wait();
  pdl_trigger();
  __shared__ int hist_s[HIST];
  float* fh_s = reinterpret_cast<float*>(hist_s);  // same storage: one of the two kinds of histogram is in use
  __shared__ float thr_s[100];
  __shared__ float red_s[NT / 32][NACC];
  const int tid = threadIdx.x, lane = tid % 32, k = blockIdx.y;
  for (int i = tid; i < HIST; i += NT) hist_s[i] = 0;
  if (tid < 100) thr_s[tid] = tid == 0 ? 0.0f - 1e-7f : (tid == 99 ? 1.0f + 1e-7f : (float)((double)tid / 99.0));
  __syncthreads();
  const bool use_focal = p.flags & SJ_EVAL_USE_FOCAL, no_use_warp = p.flags & SJ_EVAL_NO_USE_WARP,
             use_pred = p.flags & SJ_EVAL_USE_PRED, use_gt = p.flags & SJ_EVAL_USE_GT,
             is_prob = p.flags & SJ_EVAL_PRED_IS_PROB, do_loss = p.flags & SJ_EVAL_LOSS,
             do_metrics = p.flags & SJ_EVAL_METRICS, no_warp_m = p.flags & SJ_EVAL_METRICS_NO_WARP,
             float_labels = p.flags & SJ_EVAL_AUC_FLOAT_LABELS;
  const int HW = p.H * p.W;
  float acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
  // work item = one image row (b, y) x 256 columns: the index arithmetic is warp-uniform and done once per item
  const int xchunks = (p.W + NT - 1) / NT, items = p.B * p.H * xchunks;
  for (int item = blockIdx.x; item < items; item += p.nblk) {
    const int row = item / xchunks, b = row / p.H, y = row % p.H, x = (item % xchunks) * NT + tid;
    const bool valid = x < p.W;
    int key[NAUC] = {-1, -1, -1, -1};
    float lab[NAUC] = {0.f, 0.f, 0.f, 0.f};  // label value of each AUC sample (float-label semantics only)
    if (valid) {
      const int pix = y * p.W + x;
      const long long idx = (long long)b * HW + pix;
      const float4 pr = *reinterpret_cast<const float4*>(p.pred + idx * 32 + 4 * k);
      const long long g = ((long long)b * 8 + k) * HW + pix;
      const float to = p.gt_obs[g], tc = p.gt_occ[g];
      const float2 tf = *reinterpret_cast<const float2*>(p.gt_flow + 2 * g);
      const float* org = p.origin + ((long long)b * 8 + k) * HW;
      const float ta = fminf(fmaxf(to + tc, 0.f), 1.f);
      const float exists = (tf.x != 0.f || tf.y != 0.f) ? 1.f : 0.f;
      const float dx = (tf.x - pr.z) * exists, dy = (tf.y - pr.w) * exists;
      const bool need_wp = (do_loss && !no_use_warp) || (do_metrics && !no_warp_m);
      const float wp = need_wp ? sample_zero(org, p.H, p.W, (float)x + pr.z, (float)y + pr.w) : 0.f;
      float po = pr.x, pc = pr.y;  // probabilities (metrics); pr.x / pr.y stay the logits for the loss
      if (!is_prob) { po = sigmoidf(pr.x); pc = sigmoidf(pr.y); }
      if (do_loss) {
        const float so = sce(to, pr.x), sc = sce(tc, pr.y);
        acc[A_OBS_SCE] += so;
        acc[A_OCC_SCE] += sc;
        if (use_focal) {
          acc[A_OBS_FOCAL] += focal(to, po, so);  // do_loss implies logits: po / pc are their sigmoids
          acc[A_OCC_FOCAL] += focal(tc, pc, sc);
        }
        acc[A_FLOW_L1] += fabsf(dx) + fabsf(dy);
        if (use_gt) {
          const float wo = sample_zero(org, p.H, p.W, (float)x + tf.x, (float)y + tf.y);
          key[3] = (ta != 0.f ? NBIN : 0) + auc_bin(wo * ta, thr_s);
          lab[3] = ta;
        }
        if (!no_use_warp) {
          const float a = use_pred ? po + pc : sigmoidf(to) + sigmoidf(tc);
          const float joint = fminf(fmaxf(a, 0.f), 1.f) * wp;
          const float bce = bce_prob(ta, joint);
          acc[A_WARP_BCE] += bce;
          if (!use_pred) acc[A_WARP_MAIN] += use_focal ? focal(ta, joint, bce) : sce(ta, joint);
        }
      }
      acc[A_EXISTS] += exists;
      if (do_metrics) {
        key[0] = (to != 0.f ? NBIN : 0) + auc_bin(po, thr_s);
        key[1] = (tc != 0.f ? NBIN : 0) + auc_bin(pc, thr_s);
        lab[0] = to;
        lab[1] = tc;
        acc[A_IOU_OBS] += po * to; acc[A_SUM_TO] += to; acc[A_SUM_PO] += po;
        acc[A_IOU_OCC] += pc * tc; acc[A_SUM_TC] += tc; acc[A_SUM_PC] += pc;
        acc[A_EPE] += sqrtf(dx * dx + dy * dy);
        if (!no_warp_m) {
          const float gq = fminf(fmaxf(po + pc, 0.f), 1.f) * wp;
          key[2] = (gq != 0.f ? NBIN : 0) + auc_bin(ta, thr_s);  // swapped arguments, occu_metric.py:121-123
          lab[2] = gq;  // the fractional flow-grounded prediction is the LABEL here
          acc[A_IOU_G] += ta * gq; acc[A_SUM_G] += gq; acc[A_SUM_TA] += ta;
        }
      }
    }
    if (float_labels) {
      // tf.keras >= 2.6 evenly-spaced-threshold path (_update_confusion_matrix_variables_optimized): the label is not cast
      // to bool; a sample adds y_true to the true-positive mass of its bin and 1 - y_true to the false-positive mass
#pragma unroll
      for (int a = 0; a < NAUC; ++a)
        if (key[a] >= 0) {
          const int bin = key[a] >= NBIN ? key[a] - NBIN : key[a];
          if (lab[a] != 0.f) atomicAdd(&fh_s[a * 2 * NBIN + NBIN + bin], lab[a]);
          if (lab[a] != 1.f) atomicAdd(&fh_s[a * 2 * NBIN + bin], 1.0f - lab[a]);
        }
    } else {
#pragma unroll
      for (int a = 0; a < NAUC; ++a) hist_add(hist_s + a * 2 * NBIN, key[a], lane);
    }
  }
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    const float s = warp_sum(acc[i]);
    if (lane == 0) red_s[tid / 32][i] = s;
  }
  __syncthreads();
  if (tid < NACC) {
    double s = 0.0;
    for (int w = 0; w < NT / 32; ++w) s += (double)red_s[w][tid];
    p.partial[((long long)k * p.nblk + blockIdx.x) * NACC + tid] = s;
  }
  if (float_labels) {
    for (int i = tid; i < HIST; i += NT)
      if (fh_s[i] != 0.f) atomicAdd(&p.fhist[k * HIST + i], fh_s[i]);
  } else {
    for (int i = tid; i < HIST; i += NT)
      if (hist_s[i]) atomicAdd(&p.hist[k * HIST + i], hist_s[i]);
  }
}

struct FinP {
  const double* partial;
  const int* hist;
  const float* fhist;
  int B, H, W, flags, nblk;
  float ogm_weight, occ_weight, flow_origin_weight, replica;
  float* out;  // [SJ_EVAL_OUT_FLOATS]
};

__device__ __forceinline__ float div_no_nan(float a, float b) { return b == 0.f ? 0.f : a / b; }

// tf.keras.metrics.AUC.interpolate_pr_auc in float32 from the histogram of one (waypoint, AUC)
__device__ float pr_auc(const int* h) {
  const int* hf = h;         // label false
  const int* ht = h + NBIN;  // label true
  int total_t = 0;
  for (int j = 0; j < NBIN; ++j) total_t += ht[j];
  // tp[i] = #(label & pred > thr[i]) = sum_{j > i} ht[j]
  int tp_i = 0, fp_i = 0;
  for (int j = 1; j < NBIN; ++j) { tp_i += ht[j]; fp_i += hf[j]; }  // i = 0
  float auc = 0.f;
  float tp0 = (float)tp_i, p0 = (float)(tp_i + fp_i);
  for (int i = 1; i < 100; ++i) {
    tp_i -= ht[i]; fp_i -= hf[i];
    const float tp1 = (float)tp_i, p1 = (float)tp_i + (float)fp_i, fn1 = (float)(total_t - tp_i);
    const float dtp = tp0 - tp1, dp = p0 - p1;
    const float slope = div_no_nan(dtp, fmaxf(dp, 0.f));
    const float intercept = tp1 - slope * p1;
    const float ratio = (p0 > 0.f && p1 > 0.f) ? div_no_nan(p0, fmaxf(p1, 0.f)) : 1.0f;
    auc += div_no_nan(slope * (dtp + intercept * logf(ratio)), fmaxf(tp1 + fn1, 0.f));
    tp0 = tp1; p0 = p1;
  }
  return auc;
}

// the same integral from label-mass histograms (float-label semantics): tp / fp are sums of masses
__device__ float pr_auc_f(const float* h) {
  const float* hf = h;
  const float* ht = h + NBIN;
  float total_t = 0.f;
  for (int j = 0; j < NBIN; ++j) total_t += ht[j];
  float tp_i = 0.f, fp_i = 0.f;
  for (int j = 1; j < NBIN; ++j) { tp_i += ht[j]; fp_i += hf[j]; }
  float auc = 0.f;
  float tp0 = tp_i, p0 = tp_i + fp_i;
  for (int i = 1; i < 100; ++i) {
    tp_i -= ht[i]; fp_i -= hf[i];
    const float tp1 = tp_i, p1 = tp_i + fp_i, fn1 = total_t - tp_i;
    const float dtp = tp0 - tp1, dp = p0 - p1;
    const float slope = div_no_nan(dtp, fmaxf(dp, 0.f));
    const float intercept = tp1 - slope * p1;
    const float ratio = (p0 > 0.f && p1 > 0.f) ? div_no_nan(p0, fmaxf(p1, 0.f)) : 1.0f;
    auc += div_no_nan(slope * (dtp + intercept * logf(ratio)), fmaxf(tp1 + fn1, 0.f));
    tp0 = tp1; p0 = p1;
  }
  return auc;
}

__global__ void __launch_bounds__(256) eval_finalize_kernel(const FinP p) {
  pdl_wait();
  pdl_trigger();
  __shared__ double sums[8][NACC];
  __shared__ float aucs[8][NAUC];
  const int tid = threadIdx.x;
  if (tid < 8 * NACC) {
    const int k = tid / NACC, a = tid % NACC;
    double s = 0.0;
    for (int b = 0; b < p.nblk; ++b) s += p.partial[((long long)k * p.nblk + b) * NACC + a];
    sums[k][a] = s;
  }
  if (tid >= 192 && tid < 192 + 8 * NAUC) {
    const int t = tid - 192, k = t / NAUC, a = t % NAUC;
    aucs[k][a] = (p.flags & SJ_EVAL_AUC_FLOAT_LABELS) ? pr_auc_f(p.fhist + (k * NAUC + a) * 2 * NBIN)
                                                       : pr_auc(p.hist + (k * NAUC + a) * 2 * NBIN);
  }
  __syncthreads();
  if (tid != 0) return;
  const bool use_focal = p.flags & SJ_EVAL_USE_FOCAL, no_use_warp = p.flags & SJ_EVAL_NO_USE_WARP,
             use_pred = p.flags & SJ_EVAL_USE_PRED, use_gt = p.flags & SJ_EVAL_USE_GT,
             no_warp_m = p.flags & SJ_EVAL_METRICS_NO_WARP;
  const float size = (float)p.B * (float)p.H * (float)p.W, hw = (float)p.H * (float)p.W;
  float obs = 0.f, occ = 0.f, flow = 0.f, warp = 0.f, fc = 0.f;
  float m[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < 8; ++k) {
    const double* s = sums[k];
    const float res = use_gt ? ((1.0f - aucs[k][3]) < 1.0f ? 1.f : 0.f) : 1.f;
    p.out[11 + k] = res;
    fc += res;
    obs += p.ogm_weight * (float)(s[A_OBS_SCE] + (use_focal ? s[A_OBS_FOCAL] : 0.0)) / (size * p.replica);
    occ += p.occ_weight * (float)(s[A_OCC_SCE] + (use_focal ? s[A_OCC_FOCAL] : 0.0)) / (size * p.replica);
    flow += res * div_no_nan((float)s[A_FLOW_L1], (float)s[A_EXISTS] * p.replica / 2.0f);
    if (!no_use_warp) {
      const float bce_sum = (float)(s[A_WARP_BCE] / (double)hw);  // sum over the batch of per-sample means
      const float xe = use_pred ? bce_sum : (use_focal ? (float)s[A_WARP_MAIN] + bce_sum : (float)s[A_WARP_MAIN]);
      warp += res * p.flow_origin_weight * xe / (size * p.replica);
    }
    m[0] += aucs[k][0];
    m[1] += aucs[k][1];
    const float n = size;
    float inter = (float)(s[A_IOU_OBS] / n);
    m[2] += div_no_nan(inter, (float)(s[A_SUM_PO] / n) + (float)(s[A_SUM_TO] / n) - inter);
    inter = (float)(s[A_IOU_OCC] / n);
    m[3] += div_no_nan(inter, (float)(s[A_SUM_PC] / n) + (float)(s[A_SUM_TC] / n) - inter);
    m[4] += div_no_nan((float)s[A_EPE], (float)s[A_EXISTS]);
    if (!no_warp_m) {
      m[5] += aucs[k][2];
      inter = (float)(s[A_IOU_G] / n);
      m[6] += div_no_nan(inter, (float)(s[A_SUM_TA] / n) + (float)(s[A_SUM_G] / n) - inter);
    }
  }
  p.out[0] = obs / 8.0f;
  p.out[1] = occ / 8.0f;
  p.out[2] = flow / fc;  // 0/0 = NaN when every waypoint is gated off, as tf.math.add_n(...) / add_n(f_c) gives
  p.out[3] = no_use_warp ? 0.f : warp / fc;
  for (int i = 0; i < 7; ++i) p.out[4 + i] = m[i] / 8.0f;
}

int eval_blocks() { return 4 * num_sms() / 8 > 0 ? 4 * num_sms() / 8 : 1; }  // per waypoint: 4 blocks per SM in all

}  // namespace

size_t eval_workspace_bytes() {
  return (size_t)8 * eval_blocks() * NACC * sizeof(double) + (size_t)2 * 8 * HIST * sizeof(int) + 1024;
}

void eval_forward(Ctx& c, const float* pred, const float* gt_obs, const float* gt_occ, const float* gt_flow,
                  const float* origin, int B, int H, int W, int flags, float ogm_weight, float occ_weight,
                  float flow_origin_weight, float replica, float* out) {
  const int nblk = eval_blocks();
  double* partial = (double*)c.alloc((size_t)8 * nblk * NACC * sizeof(double));
  int* hist = (int*)c.alloc((size_t)2 * 8 * HIST * sizeof(int));  // integer counts, then float label masses
  if (!c.ok() || c.dry) return;
  if (!partial || !hist) return;  // workspace overflow is reported by run()
  if (cudaMemsetAsync(hist, 0, (size_t)2 * 8 * HIST * sizeof(int), c.stream) != cudaSuccess) { c.fail(SJ_ECUDA); return; }
  float* fhist = reinterpret_cast<float*>(hist + 8 * HIST);
  EvalP p{pred, gt_obs, gt_occ, gt_flow, origin, B, H, W, flags, nblk, partial, hist, fhist};
  SJ_LAUNCH(c, "eval_pass", eval_pass_kernel, dim3(nblk, 8), NT, 0, p);
  FinP f{partial, hist, fhist, B, H, W, flags, nblk, ogm_weight, occ_weight, flow_origin_weight, replica, out};
  SJ_LAUNCH(c, "eval_finalize", eval_finalize_kernel, 1, 256, 0, f);
}

}  // namespace sj
