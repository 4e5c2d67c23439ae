"""CPU oracle for the validation-side ops of the reference: `OGMFlow_loss.__call__`
(loss.py:50-170) and `compute_occupancy_flow_metrics` (occu_metric.py:26-140) -- forward values only.

TEST INFRASTRUCTURE ONLY (same rule as strajnet_oracle.py): only tests/, smoke() and bench.py's CPU legs import it.

PARITY UNPINNED for the third-party arithmetic: the loss and the metrics are built from
`tf.nn.sigmoid_cross_entropy_with_logits`, `tfa.losses.SigmoidFocalCrossEntropy` (alpha 0.25, gamma 2, reduction NONE),
`tf.keras.losses.BinaryCrossentropy(from_logits=False, reduction=NONE)` and `tf.keras.metrics.AUC(num_thresholds=100,
curve='PR', summation_method='interpolation')`, none of which can be executed here.  They are restated from their
published upstream definitions:
  * sigmoid CE: max(x,0) - x*z + log1p(exp(-|x|));
  * Keras backend binary_crossentropy on probabilities: p clipped to [eps, 1-eps], -(z*log(p+eps) + (1-z)*log(1-p+eps)),
    eps = 1e-7; the `BinaryCrossentropy` loss object then takes the MEAN over the last axis (loss.py flattens to
    [B, H*W] first), so `reduce_sum(self.bce(...))` is a sum over the batch of per-sample means;
  * tfa focal: alpha_t * (1-p_t)^gamma * ce, summed over the last axis (and loss.py sums the rest);
  * Keras AUC: thresholds [-1e-7, 1/99 .. 98/99, 1+1e-7] as float32, y_true cast to bool, `pred > threshold`, PR curve
    integrated with the Davis & Goadrich interpolation exactly as `AUC.interpolate_pr_auc` writes it, in float32.
The flow-warped AUC / IoU keep the reference's swapped arguments (occu_metric.py:121-126: the flow-grounded prediction
is passed as `true_occupancy`, the ground truth as `pred_occupancy`).

Tensors: pred [B,H,W,32] (channel 4k+0 observed logit, +1 occluded logit, +2,+3 flow; train.py:103-121),
gt_obs / gt_occ / origin [B,8,H,W], gt_flow [B,8,H,W,2] (train.py:126-140).
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from oracle.strajnet_oracle import bilinear_sample_zero

Tensor = torch.Tensor
EPS = 1e-7
NUM_THRESHOLDS = 100


def sigmoid_ce_with_logits(labels: Tensor, logits: Tensor) -> Tensor:
    return torch.clamp(logits, min=0) - logits * labels + torch.log1p(torch.exp(-logits.abs()))


def keras_bce_prob(labels: Tensor, prob: Tensor) -> Tensor:
    p = torch.clamp(prob, EPS, 1.0 - EPS)
    return -(labels * torch.log(p + EPS) + (1.0 - labels) * torch.log(1.0 - p + EPS))


def tfa_focal(labels: Tensor, pred: Tensor, from_logits: bool, alpha: float = 0.25, gamma: float = 2.0) -> Tensor:
    if from_logits:
        ce = sigmoid_ce_with_logits(labels, pred)
        prob = torch.sigmoid(pred)
    else:
        ce = keras_bce_prob(labels, pred)
        prob = pred
    p_t = labels * prob + (1.0 - labels) * (1.0 - prob)
    alpha_t = labels * alpha + (1.0 - labels) * (1.0 - alpha)
    return alpha_t * torch.pow(1.0 - p_t, gamma) * ce


def auc_thresholds() -> np.ndarray:
    """tf.keras.metrics.AUC.__init__: float32 thresholds (the variable dtype)."""
    t = [(i + 1) * 1.0 / (NUM_THRESHOLDS - 1) for i in range(NUM_THRESHOLDS - 2)]
    return np.asarray([0.0 - EPS] + t + [1.0 + EPS], dtype=np.float32)


def keras_pr_auc(y_true: Tensor, y_pred: Tensor, float_labels: bool = False) -> Tensor:
    """occu_metric.py:152-175 (`_compute_occupancy_auc`): one update_state, then result().

    `float_labels` selects which tf.keras the numbers are meant to reproduce (the reference pins no version):
    False: y_true is cast to bool (tf.keras <= 2.5, and >= 2.8 where the cast was restored);
    True:  tf.keras 2.6 / 2.7 evenly-spaced-threshold path (`_update_confusion_matrix_variables_optimized`), which keeps the
           label as a float: true-positive mass += y_true, false-positive mass += 1 - y_true, false negatives = total
           label mass - true positives.  Identical for 0/1 labels; differs for the fractional label of
           vehicles_flow_warped_occupancy_auc (arguments swapped at occu_metric.py:121-123)."""
    pred = y_pred.reshape(-1).float()
    thr = torch.from_numpy(auc_thresholds())
    above = (pred.unsqueeze(0) > thr.unsqueeze(1)).float()  # [T, n]
    if float_labels:
        lab = y_true.reshape(-1).float()
        tp = (above * lab).sum(1)
        fp = (above * (1.0 - lab)).sum(1)
        fn = lab.sum() - tp
    else:
        lab = (y_true.reshape(-1) != 0).float()
        tp = (above * lab).sum(1)
        fp = (above * (1.0 - lab)).sum(1)
        fn = ((1.0 - above) * lab).sum(1)
    return interpolate_pr_auc(tp, fp, fn)


def _div_no_nan(a: Tensor, b: Tensor) -> Tensor:
    return torch.where(b == 0, torch.zeros_like(a), a / torch.where(b == 0, torch.ones_like(b), b))


def interpolate_pr_auc(tp: Tensor, fp: Tensor, fn: Tensor) -> Tensor:
    """tf.keras.metrics.AUC.interpolate_pr_auc, float32, op for op (n = number of thresholds)."""
    n = tp.numel()
    dtp = tp[: n - 1] - tp[1:]
    p = tp + fp
    dp = p[: n - 1] - p[1:]
    prec_slope = _div_no_nan(dtp, torch.clamp(dp, min=0))
    intercept = tp[1:] - prec_slope * p[1:]
    safe_p_ratio = torch.where((p[: n - 1] > 0) & (p[1:] > 0), _div_no_nan(p[: n - 1], torch.clamp(p[1:], min=0)),
                               torch.ones_like(p[1:]))
    inc = _div_no_nan(prec_slope * (dtp + intercept * torch.log(safe_p_ratio)), torch.clamp(tp[1:] + fn[1:], min=0))
    return inc.sum()


def _identity_warp(H: int, W: int) -> Tensor:
    """loss.py:75-86 / occu_metric.py:279-292: [H,W,2] storing (x, y)."""
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    return torch.stack((xs, ys), dim=-1)


def _split(pred: Tensor, k: int):
    c = pred[..., 4 * k: 4 * k + 4]
    return c[..., 0], c[..., 1], c[..., 2:4]


def ogm_flow_loss(pred: Tensor, gt_obs: Tensor, gt_occ: Tensor, gt_flow: Tensor, origin: Tensor, ogm_weight=1000.0,
                  occ_weight=1000.0, flow_weight=1.0, replica=1.0, flow_origin_weight=1000.0, no_use_warp=False,
                  use_pred=False, use_focal_loss=True, use_gt=False) -> Dict[str, Tensor]:
    """OGMFlow_loss.__call__ (loss.py:50-170) on logits `pred`; returns the four mean losses (+ the per-waypoint
    gates `res` of the use_gt branch)."""
    B, H, W, _ = pred.shape
    ident = _identity_warp(H, W)
    size = float(B * H * W)
    obs_l, occ_l, flow_l, warp_l, f_c = [], [], [], [], []
    for k in range(8):
        lo, lc, pf = _split(pred, k)
        to, tc, tf_, org = gt_obs[:, k], gt_occ[:, k], gt_flow[:, k], origin[:, k].unsqueeze(-1)

        def occ_xe(t, logit, weight):  # _sigmoid_xe_loss / _sigmoid_occ_loss (loss.py:172-222)
            s = sigmoid_ce_with_logits(t, logit).sum()
            if use_focal_loss:
                s = tfa_focal(t, logit, True).sum() + s
            return weight * s / (size * replica)

        obs_l.append(occ_xe(to, lo, ogm_weight))
        occ_l.append(occ_xe(tc, lc, occ_weight))
        true_all = torch.clamp(to + tc, 0, 1)
        if use_gt:  # loss.py:125-135
            wp_org = bilinear_sample_zero(org, ident + tf_)[..., 0]
            auc = keras_pr_auc(true_all, wp_org * true_all)
            res = ((1 - auc) < 1.0).float()
        else:
            res = torch.tensor(1.0)
        f_c.append(res)
        # _flow_loss (loss.py:272-292)
        exists = ((tf_[..., 0] != 0) | (tf_[..., 1] != 0)).float()
        diff = (tf_ - pf) * exists.unsqueeze(-1)
        flow_l.append(res * _div_no_nan(diff.abs().sum(-1).sum(), exists.sum() * replica / 2))
        if not no_use_warp:  # loss.py:140-155
            wp = bilinear_sample_zero(org, ident + pf)[..., 0]
            if use_pred:
                a = torch.sigmoid(lo) + torch.sigmoid(lc)
            else:  # the reference passes the TRUE occupancies as "pred_occupancy_*" here (loss.py:152-154)
                a = torch.sigmoid(to) + torch.sigmoid(tc)
            joint = torch.clamp(a, 0, 1) * wp
            bce_sum = keras_bce_prob(true_all, joint).reshape(B, -1).mean(-1).sum()
            if use_pred:  # _sigmoid_xe_warp_loss_pred: the last assignment wins (loss.py:266)
                xe = bce_sum
            elif use_focal_loss:
                xe = tfa_focal(true_all, joint, False).sum() + bce_sum
            else:
                xe = sigmoid_ce_with_logits(true_all, joint).sum()
            warp_l.append(res * flow_origin_weight * xe / (size * replica))
    fc = torch.stack(f_c).sum()
    out = {
        "observed_xe": torch.stack(obs_l).sum() / 8,
        "occluded_xe": torch.stack(occ_l).sum() / 8,
        "flow": torch.stack(flow_l).sum() / fc,
        "flow_warp_xe": torch.stack(warp_l).sum() / fc if not no_use_warp else torch.tensor(0.0),
        "res": torch.stack(f_c),
    }
    return out


def occupancy_flow_metrics(pred: Tensor, gt_obs: Tensor, gt_occ: Tensor, gt_flow: Tensor, origin: Tensor,
                           pred_is_logits: bool = True, no_warp: bool = False, auc_float_labels: bool = False) -> Dict[str, Tensor]:
    """compute_occupancy_flow_metrics (occu_metric.py:26-140); `pred_is_logits` applies
    `_apply_sigmoid_to_occupancy_logits` (train.py:142-154) first."""
    B, H, W, _ = pred.shape
    ident = _identity_warp(H, W)
    names = ["vehicles_observed_auc", "vehicles_occluded_auc", "vehicles_observed_iou", "vehicles_occluded_iou",
             "vehicles_flow_epe", "vehicles_flow_warped_occupancy_auc", "vehicles_flow_warped_occupancy_iou"]
    acc = {n: [] for n in names}

    def soft_iou(t, p):  # occu_metric.py:178-201
        t, p = t.reshape(-1), p.reshape(-1)
        inter = (p * t).mean()
        return _div_no_nan(inter, p.mean() + t.mean() - inter)

    for k in range(8):
        lo, lc, pf = _split(pred, k)
        po, pc = (torch.sigmoid(lo), torch.sigmoid(lc)) if pred_is_logits else (lo, lc)
        to, tc, tf_, org = gt_obs[:, k], gt_occ[:, k], gt_flow[:, k], origin[:, k].unsqueeze(-1)
        acc["vehicles_observed_auc"].append(keras_pr_auc(to, po, auc_float_labels))
        acc["vehicles_observed_iou"].append(soft_iou(to, po))
        acc["vehicles_occluded_auc"].append(keras_pr_auc(tc, pc, auc_float_labels))
        acc["vehicles_occluded_iou"].append(soft_iou(tc, pc))
        # _compute_flow_epe (occu_metric.py:204-250)
        exists = ((tf_[..., 0] != 0) | (tf_[..., 1] != 0)).float()
        diff = (tf_ - pf) * exists.unsqueeze(-1)
        epe = torch.sqrt((diff * diff).sum(-1))
        acc["vehicles_flow_epe"].append(_div_no_nan(epe.sum(), exists.sum()))
        if not no_warp:
            true_all = torch.clamp(to + tc, 0, 1)
            pred_all = torch.clamp(po + pc, 0, 1)
            wp = bilinear_sample_zero(org, ident + pf)[..., 0]
            g = pred_all * wp
            acc["vehicles_flow_warped_occupancy_auc"].append(keras_pr_auc(g, true_all, auc_float_labels))  # swapped, as in the reference
            acc["vehicles_flow_warped_occupancy_iou"].append(soft_iou(g, true_all))
    return {n: (torch.stack(v).sum() / len(v) if v else torch.tensor(0.0)) for n, v in acc.items()}


def make_eval_inputs(B: int, H: int = 256, seed: int = 0) -> Dict[str, Tensor]:
    """Synthetic validation batch with the value ranges of train.py:84-101: boolean occupancy grids (blobs, ~3 % of the
    cells), sparse backward flow where vehicles are, flow-origin occupancy = observed|occluded of the previous waypoint,
    and logits that are informative but imperfect (so every AUC bin is populated)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ys, xs = np.mgrid[0:H, 0:H].astype(np.float32)
    obs = np.zeros((B, 8, H, H), np.float32)
    occ = np.zeros((B, 8, H, H), np.float32)
    flow = np.zeros((B, 8, H, H, 2), np.float32)
    org = np.zeros((B, 8, H, H), np.float32)
    for b in range(B):
        n = int(rng.integers(0, 14)) if b != 1 else 0  # sample 1: an empty scene (divide_no_nan paths)
        cx, cy = rng.uniform(10, H - 10, n), rng.uniform(10, H - 10, n)
        vx, vy = rng.normal(0, 3, n), rng.normal(0, 3, n)
        occl = rng.random(n) < 0.3
        prev = np.zeros((H, H), np.float32)
        for k in range(8):
            cur_o = np.zeros((H, H), np.float32)
            cur_c = np.zeros((H, H), np.float32)
            for i in range(n):
                m = (np.abs(xs - (cx[i] + vx[i] * (k + 1))) < 4) & (np.abs(ys - (cy[i] + vy[i] * (k + 1))) < 2.5)
                (cur_c if occl[i] else cur_o)[m] = 1.0
                flow[b, k][m] = (-vx[i], -vy[i])
            obs[b, k], occ[b, k] = cur_o, cur_c
            org[b, k] = prev if k else np.clip(cur_o + cur_c, 0, 1)
            prev = np.clip(cur_o + cur_c, 0, 1)
    pred = np.zeros((B, H, H, 32), np.float32)
    for k in range(8):
        pred[..., 4 * k + 0] = 4.0 * (np.clip(obs[:, k], 0, 1) - 0.55) + rng.normal(0, 2.0, (B, H, H))
        pred[..., 4 * k + 1] = 4.0 * (np.clip(occ[:, k], 0, 1) - 0.6) + rng.normal(0, 2.0, (B, H, H))
        pred[..., 4 * k + 2: 4 * k + 4] = flow[:, k] + rng.normal(0, 0.7, (B, H, H, 2))
    d = dict(pred=pred, gt_obs=obs, gt_occ=occ, gt_flow=flow, origin=org)
    return {k: torch.from_numpy(v.astype(np.float32)) for k, v in d.items()}
