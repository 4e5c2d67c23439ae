"""CPU tests of the validation-side oracle (oracle/eval_oracle.py): each restated third-party op is
pinned against an independent implementation or a published known answer."""
import math

import numpy as np
import torch
import torch.nn.functional as F

from oracle import eval_oracle as E
from oracle import strajnet_oracle as O
from tests.util import randn


def test_pr_auc_interpolation_known_answer():
    """The worked example in Keras' own metrics tests (AUC(num_thresholds=3, curve='PR', interpolation), weighted case):
    tp = [7,4,0], fp = [3,0,0], fn = [0,3,7]  ->  (0.5*(3 + 2*log(2.5)) + 4) / 7."""
    tp, fp, fn = (torch.tensor(v, dtype=torch.float32) for v in ([7, 4, 0], [3, 0, 0], [0, 3, 7]))
    got = E.interpolate_pr_auc(tp, fp, fn).item()
    assert abs(got - (0.5 * (3 + 2 * math.log(2.5)) + 4) / 7) < 1e-6
    assert abs(got - 0.916613) < 1e-5


def test_pr_auc_thresholds_and_counts():
    thr = E.auc_thresholds()
    assert thr.dtype == np.float32 and thr.shape == (100,)
    assert thr[0] < 0 < thr[1] and thr[98] < 1 < thr[99] and np.all(np.diff(thr) > 0)
    assert thr[1] == np.float32(1.0 / 99.0) and thr[98] == np.float32(98.0 / 99.0)
    # brute force over thresholds, float64 reference of the same integral
    rng = np.random.Generator(np.random.PCG64(3))
    y = (rng.random(5000) < 0.2).astype(np.float32)
    p = np.clip(0.3 * y + rng.normal(0.3, 0.2, 5000), 0, 1).astype(np.float32)
    tp = np.array([np.sum((p > t) & (y > 0)) for t in thr], np.float64)
    fp = np.array([np.sum((p > t) & (y == 0)) for t in thr], np.float64)
    fn = y.sum() - tp
    ref = 0.0
    for i in range(99):
        dtp, pa, pb = tp[i] - tp[i + 1], tp[i] + fp[i], tp[i + 1] + fp[i + 1]
        dp = pa - pb
        slope = dtp / dp if dp > 0 else 0.0
        icpt = tp[i + 1] - slope * pb
        ratio = pa / pb if (pa > 0 and pb > 0) else 1.0
        den = tp[i + 1] + fn[i + 1]
        ref += slope * (dtp + icpt * math.log(ratio)) / den if den > 0 else 0.0
    got = E.keras_pr_auc(torch.from_numpy(y), torch.from_numpy(p)).item()
    assert abs(got - ref) < 2e-5
    # a perfect ranking reaches 1, an empty ground truth gives 0 (divide_no_nan)
    assert abs(E.keras_pr_auc(torch.from_numpy(y), torch.from_numpy(y)).item() - 1.0) < 1e-6
    assert E.keras_pr_auc(torch.zeros(100), torch.rand(100)).item() == 0.0


def test_sigmoid_ce_and_focal_against_torch():
    x, z = randn((4, 1000), 1, 3.0), (randn((4, 1000), 2) > 0.8).float()
    assert torch.allclose(E.sigmoid_ce_with_logits(z, x), F.binary_cross_entropy_with_logits(x, z, reduction="none"), atol=1e-6)
    # focal loss, alpha 0.25 / gamma 2: torchvision's independent implementation of the same published formula
    from torchvision.ops import sigmoid_focal_loss
    assert torch.allclose(E.tfa_focal(z, x, True), sigmoid_focal_loss(x, z, alpha=0.25, gamma=2.0, reduction="none"), atol=1e-6)
    # Keras' probability form adds eps = 1e-7 inside the logs: equal to torch's BCE away from p = 0 / 1
    p = torch.sigmoid(0.5 * x)
    assert torch.allclose(E.keras_bce_prob(z, p), F.binary_cross_entropy(p, z, reduction="none"), atol=1e-5, rtol=1e-4)
    assert torch.allclose(E.tfa_focal(z, p, False), E.tfa_focal(z, 0.5 * x, True), atol=1e-5, rtol=1e-4)


def test_zero_border_sampler_against_grid_sample():
    """sample(..., pixel_type=0) with BorderType.ZERO == bilinear grid_sample(padding_mode='zeros', align_corners=True),
    including coordinates far outside the image."""
    img = randn((2, 20, 28, 1), 4)
    warp = torch.stack((randn((2, 20, 28), 5, 12.0) + 14, randn((2, 20, 28), 6, 9.0) + 10), -1)
    got = O.bilinear_sample_zero(img, warp)[..., 0]
    grid = torch.stack((warp[..., 0] / 27 * 2 - 1, warp[..., 1] / 19 * 2 - 1), -1)
    ref = F.grid_sample(img.permute(0, 3, 1, 2), grid, mode="bilinear", padding_mode="zeros", align_corners=True)[:, 0]
    assert torch.allclose(got, ref, atol=1e-5)


def test_loss_and_metrics_structure():
    d = E.make_eval_inputs(2, 64, seed=0)
    base = E.ogm_flow_loss(**d)
    assert set(base) == {"observed_xe", "occluded_xe", "flow", "flow_warp_xe", "res"}
    assert all(torch.isfinite(base[k]).all() for k in base)
    # no focal term: strictly smaller occupancy losses; no_use_warp: the warp loss is the constant 0.0 (loss.py:168)
    nf = E.ogm_flow_loss(**d, use_focal_loss=False)
    assert nf["observed_xe"] < base["observed_xe"] and nf["flow"] == base["flow"]
    assert E.ogm_flow_loss(**d, no_use_warp=True)["flow_warp_xe"].item() == 0.0
    # replica divides every mean (loss.py:196,292)
    r2 = E.ogm_flow_loss(**d, replica=2.0)
    assert abs(r2["observed_xe"].item() * 2 - base["observed_xe"].item()) < 1e-3 * base["observed_xe"].item()
    # perfect predictions: flow loss 0, EPE 0, IoU -> 1, AUC -> 1
    perfect = dict(d)
    pred = torch.zeros_like(d["pred"])
    for k in range(8):
        pred[..., 4 * k] = 40 * d["gt_obs"][:, k] - 20
        pred[..., 4 * k + 1] = 40 * d["gt_occ"][:, k] - 20
        pred[..., 4 * k + 2: 4 * k + 4] = d["gt_flow"][:, k]
    perfect["pred"] = pred
    assert E.ogm_flow_loss(**perfect)["flow"].item() == 0.0
    m = E.occupancy_flow_metrics(**perfect)
    assert m["vehicles_flow_epe"].item() == 0.0
    assert m["vehicles_observed_iou"].item() > 0.999 and m["vehicles_observed_auc"].item() > 0.999
    # an empty scene: divide_no_nan everywhere, no NaN
    empty = {k: torch.zeros_like(v) for k, v in d.items()}
    assert all(torch.isfinite(v) for v in E.occupancy_flow_metrics(**empty).values())
    # use_gt with every waypoint gated off: 0 / 0, as tf.math.add_n(...) / add_n(f_c) gives
    assert torch.isnan(E.ogm_flow_loss(**empty, use_gt=True)["flow"])


def test_eval_oracle_matches_golden(golden_dir):
    """Drift guard: tests/golden/eval_outputs.npz (tests/golden/make_golden.py --eval-only)."""
    g = np.load(f"{golden_dir}/eval_outputs.npz")
    d = E.make_eval_inputs(2, 64, seed=0)
    flags = {"default": dict(), "train_py": dict(use_gt=True, use_focal_loss=False), "use_pred": dict(use_pred=True)}
    for name, f in flags.items():
        r = E.ogm_flow_loss(**d, **f)
        got = np.array([r[k].item() for k in ("observed_xe", "occluded_xe", "flow", "flow_warp_xe")])
        np.testing.assert_allclose(got, g[f"loss_{name}"], rtol=2e-5)
    m = E.occupancy_flow_metrics(**d)
    np.testing.assert_allclose(np.array([v.item() for v in m.values()]), g["metrics"], rtol=2e-5)


def test_pr_auc_float_label_semantics():
    """tf.keras 2.6 / 2.7 keep y_true as a float in the evenly-spaced-threshold update: identical to the bool cast for 0/1
    labels, a mass-weighted confusion matrix otherwise (hand-worked: two samples, one threshold region)."""
    rng = np.random.Generator(np.random.PCG64(5))
    y = (rng.random(500) < 0.3).astype(np.float32)
    p = np.clip(0.6 * y + rng.normal(0.2, 0.25, 500), 0, 1).astype(np.float32)
    a = E.keras_pr_auc(torch.from_numpy(y), torch.from_numpy(p)).item()
    b = E.keras_pr_auc(torch.from_numpy(y), torch.from_numpy(p), float_labels=True).item()
    assert abs(a - b) < 1e-6
    # fractional labels: label 0.25 at prediction 0.9, label 0.0 at prediction 0.1
    yt, yp = torch.tensor([0.25, 0.0]), torch.tensor([0.9, 0.1])
    thr = torch.from_numpy(E.auc_thresholds())
    above = (yp[None] > thr[:, None]).float()
    tp, fp = (above * yt).sum(1), (above * (1 - yt)).sum(1)
    want = E.interpolate_pr_auc(tp, fp, yt.sum() - tp).item()
    got = E.keras_pr_auc(yt, yp, float_labels=True).item()
    assert abs(got - want) < 1e-7
    assert abs(got - E.keras_pr_auc(yt, yp).item()) > 0.05  # the bool cast counts the 0.25 label as a full positive
