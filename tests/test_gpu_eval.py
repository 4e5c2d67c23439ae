"""GPU parity of the validation-side ops (strajnet_b200/evaluation.py -> sj_ogm_flow_eval_fwd) against
oracle/eval_oracle.py.  Tolerance: 2e-4 relative (+1e-5 abs) on every scalar -- both sides accumulate ~1e6 fp32 terms,
in different orders (the kernel in fp64 across blocks, torch pairwise in fp32); the AUC bin counts are integers, equal up to
the few cells whose MUFU-grade sigmoid lands on the other side of a threshold."""
import pytest
import torch

from oracle import eval_oracle as E

pytestmark = pytest.mark.gpu

LOSS_FLAGS = [dict(), dict(use_focal_loss=False), dict(use_gt=True, use_focal_loss=False),  # train.py:195-196
              dict(use_pred=True), dict(no_use_warp=True), dict(use_gt=True, replica=2.0, ogm_weight=500.0,
                                                                occ_weight=250.0, flow_origin_weight=2000.0)]


def _close(a, b, what):
    a, b = float(a), float(b)
    if a != a or b != b:
        assert a != a and b != b, f"{what}: {a} vs {b}"
        return
    assert abs(a - b) <= 1e-5 + 2e-4 * abs(b), f"{what}: {a} vs {b}"


def _cuda(d):
    return {k: v.cuda() for k, v in d.items()}


@pytest.mark.parametrize("flags", LOSS_FLAGS)
def test_ogm_flow_loss_parity(flags):
    from strajnet_b200.evaluation import OGMFlow_loss
    d = E.make_eval_inputs(2, 256, seed=1)
    ref = E.ogm_flow_loss(**d, **flags)
    got = OGMFlow_loss(None, **flags).packed(**_cuda(d))
    for k in ("observed_xe", "occluded_xe", "flow", "flow_warp_xe"):
        _close(got[k], ref[k], f"{k} {flags}")


@pytest.mark.parametrize("B,H,no_warp", [(2, 256, False), (3, 64, False), (2, 256, True), (16, 256, False)])
def test_metrics_and_fused_eval_parity(B, H, no_warp):
    from strajnet_b200 import evaluation as V
    d = E.make_eval_inputs(B, H, seed=2)
    ref_m = E.occupancy_flow_metrics(**d, no_warp=no_warp)
    ref_l = E.ogm_flow_loss(**d, use_gt=True, use_focal_loss=False)
    loss, m = V.evaluate(**{("pred_logits" if k == "pred" else k): v for k, v in _cuda(d).items()},
                         loss=V.OGMFlow_loss(None, use_gt=True, use_focal_loss=False), no_warp=no_warp)
    for k in V.METRIC_NAMES:
        if no_warp and "warped" in k:
            assert getattr(m, k) == 0.0
        else:
            _close(getattr(m, k), ref_m[k], k)
    for k in V.LOSS_NAMES:
        _close(loss[k], ref_l[k], k)
    assert torch.equal(loss["res"].cpu(), ref_l["res"])


@pytest.mark.parametrize("B,H", [(2, 256), (16, 256)])
def test_metrics_auc_float_label_semantics(B, H):
    """SJ_EVAL_AUC_FLOAT_LABELS (tf.keras 2.6 / 2.7): labels keep their float value.  The warped-occupancy AUC -- whose
    label is the fractional flow-grounded prediction -- changes, every binary-label metric stays put."""
    from strajnet_b200 import evaluation as V
    d = E.make_eval_inputs(B, H, seed=4)
    ref_bool = E.occupancy_flow_metrics(**d)
    ref_float = E.occupancy_flow_metrics(**d, auc_float_labels=True)
    _, m = V.evaluate(**{("pred_logits" if k == "pred" else k): v for k, v in _cuda(d).items()}, auc_float_labels=True)
    for k in V.METRIC_NAMES:
        _close(getattr(m, k), ref_float[k], k + " (float labels)")
    assert abs(float(ref_float["vehicles_observed_auc"]) - float(ref_bool["vehicles_observed_auc"])) < 1e-6
    assert abs(float(ref_float["vehicles_flow_warped_occupancy_auc"]) - float(ref_bool["vehicles_flow_warped_occupancy_auc"])) > 1e-3


def test_reference_call_surface_and_edge_cases():
    """The reference's own calling convention (train.py:103-154, 252-283): WaypointGrids of per-waypoint tensors, logits
    for the loss, probabilities for the metrics; empty scene -> divide_no_nan zeros; every gate off -> NaN flow loss."""
    from strajnet_b200 import evaluation as V
    from strajnet_b200.loss import OGMFlow_loss
    from strajnet_b200.occu_metric import compute_occupancy_flow_metrics
    d = E.make_eval_inputs(2, 128, seed=3)
    g = _cuda(d)
    true, logits, probs = V.WaypointGrids(), V.WaypointGrids(), V.WaypointGrids()
    for k in range(8):
        true.vehicles.observed_occupancy.append(g["gt_obs"][:, k, :, :, None])
        true.vehicles.occluded_occupancy.append(g["gt_occ"][:, k, :, :, None])
        true.vehicles.flow.append(g["gt_flow"][:, k])
        true.vehicles.flow_origin_occupancy.append(g["origin"][:, k, :, :, None])
        c = g["pred"][..., 4 * k: 4 * k + 4]
        logits.vehicles.observed_occupancy.append(c[..., :1])
        logits.vehicles.occluded_occupancy.append(c[..., 1:2])
        logits.vehicles.flow.append(c[..., 2:])
        probs.vehicles.observed_occupancy.append(torch.sigmoid(c[..., :1]))
        probs.vehicles.occluded_occupancy.append(torch.sigmoid(c[..., 1:2]))
        probs.vehicles.flow.append(c[..., 2:])
    cfg = type("Cfg", (), dict(grid_height_cells=128, grid_width_cells=128, num_waypoints=8))()
    loss = OGMFlow_loss(cfg, replica=1.0, no_use_warp=False, use_pred=False, use_gt=True, use_focal_loss=False)
    got = loss(true_waypoints=true, pred_waypoint_logits=logits, curr_ogm=None)
    ref = E.ogm_flow_loss(**d, use_gt=True, use_focal_loss=False)
    for k in V.LOSS_NAMES:
        _close(got[k], ref[k], k)
    m = compute_occupancy_flow_metrics(config=cfg, true_waypoints=true, pred_waypoints=probs, no_warp=False)
    ref_m = E.occupancy_flow_metrics(**d)
    for k in V.METRIC_NAMES:
        _close(getattr(m, k), ref_m[k], k)
    with pytest.raises(ValueError):
        OGMFlow_loss(type("Cfg", (), dict(grid_height_cells=256, grid_width_cells=256, num_waypoints=8))()).packed(**g)
    empty = {k: torch.zeros_like(v) for k, v in g.items()}
    l0, m0 = V.evaluate(**{("pred_logits" if k == "pred" else k): v for k, v in empty.items()},
                        loss=V.OGMFlow_loss(None, use_gt=True))
    assert torch.isnan(l0["flow"]) and all(getattr(m0, k) == getattr(m0, k) for k in V.METRIC_NAMES)
    assert m0.vehicles_flow_epe == 0.0 and m0.vehicles_observed_auc == 0.0
    with pytest.raises(RuntimeError):
        V.OGMFlow_loss().packed(**d)  # CPU tensors: no fallback
