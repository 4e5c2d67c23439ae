"""Validation-side ops of the reference, forward values only, through `sj_ogm_flow_eval_fwd`:

* `OGMFlow_loss` -- same constructor arguments and call signature as `loss.py:22-60` (`loss_fn(pred_waypoint_logits=...,
  true_waypoints=..., curr_ogm=...)` -> dict with `observed_xe`, `occluded_xe`, `flow`, `flow_warp_xe`);
* `compute_occupancy_flow_metrics(config, true_waypoints, pred_waypoints, no_warp=False)` -- `occu_metric.py:26-140`,
  returning an object with the seven `vehicles_*` attributes of the `OccupancyFlowMetrics` proto;
* `WaypointGrids` -- the container the reference borrows from `waymo_open_dataset.utils.occupancy_flow_grids`
  (`.vehicles.observed_occupancy / occluded_occupancy / flow / flow_origin_occupancy`: lists of per-waypoint tensors), so
  `train.py:103-140` (`_get_pred_waypoint_logits`, `_warpped_gt`) keep working unchanged;
* `evaluate(...)` -- both results from ONE pass over packed tensors (what `val_step`, train.py:252-283, needs).

No backward pass (inference / validation only).  There is no CPU fallback: the kernels run on the GPU or raise.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib as L

Tensor = torch.Tensor
LOSS_NAMES = ("observed_xe", "occluded_xe", "flow", "flow_warp_xe")
METRIC_NAMES = ("vehicles_observed_auc", "vehicles_occluded_auc", "vehicles_observed_iou", "vehicles_occluded_iou",
                "vehicles_flow_epe", "vehicles_flow_warped_occupancy_auc", "vehicles_flow_warped_occupancy_iou")
FLAG = dict(use_focal=1, no_use_warp=2, use_pred=4, use_gt=8, pred_is_prob=16, loss=32, metrics=64, metrics_no_warp=128,
            auc_float_labels=256)
OUT_FLOATS = 19


class _AgentGrids:
    """occupancy_flow_grids._WaypointGridsOneType: per-waypoint lists."""

    def __init__(self):
        self.observed_occupancy: List[Tensor] = []
        self.occluded_occupancy: List[Tensor] = []
        self.flow: List[Tensor] = []
        self.flow_origin_occupancy: List[Tensor] = []


class WaypointGrids:
    """occupancy_flow_grids.WaypointGrids (only `vehicles` is used by the reference, train.py:103-140)."""

    def __init__(self):
        self.vehicles = _AgentGrids()
        self.pedestrians = _AgentGrids()
        self.cyclists = _AgentGrids()


def _dev(t, device) -> Tensor:
    return torch.as_tensor(t).to(device=device, dtype=torch.float32)


def pack_predictions(grids: WaypointGrids, device="cuda") -> Tensor:
    """Per-waypoint lists ([B,H,W,1], [B,H,W,1], [B,H,W,2]) -> [B,H,W,32], channel 4k + {0, 1, 2, 3} (the inverse of
    `_get_pred_waypoint_logits`, train.py:103-121)."""
    v = grids.vehicles
    if not (len(v.observed_occupancy) == len(v.occluded_occupancy) == len(v.flow) == 8):
        raise ValueError("expected 8 waypoints of observed_occupancy / occluded_occupancy / flow")
    parts = []
    for k in range(8):
        parts += [_dev(v.observed_occupancy[k], device), _dev(v.occluded_occupancy[k], device), _dev(v.flow[k], device)]
    return torch.cat(parts, dim=-1).contiguous()


def pack_truth(grids: WaypointGrids, device="cuda"):
    """Per-waypoint lists -> gt_obs, gt_occ, origin [B,8,H,W] and gt_flow [B,8,H,W,2] (the inverse of `_warpped_gt`)."""
    v = grids.vehicles
    if not (len(v.observed_occupancy) == len(v.occluded_occupancy) == len(v.flow) == len(v.flow_origin_occupancy) == 8):
        raise ValueError("expected 8 waypoints of observed_occupancy / occluded_occupancy / flow / flow_origin_occupancy")

    def stack(xs, c):
        xs = [_dev(x, device) for x in xs]
        out = torch.stack(xs, dim=1)  # [B,8,H,W,c]
        if out.shape[-1] != c:
            raise ValueError(f"expected {c} channel(s), got {out.shape[-1]}")
        return (out[..., 0] if c == 1 else out).contiguous()

    return (stack(v.observed_occupancy, 1), stack(v.occluded_occupancy, 1), stack(v.flow, 2),
            stack(v.flow_origin_occupancy, 1))


def _run(pred: Tensor, gt_obs: Tensor, gt_occ: Tensor, gt_flow: Tensor, origin: Tensor, flags: int, ogm_weight=1000.0,
         occ_weight=1000.0, flow_origin_weight=1000.0, replica=1.0) -> Tensor:
    if not pred.is_cuda:
        raise RuntimeError("strajnet_b200.evaluation runs on the GPU only (no CPU fallback)")
    dev = pred.device
    pred = pred.to(torch.float32).contiguous()
    if pred.dim() != 4 or pred.shape[-1] != 32:
        raise ValueError(f"pred must be [B,H,W,32], got {tuple(pred.shape)}")
    B, H, W, _ = pred.shape
    gt_obs, gt_occ, origin = (_dev(t, dev).contiguous() for t in (gt_obs, gt_occ, origin))
    gt_flow = _dev(gt_flow, dev).contiguous()
    for name, t, shp in (("gt_obs", gt_obs, (B, 8, H, W)), ("gt_occ", gt_occ, (B, 8, H, W)), ("origin", origin, (B, 8, H, W)),
                         ("gt_flow", gt_flow, (B, 8, H, W, 2))):
        if tuple(t.shape) != shp:
            raise ValueError(f"{name} must be {shp}, got {tuple(t.shape)}")
    lib = L.lib()
    with torch.cuda.device(dev):
        ws = torch.empty(lib.sj_ogm_flow_eval_workspace_bytes() + 256, dtype=torch.uint8, device=dev)
        wp = (ws.data_ptr() + 255) & ~255
        out = torch.empty(OUT_FLOATS, dtype=torch.float32, device=dev)
        prm = L.SjEvalParams(flags, ogm_weight, occ_weight, flow_origin_weight, replica)
        st = lib.sj_ogm_flow_eval_fwd(pred.data_ptr(), gt_obs.data_ptr(), gt_occ.data_ptr(), gt_flow.data_ptr(),
                                      origin.data_ptr(), B, H, W, C.byref(prm), out.data_ptr(), wp,
                                      ws.numel() - (wp - ws.data_ptr()), torch.cuda.current_stream(dev).cuda_stream)
        L.check(st, "ogm_flow_eval")
        ws.record_stream(torch.cuda.current_stream(dev))
    return out


class OGMFlow_loss:
    """loss.py:20-170, forward value only.  `flow_weight` is accepted and unused, as in the reference (its
    `_flow_loss` is always called with the default weight 1, loss.py:137)."""

    def __init__(self, config=None, ogm_weight=1000.0, occ_weight=1000.0, flow_weight=1.0, replica=1.0,
                 flow_origin_weight=1000.0, no_use_warp=False, use_pred=False, use_focal_loss=True, use_gt=False):
        self.config = config
        self.ogm_weight, self.occ_weight, self.flow_weight = ogm_weight, occ_weight, flow_weight
        self.replica, self.flow_origin_weight = replica, flow_origin_weight
        self.no_use_warp, self.use_pred, self.use_focal_loss, self.use_gt = no_use_warp, use_pred, use_focal_loss, use_gt

    def flags(self) -> int:
        return (FLAG["use_focal"] * bool(self.use_focal_loss) | FLAG["no_use_warp"] * bool(self.no_use_warp) |
                FLAG["use_pred"] * bool(self.use_pred) | FLAG["use_gt"] * bool(self.use_gt))

    def _check_config(self, H, W):
        c = self.config
        if c is not None and (getattr(c, "grid_height_cells", H) != H or getattr(c, "grid_width_cells", W) != W or
                              getattr(c, "num_waypoints", 8) != 8):
            raise ValueError("config does not match the grids (height/width cells, 8 waypoints)")

    def packed(self, pred: Tensor, gt_obs: Tensor, gt_occ: Tensor, gt_flow: Tensor, origin: Tensor) -> Dict[str, Tensor]:
        """pred [B,H,W,32] logits (the model output as is), gt_* as train.py:84-101 decodes them ([B,8,H,W,(2)])."""
        self._check_config(pred.shape[1], pred.shape[2])
        out = _run(pred, gt_obs, gt_occ, gt_flow, origin, self.flags() | FLAG["loss"], self.ogm_weight, self.occ_weight,
                   self.flow_origin_weight, self.replica)
        d = {n: out[i] for i, n in enumerate(LOSS_NAMES)}
        if self.no_use_warp:
            d["flow_warp_xe"] = 0.0  # loss.py:168
        return d

    def __call__(self, pred_waypoint_logits: WaypointGrids, true_waypoints: WaypointGrids, curr_ogm=None) -> Dict[str, Tensor]:
        pred = pack_predictions(pred_waypoint_logits)
        return self.packed(pred, *pack_truth(true_waypoints, pred.device))


def compute_occupancy_flow_metrics(config, true_waypoints: WaypointGrids, pred_waypoints: WaypointGrids,
                                   no_warp: bool = False, auc_float_labels: bool = False):
    """occu_metric.py:26-140.  `pred_waypoints` holds occupancy PROBABILITIES (train.py:142-154).
    `auc_float_labels`: tf.keras 2.6 / 2.7 AUC semantics (labels kept as floats) instead of the bool cast of every
    other tf.keras version; only vehicles_flow_warped_occupancy_auc can differ (SJ_EVAL_AUC_FLOAT_LABELS)."""
    pred = pack_predictions(pred_waypoints)
    out = _run(pred, *pack_truth(true_waypoints, pred.device),
               FLAG["metrics"] | FLAG["pred_is_prob"] | (FLAG["metrics_no_warp"] if no_warp else 0)
               | (FLAG["auc_float_labels"] if auc_float_labels else 0))
    return _metrics_object(out, no_warp)


def _metrics_object(out: Tensor, no_warp: bool):
    vals = out[4:11].tolist()  # host read, as the reference's `_mean(...).numpy()` (occu_metric.py:143-149)
    m = SimpleNamespace(**dict(zip(METRIC_NAMES, vals)))
    if no_warp:  # unset proto fields read as 0
        m.vehicles_flow_warped_occupancy_auc = 0.0
        m.vehicles_flow_warped_occupancy_iou = 0.0
    return m


def evaluate(pred_logits: Tensor, gt_obs: Tensor, gt_occ: Tensor, gt_flow: Tensor, origin: Tensor,
             loss: Optional[OGMFlow_loss] = None, no_warp: bool = False, auc_float_labels: bool = False):
    """val_step (train.py:252-283) in one pass: the model's logits [B,H,W,32] -> (loss dict, metrics object)."""
    loss = loss or OGMFlow_loss()
    out = _run(pred_logits, gt_obs, gt_occ, gt_flow, origin,
               loss.flags() | FLAG["loss"] | FLAG["metrics"] | (FLAG["metrics_no_warp"] if no_warp else 0)
               | (FLAG["auc_float_labels"] if auc_float_labels else 0), loss.ogm_weight,
               loss.occ_weight, loss.flow_origin_weight, loss.replica)
    d = {n: out[i] for i, n in enumerate(LOSS_NAMES)}
    if loss.no_use_warp:
        d["flow_warp_xe"] = 0.0
    d["res"] = out[11:19]
    return d, _metrics_object(out, no_warp)
