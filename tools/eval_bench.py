"""Timing of the fused validation pass (sj_ogm_flow_eval_fwd: OGMFlow_loss + occupancy-flow metrics) at batch 16,
256x256x8: algorithmic bytes (every input read once) / CUDA-event time = achieved HBM GB/s."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strajnet_b200 import evaluation as V  # noqa: E402


def main():
    B, H = 16, 256
    g = torch.Generator(device="cuda").manual_seed(0)
    pred = torch.randn(B, H, H, 32, device="cuda", generator=g)
    obs = (torch.rand(B, 8, H, H, device="cuda", generator=g) < 0.03).float()
    occ = (torch.rand(B, 8, H, H, device="cuda", generator=g) < 0.01).float()
    flow = torch.randn(B, 8, H, H, 2, device="cuda", generator=g) * (obs + occ).clamp(0, 1)[..., None]
    org = (torch.rand(B, 8, H, H, device="cuda", generator=g) < 0.04).float()
    loss = V.OGMFlow_loss(None, use_gt=True, use_focal_loss=False)  # train.py:195-196
    big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # L2 flush between iterations
    for _ in range(3):
        V.evaluate(pred, obs, occ, flow, org, loss=loss)
    ms = []
    for _ in range(10):
        big.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = V._run(pred, obs, occ, flow, org, loss.flags() | V.FLAG["loss"] | V.FLAG["metrics"])
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    med = ms[len(ms) // 2]
    nbytes = pred.numel() * 4 + 3 * obs.numel() * 4 + flow.numel() * 4
    peak = 6551.0
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    print(json.dumps({"kernel": "eval_pass_kernel + eval_finalize_kernel (+ memset)", "batch": B, "ms": med, "ms_min": ms[0],
                      "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / med / 1e6, "hbm_peak_gbs": peak,
                      "frac": nbytes / med / 1e6 / peak, "frames_per_s": B / med * 1e3,
                      "out": [round(x, 5) for x in out.tolist()[:11]]}))


if __name__ == "__main__":
    main()
