#!/usr/bin/env python
"""How much of the step is launch gaps?  Times the batch-16 bf16 forward as plain stream launches and as a replayed
CUDA graph (the C ABI is capture-safe).  python tools/graph_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import strajnet_b200 as sj  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B = 16
    model = sj.STrajNet(bench.CFG256, fg_msa=True, fg=True, large_ogm=False, dtype="bfloat16", device=dev)
    model.build()
    inp = {k: v.to(dev) for k, v in bench.synth_inputs(B, 256, seed=0).items()}
    out = torch.empty(B, 256, 256, 32, dtype=torch.float32, device=dev)

    def fwd():
        model.forward_into(out, inp["ogm"], inp["map_img"], inp["obs"], inp["occ"], inp["flow"])

    def timed(fn, n=20):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    s = torch.cuda.Stream(dev)
    with torch.cuda.stream(s):
        for _ in range(5):
            fwd()
        ms_stream = timed(fwd)
        ref = out.clone()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fwd()
        for _ in range(3):
            g.replay()
        ms_graph = timed(g.replay)
        same = torch.equal(ref, out)
    print(f"stream launches: {ms_stream:.3f} ms/step   graph replay: {ms_graph:.3f} ms/step   identical output: {same}")


if __name__ == "__main__":
    main()
