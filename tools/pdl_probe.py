"""Step time of the batch-16 bf16 forward (CUDA-graph replay and plain stream launches) for each programmatic-dependent-
launch mask (sj_set_pdl): which class of launches gains from overlapping its prologue with the predecessor's tail."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import strajnet_b200 as sj  # noqa: E402
from strajnet_b200 import _lib  # noqa: E402
from bench import CFG256, synth_inputs  # noqa: E402


def main():
    B, steps = 16, int(os.environ.get("STEPS", "40"))
    dev = torch.device("cuda", 0)
    lib = _lib.lib()
    m = sj.STrajNet(CFG256, fg_msa=True, fg=True, large_ogm=False, dtype="bfloat16", device=dev)
    m.build()
    inp = {k: v.to(dev) for k, v in synth_inputs(B).items()}
    out = torch.empty(B, 256, 256, 32, device=dev)
    s = torch.cuda.Stream(dev)
    masks = [int(x) for x in os.environ.get("MASKS", "0,1,2,3,0,3").split(",")]
    with torch.cuda.stream(s):
        for mask in masks:
            lib.sj_set_pdl(mask)
            m._graphs.clear()
            res = {}
            for graph in (True, False):
                for _ in range(5):
                    m.forward_into(out, inp["ogm"], inp["map_img"], inp["obs"], inp["occ"], inp["flow"], graph=graph)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.synchronize()
                e0.record()
                for _ in range(steps):
                    m.forward_into(out, inp["ogm"], inp["map_img"], inp["obs"], inp["occ"], inp["flow"], graph=graph)
                e1.record()
                s.synchronize()
                res[graph] = e0.elapsed_time(e1) / steps
            print(f"pdl mask {mask}: graph {res[True]:.4f} ms/step, stream {res[False]:.4f} ms/step", flush=True)


if __name__ == "__main__":
    main()
