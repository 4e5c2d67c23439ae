"""In-step duration of each forward role (sj_probe_start / sj_probe_stop: CUDA events around the launches whose role starts
with a prefix), batch-16 bf16 forward issued as plain stream launches.  One line per role: launches per step, us per step."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import strajnet_b200 as sj  # noqa: E402
from strajnet_b200 import _lib  # noqa: E402
from bench import CFG256, synth_inputs  # noqa: E402

ROLES = ["enc", "fgmsa", "traj", "dec.upconv0", "dec.res0", "dec.upconv1", "dec.res1", "dec.resf", "dec.upconv2",
         "dec.upconv3", "dec.upconvf0", "dec.upconvf1", "dec.outconv"]


def main():
    B, steps = 16, int(os.environ.get("STEPS", "10"))
    dev = torch.device("cuda", 0)
    lib = _lib.lib()
    m = sj.STrajNet(CFG256, fg_msa=True, fg=True, large_ogm=False, dtype="bfloat16", device=dev)
    m.build()
    inp = {k: v.to(dev) for k, v in synth_inputs(B).items()}
    out = torch.empty(B, 256, 256, 32, device=dev)
    roles = os.environ.get("ROLES", ",".join(ROLES)).split(",")
    s = torch.cuda.Stream(dev)
    tot = 0.0
    with torch.cuda.stream(s):
        for _ in range(5):
            m.forward_into(out, inp["ogm"], inp["map_img"], inp["obs"], inp["occ"], inp["flow"])
        for role in roles:
            lib.sj_probe_start(role.encode())
            for _ in range(steps):
                m.forward_into(out, inp["ogm"], inp["map_img"], inp["obs"], inp["occ"], inp["flow"])
            ms, n = ctypes.c_double(0), ctypes.c_int(0)
            lib.sj_probe_stop(ctypes.byref(ms), ctypes.byref(n))
            us = ms.value / steps * 1e3
            tot += us
            print(f"{role:14s} {n.value // steps:3d} launches  {us:8.1f} us/step", flush=True)
    print(f"{'sum':14s}               {tot:8.1f} us/step")


if __name__ == "__main__":
    main()
