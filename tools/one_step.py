"""Three batch-16 bf16 forwards as plain stream launches (for `ncu` launch lists: the last forward is the warm one)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import strajnet_b200 as sj  # noqa: E402
from strajnet_b200 import _lib  # noqa: E402
from bench import CFG256, synth_inputs  # noqa: E402

dev = torch.device("cuda", 0)
m = sj.STrajNet(CFG256, fg_msa=True, fg=True, large_ogm=False, dtype="bfloat16", device=dev)
m.build()
inp = {k: v.to(dev) for k, v in synth_inputs(16).items()}
out = torch.empty(16, 256, 256, 32, device=dev)
lib = _lib.lib()
for i in range(3):
    lib.sj_launch_count(1)
    m.forward_into(out, inp["ogm"], inp["map_img"], inp["obs"], inp["occ"], inp["flow"])
    torch.cuda.synchronize()
print("launches per forward:", lib.sj_launch_count(0))
