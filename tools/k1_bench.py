"""K1 (tc_wmsa_kernel) alone at growing tile counts: time per launch and tiles per SM.  Under ncu
(`ncu --set full -k regex:tc_wmsa ...`) this gives the tensor-pipe % of the kernel as a function of the launch size.
usage: python tools/k1_bench.py [B ...]   (64x64 tokens, C = 96: 32 tiles per sample)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import strajnet_b200 as sj  # noqa: E402
from strajnet_b200 import _lib  # noqa: E402

lib = _lib.lib()
for B in [int(a) for a in sys.argv[1:]] or [16, 64, 256]:
    blk = sj.SwinTransformerBlock(96, (64, 64), 3, window_size=8, shift_size=4, dtype="bfloat16")
    blk.build()
    x = torch.randn(B, 4096, 96, device="cuda").to(torch.bfloat16)
    for _ in range(3):
        y = blk(x)
    torch.cuda.synchronize()
    import ctypes as C
    res = {}
    for role in ("",):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            y = blk(x)
        e1.record()
        torch.cuda.synchronize()
    print(f"B={B}: {B * 32} tiles ({B * 32 / 148:.1f} per SM), block (ln_stats + tc_wmsa + tc_mlp96) {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")
