"""Time the un-grouped tcgen05 GEMM shapes of the batch-16 step for every legal n-tile width (SJ_TCG_BN), to check the
width heuristic of tc_gemm.cu.  One line per shape: us per launch by BN, the heuristic's own choice marked with *."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import strajnet_b200 as sj  # noqa: E402

SHAPES = [  # (what, M, N, K, act)
    ("L1 fc1", 16384, 768, 192, "gelu"), ("L2 qkv / fg qkv", 4096, 1152, 384, None), ("L2 proj / fg out", 4096, 384, 384, None),
    ("L2 fc1", 4096, 1536, 384, "gelu"), ("L2 fc2", 4096, 384, 1536, None), ("L1 qkv", 16384, 576, 192, None),
]


def time_layer(layer, x, n=40):
    """us per launch with the launches replayed from a CUDA graph (eager calls are bound by ~13 us of Python each)"""
    for _ in range(3):
        layer(x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            layer(x)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * n) * 1e3


def main():
    dev = torch.device("cuda", 0)
    for what, M, N, K, act in SHAPES:
        layer = sj.Dense(N, K, activation=act, dtype="bfloat16")
        g = torch.Generator().manual_seed(1)
        layer.set_weights({"kernel": torch.randn(K, N, generator=g) * K ** -0.5, "bias": torch.randn(N, generator=g) * 0.1})
        x = torch.randn(M, K, generator=g).to(dev, torch.bfloat16)
        os.environ.pop("SJ_TCG_BN", None)
        time_layer(layer, x)
        base = time_layer(layer, x)
        out = [f"{what:18s} M={M:5d} N={N:4d} K={K:4d}  heuristic {base:6.1f} us |"]
        for bn in range(256, 31, -16):
            if N % bn:
                continue
            os.environ["SJ_TCG_BN"] = str(bn)
            out.append(f" {bn}:{time_layer(layer, x):5.1f}")
        os.environ.pop("SJ_TCG_BN", None)
        print("".join(out), flush=True)


if __name__ == "__main__":
    main()
