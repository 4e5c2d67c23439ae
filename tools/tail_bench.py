"""Times the decoder tail (2 x tc_upconv4h + head_tapsum) at batch 16 with CUDA events, per role.
usage: python tools/tail_bench.py [B]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strajnet_b200 import _lib, weights  # noqa: E402
from oracle import strajnet_oracle as O  # noqa: E402  (weights only)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
lib = _lib.lib()
dev = torch.device("cuda")
w = O.make_weights(O.CFG256, seed=0)
dw = {k[len("decoder."):]: v for k, v in w.items() if k.startswith("decoder.")}
pk = weights.Packer(dw, dev, tc=True)
dec = pk.decoder("")
x3 = torch.randn(B * 8, 128, 128, 96, device=dev).to(torch.bfloat16)
f3 = torch.randn(B * 8, 128, 128, 96, device=dev).to(torch.bfloat16)
out = torch.empty(B, 256, 256, 32, dtype=torch.float32, device=dev)
n = lib.sj_decoder_tail_workspace_bytes(B, _lib.SJ_BF16)
ws = torch.empty(n, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream


def run():
    _lib.check(lib.sj_decoder_tail_fwd(x3.data_ptr(), f3.data_ptr(), out.data_ptr(), C.byref(dec), B, 1, _lib.SJ_BF16,
                                       ws.data_ptr(), n, st), "tail")


for _ in range(3):
    run()
torch.cuda.synchronize()
res = {}
for role in ("dec.upconv3", "dec.upconvf1", "dec.outconv"):
    lib.sj_probe_start(role.encode())
    for _ in range(10):
        run()
    ms, cnt = C.c_double(0), C.c_int(0)
    lib.sj_probe_stop(C.byref(ms), C.byref(cnt))
    res[role] = ms.value / max(cnt.value, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
print(f"B={B} tail {e0.elapsed_time(e1) / 10:.4f} ms | " +
      " | ".join(f"{k} {v * 1e3:.1f} us" for k, v in res.items()))
