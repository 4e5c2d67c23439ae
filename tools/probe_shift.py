"""Needs a library built with SJ_DEBUG_PROBES=1 (`SJ_DEBUG_PROBES=1 python -m strajnet_b200.build --force`): the probe entry
point sj_debug_gemm_shift is not part of the product ABI."""
import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strajnet_b200 import _lib
L = _lib.lib()
import ctypes as C
L.sj_debug_gemm_shift.restype = C.c_int
L.sj_debug_gemm_shift.argtypes = [C.c_void_p] * 3 + [C.c_int] * 5 + [C.c_void_p]
torch.manual_seed(0)
M, N, K = 384, 64, 128
x = torch.randn(M + 256, K, device='cuda').to(torch.bfloat16)
w = (torch.randn(N, K, device='cuda') * K ** -0.5).to(torch.bfloat16)
for shift in (0, 1, 2, 3, 4, 7, 8, 9, 16, 17, 64, 65, 100):
    for bo in (0, 1):
        y = torch.zeros(M, N, device='cuda', dtype=torch.bfloat16)
        st = L.sj_debug_gemm_shift(x.data_ptr(), y.data_ptr(), w.data_ptr(), M, N, K, shift, bo, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ref = (x[shift:shift + M].float() @ w.float().t())
        err = (y.float() - ref).abs().max().item()
        print(f"shift {shift:3d} base_offset {bo}: status {st} max err {err:.4f}")
