"""Diagnostic: steady-state main-loop rate of the GEMM kernels on large square problems."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200"))
import torch
from vitad import _lib, ops

def t(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

for (m, n, k) in [(8192, 8192, 8192), (6336, 3072, 768), (6336, 2304, 768), (6336, 768, 768), (6336, 768, 3072)]:
    a = (torch.randn(m, k) * 0.5).half().cuda(); w = (torch.randn(n, k) * 0.05).half().cuda(); b = torch.zeros(n).cuda()
    out = torch.empty(m, n, device="cuda", dtype=torch.float16)
    for pair in (0, 1):
        for bn in (96, 128, 256):
            _lib.lib.vitad_set_cta_pair(pair)
            ms = t(lambda: ops.linear(a, w, b, _lib.EPI_BIAS_F16, out=out, block_n=bn))
            print(f"M{m} N{n} K{k} pair={pair} bn={bn}: {ms*1e3:8.1f} us {2*m*n*k/ms/1e9:8.1f} TFLOP/s")
    ms = t(lambda: torch.nn.functional.linear(a, w))
    print(f"M{m} N{n} K{k} cuBLAS        : {ms*1e3:8.1f} us {2*m*n*k/ms/1e9:8.1f} TFLOP/s")
_lib.lib.vitad_set_cta_pair(1)
