"""Diagnostic: per-shape timing of the encoder GEMMs (epilogue x BLOCK_N), launch-overhead free (each variant is
captured in a CUDA graph of REP back-to-back launches), next to cuBLAS on the same shape.
usage: python tools/gpu_bench_gemm.py [bn,bn,...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200"))
import torch
from vitad import _lib, ops

REP = 20
bns = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [96, 128, 256]
torch.manual_seed(0)
M = 6336
shapes = [("qkv", 2304, 768), ("proj", 768, 768), ("fc1", 3072, 768), ("fc2", 768, 3072)]


def timed_graph(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); fn()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(REP):
                fn()
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(5):
            e0.record(s); g.replay(); e1.record(s); s.synchronize()
            best = min(best, e0.elapsed_time(e1) / REP * 1e3)
    return best


for name, n, k in shapes:
    a = (torch.randn(M, k) * 0.5).half().cuda()
    w = (torch.randn(n, k) * 0.05).half().cuda()
    b = (torch.randn(n) * 0.1).cuda()
    x = torch.randn(M, n).cuda()
    flops = 2.0 * M * n * k
    t = timed_graph(lambda: torch.nn.functional.linear(a, w))
    print(f"{name:5s} N{n} K{k}: cuBLAS {t:6.1f} us {flops/t/1e6:7.1f} TFLOP/s")
    out_h = torch.empty(M, n, dtype=torch.float16, device="cuda")
    q = torch.empty(32, 12, 198, 64, dtype=torch.float16, device="cuda")
    kk = torch.empty_like(q)
    vt = torch.zeros(32, 12, 64, 256, dtype=torch.float16, device="cuda")
    for bn, ew in [(b_, e_) for b_ in bns for e_ in (8, 16)]:
        if n % bn:
            continue
        _lib.lib.vitad_set_epilogue_warps(ew)
        res = []
        for epi, label in ((_lib.EPI_BIAS_F16, "bias"), (_lib.EPI_BIAS_GELU_F16, "gelu"), (_lib.EPI_RESIDUAL_F32, "resid")):
            if epi == _lib.EPI_RESIDUAL_F32:
                f = lambda: ops.linear(a, w, b, epi, out=x, resid=x, block_n=bn)
            else:
                f = lambda: ops.linear(a, w, b, epi, out=out_h, block_n=bn)
            t = timed_graph(f)
            res.append(f"{label} {t:6.1f} us {flops/t/1e6:6.1f}")
        if name == "qkv":
            t = timed_graph(lambda: ops.linear_qkv(a, w, b, 32, 198, 12, 256, q, kk, vt, 0.125, block_n=bn))
            res.append(f"qkv {t:6.1f} us {flops/t/1e6:6.1f}")
        print(f"   bn{bn:3d} ew{ew:2d}: " + " | ".join(res))
    _lib.lib.vitad_set_epilogue_warps(0)
# fixed overhead: one k-block, one wave
a = (torch.randn(M, 64) * 0.5).half().cuda(); w = (torch.randn(768, 64) * 0.05).half().cuda(); b = torch.zeros(768).cuda()
o = torch.empty(M, 768, dtype=torch.float16, device="cuda")
for bn in bns:
    print(f"K=64 N=768 bn{bn}: {timed_graph(lambda: ops.linear(a, w, b, _lib.EPI_BIAS_F16, out=o, block_n=bn)):.1f} us")
