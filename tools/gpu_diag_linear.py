"""Diagnostic (not a test): prints error statistics and timings of the GEMM on the GPU box."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200"))
import torch
from vitad import _lib, ops

torch.manual_seed(0)
print("device", torch.cuda.get_device_name(0), "abi", _lib.lib.vitad_abi_version())
for (m, n, k, bn) in [(128, 256, 64, 256), (128, 128, 64, 128), (256, 512, 128, 256), (6336, 768, 768, 256),
                      (6336, 2304, 768, 256), (6336, 3072, 768, 256), (6336, 768, 3072, 256), (6336, 768, 768, 128)]:
    a = (torch.randn(m, k) * 0.5).to(torch.float16).cuda()
    w = (torch.randn(n, k) * 0.05).to(torch.float16).cuda()
    b = (torch.randn(n) * 0.1).cuda()
    out = ops.linear(a, w, b, _lib.EPI_F32, block_n=bn)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + b
    d = (out - ref).abs()
    print(f"M{m} N{n} K{k} bn{bn}: max_err {d.max().item():.3e} ref_max {ref.abs().max().item():.3e} "
          f"bad_frac {(d > 1e-3).float().mean().item():.4f}")
    if d.max().item() > 1e-2:
        bad = (d > 1e-3).nonzero()
        print("  first bad idx", bad[:5].tolist(), "rows bad", bad[:, 0].unique()[:10].tolist(), "cols bad",
              bad[:, 1].unique()[:10].tolist())
    # timing
    for _ in range(3):
        ops.linear(a, w, b, _lib.EPI_BIAS_F16, block_n=bn)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = 20
    for _ in range(iters):
        ops.linear(a, w, b, _lib.EPI_BIAS_F16, block_n=bn)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"   {ms*1e3:.1f} us  {2*m*n*k/ms/1e9:.1f} TFLOP/s")
    t0 = time.perf_counter()
    for _ in range(iters):
        torch.nn.functional.linear(a, w)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.nn.functional.linear(a, w)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"   cuBLAS {ms*1e3:.1f} us  {2*m*n*k/ms/1e9:.1f} TFLOP/s")
