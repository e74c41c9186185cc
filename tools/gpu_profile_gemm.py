"""Diagnostic: GPU-side duration (event pair inside the library) of single GEMM launches, small shapes."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200"))
import torch
from vitad import _lib, ops
lib = _lib.lib
lib.vitad_profile_enable.argtypes = [C.c_int]; lib.vitad_profile_report.argtypes = [C.c_char_p, C.c_int]; lib.vitad_profile_report.restype = C.c_int

def run(m, n, k, epi, bn, pair=1):
    lib.vitad_set_cta_pair(pair)
    a = (torch.randn(m, k) * 0.5).half().cuda(); w = (torch.randn(n, k) * 0.05).half().cuda(); b = torch.zeros(n).cuda()
    kw = {}
    if epi == _lib.EPI_RESIDUAL_F32:
        r = torch.randn(m, n, device="cuda"); kw = dict(out=r, resid=r)
    for _ in range(3): ops.linear(a, w, b, epi, block_n=bn, **kw)
    torch.cuda.synchronize()
    lib.vitad_profile_enable(1)
    for _ in range(10):
        ops.linear(a, w, b, epi, block_n=bn, **kw); torch.cuda.synchronize()
    buf = C.create_string_buffer(4096); lib.vitad_profile_report(buf, len(buf)); lib.vitad_profile_enable(0)
    r = buf.value.decode().split("\n")[0].split()
    print(f"M{m:5d} N{n:5d} K{k:5d} epi{epi} bn{bn:3d} pair{pair}: {float(r[2])/int(r[1]):7.1f} us")

E = _lib
for pair in (1, 0):
    run(256, 96, 64, E.EPI_F32, 96, pair)
    run(256, 256, 64, E.EPI_BIAS_F16, 256, pair)
    run(6336, 768, 64, E.EPI_BIAS_F16, 96, pair)
    run(6336, 768, 64, E.EPI_RESIDUAL_F32, 96, pair)
    run(6336, 768, 768, E.EPI_BIAS_F16, 96, pair)
    run(6336, 768, 768, E.EPI_RESIDUAL_F32, 96, pair)
    run(6336, 768, 768, E.EPI_RESIDUAL_F32, 256, pair)
    run(6336, 3072, 768, E.EPI_BIAS_F16, 96, pair)
    run(6336, 3072, 768, E.EPI_BIAS_GELU_F16, 96, pair)
    run(6336, 3072, 768, E.EPI_BIAS_GELU_F16, 256, pair)
    run(6336, 768, 3072, E.EPI_RESIDUAL_F32, 96, pair)
lib.vitad_set_cta_pair(1)
