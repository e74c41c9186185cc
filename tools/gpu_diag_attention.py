"""Diagnostic: attention kernel duration for head dim / bias / mask variants (library event profiler)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200"))
import torch
from vitad import _lib, ops
lib = _lib.lib
lib.vitad_profile_enable.argtypes = [C.c_int]; lib.vitad_profile_report.argtypes = [C.c_char_p, C.c_int]; lib.vitad_profile_report.restype = C.c_int

def run(bw, h, t, hd, bias, region, tag):
    q = torch.randn(bw, h, t, hd, device="cuda").half(); k = torch.randn_like(q); vt = torch.zeros(bw, h, hd, 256, device="cuda", dtype=torch.float16)
    vt[..., :t] = torch.randn(bw, h, hd, t, device="cuda").half()
    b = torch.randn(h, t, t, device="cuda") if bias else None
    r = torch.zeros(1, t, dtype=torch.int8, device="cuda") if region else None
    w2t = None
    for _ in range(3): ops.attention(q, k, vt, t, windows=1, bias=b, region=r, win2tok=w2t)
    torch.cuda.synchronize(); lib.vitad_profile_enable(1)
    for _ in range(5):
        ops.attention(q, k, vt, t, windows=1, bias=b, region=r, win2tok=w2t); torch.cuda.synchronize()
    buf = C.create_string_buffer(4096); lib.vitad_profile_report(buf, len(buf)); lib.vitad_profile_enable(0)
    rr = buf.value.decode().split("\n")[0].split()
    print(f"{tag:28s} bw{bw} h{h} t{t} hd{hd}: {float(rr[2])/int(rr[1]):8.1f} us")

run(32, 12, 198, 64, False, False, "deit")
run(32, 12, 196, 64, False, False, "hd64 t196")
run(32, 12, 196, 32, False, False, "hd32 no bias")
run(32, 12, 196, 32, True, False, "hd32 bias")
run(32, 12, 196, 32, True, True, "hd32 bias+region")
run(32, 12, 196, 64, True, False, "hd64 bias")
run(32, 24, 49, 32, True, False, "hd32 t49 bias")
