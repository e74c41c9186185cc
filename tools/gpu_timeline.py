"""Diagnostic: in-kernel timeline of the CTA-pair GEMM (needs the `make TL=1` build; run with
VITAD_LIB=vit-ad_b200/lib/libvitad_tl.so).  Prints, for a few clusters, when each role reached its milestones
(cycles since the CTA's entry) and the whole-grid span.  usage: gpu_timeline.py name[,name] [bn]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200"))
import torch
from vitad import _lib, ops

lib = _lib.lib
lib.vitad_debug_timeline.argtypes = [C.c_void_p]
M = 6336
shapes = {"qkv": (2304, 768, _lib.EPI_BIAS_F16), "proj": (768, 768, _lib.EPI_RESIDUAL_F32),
          "fc1": (3072, 768, _lib.EPI_BIAS_GELU_F16), "fc2": (768, 3072, _lib.EPI_RESIDUAL_F32)}
names = sys.argv[1].split(",") if len(sys.argv) > 1 else list(shapes)
bn = int(sys.argv[2]) if len(sys.argv) > 2 else 0
torch.manual_seed(0)
for name in names:
    n, k, epi = shapes[name]
    a = (torch.randn(M, k) * 0.5).half().cuda(); w = (torch.randn(n, k) * 0.05).half().cuda(); b = torch.zeros(n).cuda()
    x = torch.randn(M, n).cuda(); oh = torch.empty(M, n, dtype=torch.float16, device="cuda")
    run = (lambda: ops.linear(a, w, b, epi, out=x, resid=x, block_n=bn)) if epi == _lib.EPI_RESIDUAL_F32 else \
          (lambda: ops.linear(a, w, b, epi, out=oh, block_n=bn))
    for _ in range(3): run()
    torch.cuda.synchronize()
    tl = torch.zeros(148, 64, dtype=torch.int64, device="cuda")
    lib.vitad_debug_timeline(tl.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    lib.vitad_debug_timeline(None)
    t = tl.cpu()
    g0 = t[:, 1][t[:, 1] > 0].min().item(); g1 = t[:, 61].max().item()
    print(f"== {name} N{n} K{k} bn{bn}: event {e0.elapsed_time(e1)*1e3:.1f} us; globaltimer first entry -> last exit {(g1-g0)/1e3:.1f} us; "
          f"entry skew {(t[:,1][t[:,1]>0].max().item()-g0)/1e3:.1f} us")
    for cta in (0, 1, 74, 146):
        r = t[cta]; z = r[0].item()
        rel = lambda s: (r[s].item() - z) if r[s].item() else None
        print(f" cta {cta}: prologue {rel(2)} first_tma {rel(3)} last_tma {rel(4)} exit {rel(60)} (cycles)")
        if cta % 2 == 0:
            print("   mma  [acc_free, first_full, commit]:", [(rel(8+3*i), rel(9+3*i), rel(10+3*i)) for i in range(12) if r[10+3*i].item()])
        print("   epi  [acc_full, drained]:", [(rel(44+2*i), rel(45+2*i)) for i in range(8) if r[45+2*i].item()])
