"""Diagnostic: in-kernel timeline of the fused residual GEMM + LayerNorm kernel (needs the `make TL=1` build; run with
VITAD_LIB=vit-ad_b200/lib/libvitad_tl.so).  Median over CTAs of the cycles between milestones.  usage: gpu_timeline_ln.py [K]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200"))
import torch
from vitad import _lib
lib = _lib.lib
lib.vitad_debug_timeline_ln.argtypes = [C.c_void_p]
M = 6336
for K in ([int(sys.argv[1])] if len(sys.argv) > 1 else [768, 3072]):
  for warps in (8,):
    torch.manual_seed(0)
    a = (torch.randn(M, K) * 0.5).half().cuda(); w = (torch.randn(768, K) * 0.05).half().cuda()
    bias, gamma, beta = torch.zeros(768).cuda(), torch.ones(768).cuda(), torch.zeros(768).cuda()
    x = torch.randn(M, 768).cuda(); h = torch.empty(M, 768, dtype=torch.float16, device="cuda")
    args = _lib.LinearLnArgs()
    args.a, args.w, args.bias, args.m, args.k, args.lda, args.ldw = a.data_ptr(), w.data_ptr(), bias.data_ptr(), M, K, K, K
    args.x, args.gamma, args.beta, args.eps, args.h, args.ldh = x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-6, h.data_ptr(), 768
    run = lambda: _lib.check(lib.vitad_linear_resid_ln_f16(C.byref(args), torch.cuda.current_stream().cuda_stream))
    for _ in range(3): run()
    torch.cuda.synchronize()
    tl = torch.zeros(100, 64, dtype=torch.int64, device="cuda")
    lib.vitad_debug_timeline_ln(tl.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    lib.vitad_debug_timeline_ln(None)
    t = tl.cpu()
    g0 = t[:, 1].min().item(); g1 = t[:, 61].max().item()
    print(f"== K{K} warps{warps}: event {e0.elapsed_time(e1)*1e3:.1f} us; globaltimer first entry -> last exit {(g1-g0)/1e3:.1f} us")
    rel = (t - t[:, :1]).double()
    def med(s): 
        v = rel[:, s][t[:, s] > 0]
        return None if v.numel() == 0 else int(v.median().item())
    names = {2: "griddep_wait done", 3: "first TMA", 4: "last TMA", 8: "first MMA (leaders)", 10: "tmem_full commit", 20: "epi: tmem_full seen",
             21: "epi: resid loads issued", 22: "unit0 resid landed", 23: "unit1", 24: "unit2", 25: "unit3", 26: "unit4", 27: "unit5",
             30: "pass1 done", 31: "stats barrier passed", 32: "x stores read", 40: "pass2 done", 41: "h stores read", 60: "exit"}
    for s, n in names.items():
        print(f"   {n:28s} {med(s)}")
