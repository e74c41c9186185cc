"""Diagnostic: in-kernel timeline of the attention kernel (needs `make TL=1`; VITAD_LIB=vit-ad_b200/lib/libvitad_tl.so)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200"))
import torch
from vitad import _lib, ops

lib = _lib.lib
lib.vitad_debug_timeline_attention.argtypes = [C.c_void_p]
B, H, T, hd, Tpad = 32, 12, 198, 64, 256
torch.manual_seed(0)
q = (torch.randn(B, H, T, hd) * 0.3).half().cuda(); k = (torch.randn(B, H, T, hd)).half().cuda()
vt = torch.zeros(B, H, hd, Tpad).half().cuda(); vt[..., :T] = torch.randn(B, H, hd, T).half().cuda()
out = torch.empty(B * T, H * hd, dtype=torch.float16, device="cuda")
a = _lib.AttentionArgs()
a.q, a.k, a.vt, a.out = q.data_ptr(), k.data_ptr(), vt.data_ptr(), out.data_ptr()
a.batch_windows, a.heads, a.tokens, a.tokens_pad, a.head_dim, a.windows = B, H, T, Tpad, hd, 1
run = lambda: _lib.check(lib.vitad_attention_f16(C.byref(a), torch.cuda.current_stream().cuda_stream))
for _ in range(3): run()
torch.cuda.synchronize()
nblk = 2 * B * H
tl = torch.zeros(nblk, 64, dtype=torch.int64, device="cuda")
lib.vitad_debug_timeline_attention(tl.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
lib.vitad_debug_timeline_attention(None)
t = tl.cpu()
g0 = t[:, 1].min().item()
print(f"event {e0.elapsed_time(e1)*1e3:.1f} us; first entry -> last exit {(t[:,61].max().item()-g0)/1e3:.1f} us")
names = {2: "prologue", 3: "qk_arrived", 5: "S_ready", 6: "pass1", 7: "pass2", 8: "sync", 13: "pv_issued", 9: "O_ready", 11: "pre_ld", 12: "ld_done", 10: "stored", 60: "exit"}
starts = ((t[:, 1] - g0).float() / 1e3)
ends = ((t[:, 61] - g0).float() / 1e3)
print("entry time (us) quantiles:", [round(float(starts.quantile(x)), 1) for x in (0, .25, .5, .75, 1)])
print("exit  time (us) quantiles:", [round(float(ends.quantile(x)), 1) for x in (0, .25, .5, .75, 1)])
print("CTA lifetime (us) median:", float((ends - starts).median()))
for blk in (0, 1, 300, 301, 766, 767):
    r = t[blk]; z = r[0].item()
    print(f" blk {blk} (entry {starts[blk]:.1f} us):", {n: r[s].item() - z for s, n in names.items()})
