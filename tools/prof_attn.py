"""Profiling target: the DeiT-shaped attention kernel alone (B=32, H=12, T=198, hd=64)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200"))
import torch
from vitad import ops
B, H, T, hd, Tpad = 32, 12, 198, 64, 256
torch.manual_seed(0)
q = (torch.randn(B, H, T, hd) * 0.3).half().cuda(); k = torch.randn(B, H, T, hd).half().cuda()
vt = torch.zeros(B, H, hd, Tpad).half().cuda(); vt[..., :T] = torch.randn(B, H, hd, T).half().cuda()
for _ in range(4):
    out = ops.attention(q, k, vt, T)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = ops.attention(q, k, vt, T)
e1.record(); torch.cuda.synchronize()
print("attention avg us (launch-bound loop):", e0.elapsed_time(e1) / 20 * 1e3)
