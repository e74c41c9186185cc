"""Diagnostic: fused GMM kernel (batch 32, K=100) with the side CTA-pair launch on the SMs the 4-CTA clusters leave idle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200")); sys.path.insert(0, ROOT)
import torch
from vitad import _lib
from vitad import synth_weights as W
from vitad.mdn import GaussianMixtureDensityNetwork

B, K = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 100
head = GaussianMixtureDensityNetwork(768, 768, K); head.load_state_dict(W.make_mdn_state_dict(21, K)); head = head.cuda().eval()
x = torch.randn(B, 196, 768, device="cuda"); gn = torch.randn(B, 196, K, device="cuda")
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
import subprocess
def clock():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    except Exception:
        return "?"
ref = None
with torch.no_grad():
    for rep in range(3):
        for split in (0, 72, 70, 0, 64, 72, 48, 70):
            _lib.lib.vitad_set_gmm_split(split)
            L = head.patch_log_likelihood(x, gn)
            if ref is None: ref = L.clone()
            same = torch.equal(L, ref)
            t = timeit(lambda: head.patch_log_likelihood(x, gn), iters=100)
            print(f"rep {rep} split {split:4d}: head {t*1e3:7.1f} us   bit-identical: {same}   [{clock()}]")
_lib.lib.vitad_set_gmm_split(-1)
