"""ncu target: three forwards of the 20-step normalizing-flow head at batch 32."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from vitad import synth_weights as W
from vitad.nf import NormalizingFlow
np.random.seed(0)
nf = NormalizingFlow(768, 224, 196, 0.16, 20); nf.load_state_dict(W.make_nf_state_dict(31)); nf = nf.cuda().eval()
x = torch.randn(32, 196, 768, device="cuda")
with torch.no_grad():
    for _ in range(3):
        nf.forward_tokens(x)
torch.cuda.synchronize()
