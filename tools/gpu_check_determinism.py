"""Diagnostic: run-to-run and batch-composition determinism of the encoder and heads."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from vitad import synth_weights as W
from vitad.encoders import EncoderDeit
from vitad.nf import NormalizingFlow
from vitad.mdn import GaussianMixtureDensityNetwork
enc = EncoderDeit(224); enc.load_state_dict(W.make_deit_state_dict(11, stress=True)); enc = enc.cuda().eval()
np.random.seed(0)
nf = NormalizingFlow(768, 224, 196, 0.16, 20); nf.load_state_dict(W.make_nf_state_dict(31, stress=True)); nf = nf.cuda().eval()
head = GaussianMixtureDensityNetwork(768, 768, 100); head.load_state_dict(W.make_mdn_state_dict(21, 100, stress=True)); head = head.cuda().eval()
imgs = W.synthetic_images(3, 32).cuda(); other = W.synthetic_images(4, 21).cuda()
g = torch.randn(32, 196, 100, device="cuda")
with torch.no_grad():
    a = enc(imgs).patch_embedding.clone()
    _ = enc(other)                       # different batch size in between (workspace reuse)
    b = enc(imgs).patch_embedding.clone()
    print("deit run-to-run max diff", (a - b).abs().max().item())
    c = enc(imgs[:7]).patch_embedding
    print("deit batch-composition max diff (first 7 images)", (a[:7] - c).abs().max().item())
    r1 = nf.forward_tokens(a).anomaly_score_map.clone(); _ = nf.forward_tokens(enc(other).patch_embedding)
    r2 = nf.forward_tokens(a).anomaly_score_map.clone()
    print("nf run-to-run max diff", (r1 - r2).abs().max().item())
    r3 = nf.forward_tokens(a[:7].contiguous()).anomaly_score_map
    print("nf batch-composition max diff", (r1[:7] - r3).abs().max().item())
    L1 = head.patch_log_likelihood(a, g).clone(); L2 = head.patch_log_likelihood(a, g).clone()
    print("gmm run-to-run max diff", (L1 - L2).abs().max().item())
    L3 = head.patch_log_likelihood(a[:7].contiguous(), g[:7].contiguous())
    print("gmm batch-composition max diff", (L1[:7] - L3).abs().max().item())
