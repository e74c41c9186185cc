"""Short profiling target: two steps of the bs-32 scoring path (DeiT + GMM K=100 + tail)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200")); sys.path.insert(0, ROOT)
import torch
from vitad import synth_weights as W
from vitad import ops
from vitad.encoders import EncoderDeit
from vitad.mdn import GaussianMixtureDensityNetwork

B, K = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 100
enc = EncoderDeit(224); enc.load_state_dict(W.make_deit_state_dict(11)); enc = enc.cuda().eval()
head = GaussianMixtureDensityNetwork(768, 768, K); head.load_state_dict(W.make_mdn_state_dict(21, K)); head = head.cuda().eval()
imgs = W.synthetic_images(1, B).cuda(); gn = torch.randn(B, 196, K, device="cuda")
with torch.no_grad():
    for _ in range(2):
        f = enc(imgs); prob, sc = head.score(f.patch_embedding, gn)
        ops.bilinear_up(prob.view(-1, 14, 14), 224, True, post_one_minus=True)
torch.cuda.synchronize()
print("ok", sc[:4].tolist())
