"""ncu target: three forwards of the EsViT Swin-T encoder at batch 32 (skip the first two with -s)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200")); sys.path.insert(0, ROOT)
import torch
from vitad import synth_weights as W
from vitad.encoders import EncoderEsVit
es = EncoderEsVit(224, requires_grad=True); es.load_state_dict(W.make_esvit_state_dict(51)); es = es.cuda().eval()
imgs = W.synthetic_images(1, 32).cuda()
with torch.no_grad():
    for _ in range(3):
        es(imgs)
torch.cuda.synchronize()
