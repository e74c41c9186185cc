"""ncu target: two forwards of the reverse-ResNet decoder at batch 32 (capture the second with -s)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200")); sys.path.insert(0, ROOT)
import torch
from vitad import synth_weights as W
from vitad.autoencoders import DecoderResNetVariableEmbeddingSize
dec = DecoderResNetVariableEmbeddingSize(768)
dec.load_state_dict({k[len("decoder."):]: v for k, v in W.make_resnet_decoder_state_dict(43).items()})
dec = dec.cuda().eval()
lat = torch.randn(32, 768, device="cuda") * 0.7
with torch.no_grad():
    for _ in range(2):
        dec(lat)
torch.cuda.synchronize()
