"""Diagnostic: 8 vs 16 epilogue warps (vitad_set_epilogue_warps) on the Swin-T forward and the DeiT + GMM step at batch 32."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200")); sys.path.insert(0, ROOT)
import torch
from vitad import _lib
from vitad import synth_weights as W
from vitad.encoders import EncoderDeit, EncoderEsVit

def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

B = 32
imgs = W.synthetic_images(1, B).cuda()
with torch.no_grad():
    es = EncoderEsVit(224, requires_grad=True); es.load_state_dict(W.make_esvit_state_dict(51)); es = es.cuda().eval()
    deit = EncoderDeit(224); deit.load_state_dict(W.make_deit_state_dict(11)); deit = deit.cuda().eval()
    for rep in range(2):
        for warps in (0, 8, 16):
            _lib.lib.vitad_set_epilogue_warps(warps)
            t1 = timeit(lambda: es(imgs)); t2 = timeit(lambda: deit(imgs))
            print(f"rep {rep} epilogue warps {warps or 'default':>7}: Swin-T forward {t1:.3f} ms, DeiT forward {t2:.3f} ms")
_lib.lib.vitad_set_epilogue_warps(0)
