"""Diagnostic: device resize (Pillow-exact BILINEAR) throughput and the reverse-ResNet decoder with 8 / 16 epilogue warps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200")); sys.path.insert(0, ROOT)
import torch
from vitad import _lib, ops
from vitad import synth_weights as W

def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

for (B, H, Wd) in [(32, 1024, 1024), (32, 900, 900), (32, 700, 700), (32, 256, 256)]:
    x = torch.randint(0, 256, (B, H, Wd, 3), dtype=torch.uint8, device="cuda")
    t = timeit(lambda: ops.resize_u8(x, 224))
    algo = B * (H * Wd * 3 + 2 * H * 224 * 3 + 3 * 224 * 224)
    print(f"resize {H}x{Wd} -> 224, B={B}: {t*1e3:.1f} us, {algo/t/1e6:.0f} GB/s algorithmic ({B/t*1e3:.0f} img/s)")

from vitad.autoencoders import DecoderResNetVariableEmbeddingSize
dec = DecoderResNetVariableEmbeddingSize(768)
dec.load_state_dict({k[len("decoder."):]: v for k, v in W.make_resnet_decoder_state_dict(43).items()})
dec = dec.cuda().eval()
lat = torch.randn(32, 768, device="cuda") * 0.7
for warps in (0, 8, 16):
    _lib.lib.vitad_set_epilogue_warps(warps)
    with torch.no_grad():
        t = timeit(lambda: dec(lat))
    print(f"reverse-ResNet decoder B=32, epilogue warps forced to {warps or 'default'}: {t:.3f} ms")
_lib.lib.vitad_set_epilogue_warps(0)
