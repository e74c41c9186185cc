"""Diagnostic: timings of the other BASELINE configs at batch 32 (NF head, EsViT + GMM-130, recon tail)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from vitad import synth_weights as W
from vitad import _lib, ops
from vitad.encoders import EncoderDeit, EncoderEsVit
from vitad.mdn import GaussianMixtureDensityNetwork
from vitad.nf import NormalizingFlow
lib = _lib.lib
lib.vitad_profile_enable.argtypes = [C.c_int]; lib.vitad_profile_report.argtypes = [C.c_char_p, C.c_int]; lib.vitad_profile_report.restype = C.c_int

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def profile(fn, n=3):
    lib.vitad_profile_enable(1)
    for _ in range(n): fn()
    buf = C.create_string_buffer(1 << 16); lib.vitad_profile_report(buf, len(buf)); lib.vitad_profile_enable(0)
    rows = [l.split() for l in buf.value.decode().strip().split("\n")]
    for r in sorted(rows, key=lambda r: -float(r[2]))[:14]:
        print(f"      {r[0]:34s} n/step {int(r[1])/n:5.1f} per-step {float(r[2])/n:8.1f} us")

B = 32
imgs = W.synthetic_images(1, B).cuda()
with torch.no_grad():
    deit = EncoderDeit(224); deit.load_state_dict(W.make_deit_state_dict(11)); deit = deit.cuda().eval()
    np.random.seed(0)
    nf = NormalizingFlow(768, 224, 196, 0.16, 20); nf.load_state_dict(W.make_nf_state_dict(31)); nf = nf.cuda().eval()
    x = deit(imgs).patch_embedding
    t = timeit(lambda: nf.forward_tokens(x)); print(f"NF head (20 steps) B={B}: {t:.3f} ms")
    profile(lambda: nf.forward_tokens(x))
    t = timeit(lambda: nf.forward_tokens(deit(imgs).patch_embedding)); print(f"config 2  DeiT + NF: {t:.3f} ms -> {B/t*1e3:.0f} img/s")
    es = EncoderEsVit(224, requires_grad=True); es.load_state_dict(W.make_esvit_state_dict(51)); es = es.cuda().eval()
    t = timeit(lambda: es(imgs)); print(f"EsViT Swin-T forward B={B}: {t:.3f} ms -> {B/t*1e3:.0f} img/s ({9.78e9*B/t/1e9:.1f} TFLOP/s)")
    profile(lambda: es(imgs))
    head = GaussianMixtureDensityNetwork(768, 768, 130); head.load_state_dict(W.make_mdn_state_dict(21, 130)); head = head.cuda().eval()
    gn = torch.randn(B, 49, 130, device="cuda")
    def c3():
        f = es(imgs); prob, sc = head.score(f.patch_embedding, gn); ops.bilinear_up(prob.view(-1, 7, 7), 224, True, post_one_minus=True)
    t = timeit(c3); print(f"config 3  EsViT + GMM-130: {t:.3f} ms -> {B/t*1e3:.0f} img/s")
    recon = torch.tanh(torch.randn(B, 3, 224, 224, device="cuda"))
    t = timeit(lambda: ops.l2_map_score(recon, imgs), iters=50); bytes_ = B * (2 * 3 + 1) * 224 * 224 * 4
    print(f"recon L2 map+amax B={B}: {t*1e3:.1f} us -> {bytes_/t/1e6:.0f} GB/s algorithmic")
    prob = torch.rand(B, 14, 14, device="cuda")
    t = timeit(lambda: ops.bilinear_up(prob, 224, True, post_one_minus=True), iters=50)
    print(f"bilinear 14->224 B={B}: {t*1e3:.1f} us -> {B*224*224*4/t/1e6:.0f} GB/s written")
    # config 4: DeiT + small CNN decoder (vitad_cnn_decoder_forward) + L2 map
    from vitad.model_helper import get_model
    ae = get_model("ae_deit_small", 224, requires_grad=True).cuda().eval()
    def c4():
        o = ae(imgs); return ae.anomaly_map_and_score(o.reconstruction, imgs)
    t = timeit(c4); print(f"config 4  DeiT + CNN decoder + L2 map: {t:.3f} ms -> {B/t*1e3:.0f} img/s")
    lat = ae.encoder(imgs).latent_space
    t = timeit(lambda: ae.decoder(lat)); print(f"   decoder alone (CUDA path; torch/cuDNN fp32 was 0.814 ms): {t:.3f} ms")
    profile(lambda: ae.decoder(lat))
    # config 4 with the reference's default decoder: DeiT + reverse-ResNet decoder (vitad_resnet_decoder_forward) + L2 map
    ae = get_model("ae_deit", 224, requires_grad=True)
    ae.decoder.load_state_dict({k[len("decoder."):]: v for k, v in W.make_resnet_decoder_state_dict(43).items()})
    ae = ae.cuda().eval()
    t = timeit(c4); print(f"config 4  DeiT + reverse-ResNet decoder + L2 map: {t:.3f} ms -> {B/t*1e3:.0f} img/s")
    t = timeit(lambda: ae.decoder(lat)); print(f"   reverse-ResNet decoder alone: {t:.3f} ms ({8.18e9*B/t/1e9:.1f} TFLOP/s of the reference's 8.18 GFLOP/img)")
    profile(lambda: ae.decoder(lat))
