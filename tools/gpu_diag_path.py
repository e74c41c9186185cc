"""Diagnostic (not a test): timings of the encoder and GMM head at bs 32 on the GPU box."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200")); sys.path.insert(0, ROOT)
import torch
from vitad import synth_weights as W
from vitad import _lib, ops
from vitad.encoders import EncoderDeit
from vitad.mdn import GaussianMixtureDensityNetwork

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K = int(sys.argv[2]) if len(sys.argv) > 2 else 100
enc = EncoderDeit(224); enc.load_state_dict(W.make_deit_state_dict(11, stress=True)); enc = enc.cuda().eval()
head = GaussianMixtureDensityNetwork(768, 768, K); head.load_state_dict(W.make_mdn_state_dict(21, K)); head = head.cuda().eval()
imgs = W.synthetic_images(1, B).cuda()
gn = torch.randn(B, 196, K, device="cuda")
with torch.no_grad():
    f = enc(imgs); torch.cuda.synchronize()
    t_enc = timeit(lambda: enc(imgs))
    print(f"deit forward B={B}: {t_enc:.3f} ms  -> {B/t_enc*1e3:.0f} img/s  ({35.31e9*B/t_enc/1e9:.1f} TFLOP/s)")
    x = f.patch_embedding
    t_head = timeit(lambda: head.score(x, gn))
    flops = 2*B*196*768*2*K*768
    print(f"gmm head K={K} B={B}: {t_head:.3f} ms ({flops/t_head/1e9:.1f} TFLOP/s)")
    def full():
        f = enc(imgs); prob, sc = head.score(f.patch_embedding, gn)
        ops.bilinear_up(prob.view(-1, 14, 14), 224, True, post_one_minus=True)
    t_full = timeit(full)
    print(f"full path B={B}: {t_full:.3f} ms -> {B/t_full*1e3:.0f} img/s; launches/iter ~ {(_lib.launch_count())}")
    # kernel-level breakdown with events around pieces of the head
    M = B*196
    xf = x.reshape(M, 768); xaug = x._vitad_xaug[0]
    n_kc, kc, _ = _lib.gmm_plan(K)
    lp2 = torch.empty(M, n_kc*kc, device="cuda"); g2 = gn.reshape(M, K).contiguous()
    pk = head._packed
    s = torch.cuda.current_stream().cuda_stream
    t = timeit(lambda: _lib.check(_lib.lib.vitad_gmm_log_pi(xf.data_ptr(), 768, pk["pi_w"].data_ptr(), pk["pi_b"].data_ptr(), g2.data_ptr(), lp2.data_ptr(), M, 768, K, s)))
    print(f"  log_pi kernel: {t*1e3:.1f} us")
    ll = torch.empty(768, M, device="cuda"); L = torch.empty(M, device="cuda")
    t = timeit(lambda: _lib.check(_lib.lib.vitad_gmm_patch_loglik(xaug.data_ptr(), pk["w"].data_ptr(), lp2.data_ptr(), xf.data_ptr(), 768, ll.data_ptr(), M, L.data_ptr(), M, 768, K, s)))
    print(f"  fused projection+logsumexp (+mean): {t*1e3:.1f} us ({flops/t/1e9:.1f} TFLOP/s)")
