"""Diagnostic: in-situ per-launch-site timing of one bs-32 step using the library's event profiler."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200")); sys.path.insert(0, ROOT)
import torch
from vitad import synth_weights as W
from vitad import _lib, ops
from vitad.encoders import EncoderDeit
from vitad.mdn import GaussianMixtureDensityNetwork

lib = _lib.lib
lib.vitad_profile_enable.argtypes = [C.c_int]; lib.vitad_profile_report.argtypes = [C.c_char_p, C.c_int]; lib.vitad_profile_report.restype = C.c_int
B, K = (int(sys.argv[2]) if len(sys.argv) > 2 else 32), (int(sys.argv[1]) if len(sys.argv) > 1 else 100)
enc = EncoderDeit(224); enc.load_state_dict(W.make_deit_state_dict(11)); enc = enc.cuda().eval()
head = GaussianMixtureDensityNetwork(768, 768, K); head.load_state_dict(W.make_mdn_state_dict(21, K)); head = head.cuda().eval()
imgs = W.synthetic_images(1, B).cuda(); gn = torch.randn(B, 196, K, device="cuda")
def step():
    f = enc(imgs); prob, sc = head.score(f.patch_embedding, gn)
    ops.bilinear_up(prob.view(-1, 14, 14), 224, True, post_one_minus=True)
with torch.no_grad():
    for _ in range(3): step()
    torch.cuda.synchronize()
    lib.vitad_profile_enable(1)
    N = 5
    for _ in range(N): step()
    buf = C.create_string_buffer(1 << 16)
    lib.vitad_profile_report(buf, len(buf))
    lib.vitad_profile_enable(0)
rows = [l.split() for l in buf.value.decode().strip().split("\n")]
tot = sum(float(r[2]) for r in rows if r[0] != "__span__")
span = [float(r[2]) for r in rows if r[0] == "__span__"][0]
print(f"per step: sum of launch-site times {tot/N:.1f} us, span {span/N:.1f} us (includes event overhead)")
for r in sorted(rows, key=lambda r: -float(r[2])):
    if r[0] != "__span__":
        print(f"  {r[0]:36s} n/step {int(r[1])/N:5.1f}  avg {float(r[2])/int(r[1]):8.1f} us  per-step {float(r[2])/N:8.1f} us")
