"""Diagnostic (torchrun, N GPUs): gathered sharded validator results (GMM and NF heads, device-resident rows) must equal
the unsharded results bit for bit.  The one-GPU form of the same check is tests/test_gmm_gpu.py::
test_sharded_validators_reproduce_the_unsharded_results_bit_for_bit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200"))
import numpy as np, torch, torch.distributed as dist
from vitad import synth_weights as W
from vitad.encoders import EncoderDeit
from vitad.mdn import GaussianMixtureDensityNetwork
from vitad.nf import NormalizingFlow
from vitad.parallel import gather_results, init_from_env
from vitad.synthetic import batches, make_category
from vitad.validators import ValidatorMdn, ValidatorNF
rank, world, local = init_from_env(); torch.cuda.set_device(local); dev = torch.device("cuda", local)
enc = EncoderDeit(224); enc.load_state_dict(W.make_deit_state_dict(11, stress=True))
np.random.seed(0); nf = NormalizingFlow(768, 224, 196, 0.16, 20)
props = {"dataset": "s", "dataclass": "x", "fp_thres": 0.3, "num_gaussians": 100}
nf_sd, mdn_sd = W.make_nf_state_dict(31, stress=True), W.make_mdn_state_dict(21, 100, stress=True)
bl = batches(*make_category("cable", 150, seed=501))
def run(kind, r, w):
    if kind == "nf":
        return ValidatorNF([nf], enc, None, props, weights_object=[nf_sd], rank=r, world_size=w).valid_loop_transformer_nf(
            bl, keep_origs=False, on_device=True)
    v = ValidatorMdn([GaussianMixtureDensityNetwork(768, 768, 100)], enc, None, props, weights_object=[mdn_sd], rank=r,
                     world_size=w, gumbel_seed=77)
    return v.valid_loop_transformer(bl, keep_origs=False, on_device=True)
for kind in ("nf", "gmm"):
    full, part = run(kind, 0, 1), run(kind, rank, world)
    g = gather_results(part, len(bl), dev)
    ok = {k: bool(torch.equal(g[k], full[k])) for k in ("image_scores", "pixel_scores", "image_labels", "pixel_labels")}
    print(rank, kind, ok, "n", g["image_scores"].shape[0], full["image_scores"].shape[0])
    assert all(ok.values()), (kind, ok)
dist.destroy_process_group()
