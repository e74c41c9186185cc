"""Diagnostic (torchrun, N GPUs): gathered sharded validator results must equal the unsharded results."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200"))
import numpy as np, torch, torch.distributed as dist
from vitad import synth_weights as W
from vitad.encoders import EncoderDeit
from vitad.nf import NormalizingFlow
from vitad.parallel import gather_results, init_from_env
from vitad.synthetic import batches, make_category
from vitad.validators import ValidatorNF
rank, world, local = init_from_env(); torch.cuda.set_device(local); dev = torch.device("cuda", local)
enc = EncoderDeit(224); enc.load_state_dict(W.make_deit_state_dict(11, stress=True))
np.random.seed(0); nf = NormalizingFlow(768, 224, 196, 0.16, 20)
props = {"dataset": "s", "dataclass": "x", "fp_thres": 0.3}
sd = W.make_nf_state_dict(31, stress=True)
full = ValidatorNF([nf], enc, None, props, weights_object=[sd])
part = ValidatorNF([nf], enc, None, props, weights_object=[sd], rank=rank, world_size=world)
bl = batches(*make_category("cable", 150, seed=501))
rf = full.valid_loop_transformer_nf(bl)
rp = part.valid_loop_transformer_nf(bl); rp.pop("origs")
g = gather_results(rp, len(bl), dev)
print(rank, "scores equal", np.array_equal(g["image_scores"], rf["image_scores"]), "labels equal", np.array_equal(g["image_labels"], rf["image_labels"]),
      "maps equal", np.array_equal(g["pixel_scores"], rf["pixel_scores"]), "max diff", np.abs(g["image_scores"] - rf["image_scores"]).max(),
      "n", len(g["image_scores"]), len(rf["image_scores"]))
if rank == 0:
    bad = np.nonzero(g["image_scores"] != rf["image_scores"])[0]
    print("mismatch idx", bad[:20])
dist.destroy_process_group()
