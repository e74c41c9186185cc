"""north_star: image-level AUROC identical to 4 decimals between the CUDA path and the reference algorithm
(oracle) on a fixed synthetic anomaly set; scores/maps within 1e-3 per element (helpers.assert_rel).  One batch of 24
images (the GMM score couples a batch through its global max, so the whole set is scored as one batch on both sides)."""
import numpy as np
import pytest
import torch

from helpers import gumbel

pytestmark = pytest.mark.gpu


def test_image_auroc_identical_to_4_decimals_on_synthetic_anomaly_set():
    """The whole designed set (helpers.DESIGNED_GMM, 24 images, oracle scores pairwise >= 10x the allowed noise apart —
    asserted): per-element 1e-3 parity of scores and maps, identical score order, image AUROC identical to 4 decimals."""
    from sklearn.metrics import roc_auc_score

    from helpers import DESIGNED_GMM, GMM_NOISE_SEED_BASE, assert_designed_separation, assert_map_parity, assert_rel
    from oracle import vitad_oracle as O
    from oracle import weights as W
    from vitad.encoders import EncoderDeit
    from vitad.mdn import GaussianMixtureDensityNetwork
    from vitad.synthetic import batches, make_designed_set
    from vitad.validators import ValidatorMdn

    K = 100
    n = len(DESIGNED_GMM)
    images, labels, masks = make_designed_set(DESIGNED_GMM)
    enc_sd = W.make_deit_state_dict(seed=11, stress=True)
    mdn_sd = W.make_mdn_state_dict(seed=21, num_gaussians=K, stress=True)
    g = torch.stack([O.gumbel_noise((196, K), torch.Generator().manual_seed(GMM_NOISE_SEED_BASE + s)) for s in DESIGNED_GMM])

    with torch.no_grad():
        tok, _ = O.deit_forward(enc_sd, images)
        ref_scores, ref_maps = O.mdn_scores(O.mdn_probability_map(O.mdn_patch_loglik(tok, mdn_sd, g)), 224, 16)
    ref_scores, ref_maps = ref_scores.numpy(), ref_maps.numpy()
    assert_designed_separation(ref_scores, labels.numpy(), factor=10.0)

    enc = EncoderDeit(224)
    enc.load_state_dict(enc_sd)
    head = GaussianMixtureDensityNetwork(768, 768, K)
    props = {"dataset": "synthetic", "dataclass": "designed", "num_gaussians": K, "fp_thres": 0.3}
    val = ValidatorMdn([head], enc, None, props, weights_object=[mdn_sd], gumbel=lambda bi, shape: g)
    res = val.valid_loop_transformer(batches(images, labels, masks, batch_size=n))

    assert_rel(res["image_scores"], ref_scores, 1e-3, what="image scores")
    assert_map_parity(res["pixel_scores"], ref_maps, what="anomaly maps")
    lab = labels.numpy()
    auroc_ref = roc_auc_score(lab, ref_scores)
    auroc = roc_auc_score(res["image_labels"], res["image_scores"])
    assert 0.05 < auroc_ref < 0.95, auroc_ref
    assert round(auroc, 4) == round(auroc_ref, 4), (auroc, auroc_ref)
    assert np.array_equal(np.argsort(ref_scores), np.argsort(res["image_scores"]))
    pix_ref = roc_auc_score(masks.numpy().ravel() > 0.5, ref_maps.ravel())
    pix = roc_auc_score(res["pixel_labels"].ravel() > 0.5, res["pixel_scores"].ravel())
    assert abs(pix - pix_ref) <= 5e-4, (pix, pix_ref)


def test_device_metrics_equal_sklearn_on_validator_output():
    """calc_all_metrics on the GPU (sort + cumulative counts) equals the reference's sklearn calls
    (ValidationHelper.py:131-211) to 4 decimals on pixel-level data with ties (1.2 M pixels, thresholded maps)."""
    from vitad.gpu_metrics import calc_all_metrics_device
    from vitad.metrics import calc_all_metrics
    from vitad.synthetic import make_category

    n = 24
    _images, labels, masks = make_category("cable", n, seed=78)
    rng = np.random.default_rng(5)
    maps = rng.random((n, 1, 224, 224), dtype=np.float32) * 0.5 + 0.5 * masks.numpy() * rng.random((n, 1, 224, 224), dtype=np.float32)
    maps = np.round(maps * 512) / 512  # ties
    scores = maps.reshape(n, -1).max(1)
    result = {"image_scores": scores, "image_labels": labels.numpy(), "pixel_scores": maps, "pixel_labels": masks.numpy()}
    ref = calc_all_metrics(result, fp_thres=0.3, dataset_name="x")
    got = calc_all_metrics_device(result, fp_thres=0.3, dataset_name="x")
    for k, v in ref.items():
        if isinstance(v, float):
            assert round(got[k], 4) == round(v, 4), (k, got[k], v)
    assert set(got) == set(ref)


def test_sweep_metrics_equal_the_per_validator_sklearn_path():
    """vitad.sweep.run_sweep (device-resident rows, metrics on a side stream, one sort per validation) against the
    reference-shaped path: valid_loop_* → numpy result dictionary → sklearn (vitad.metrics = ValidationHelper.py:131-211),
    category by category, on host-resident and on device-resident data."""
    from vitad.metrics import calc_all_metrics
    from vitad.sweep import build_sweep_models, run_sweep
    from vitad.synthetic import batches, make_category

    dev = torch.device("cuda", torch.cuda.current_device())
    v_gmm, v_nf = build_sweep_models(0, 1, dev)
    data = {"alpha": make_category("alpha", 37, seed=901), "beta": make_category("beta", 70, seed=902)}
    data = {k: (t[0].pin_memory(), t[1], (t[2] != 0).to(torch.uint8)) for k, t in data.items()}
    out = run_sweep(v_gmm, v_nf, data, batch_size=32)
    resident = {k: (t[0].to(dev), t[1], t[2].to(dev)) for k, t in data.items()}
    out_dev = run_sweep(v_gmm, v_nf, resident, batch_size=32)
    assert out["images"] == 107 and out["metrics"] == out_dev["metrics"]
    for ci, (name, (images, labels, masks)) in enumerate(data.items()):
        bl = batches(images, labels, masks.float(), batch_size=32)
        v_gmm.gumbel_seed = 1234 + ci
        v_gmm.shard.offset = v_nf.shard.offset = 0
        for tag, res in (("gmm", v_gmm.valid_loop_transformer(bl, keep_origs=False)),
                         ("nf", v_nf.valid_loop_transformer_nf(bl, keep_origs=False))):
            ref = calc_all_metrics(res, fp_thres=0.3, dataset_name=name)
            got = out["metrics"][f"{name}/{tag}"]
            for k, v in ref.items():
                if isinstance(v, float) and k != "fp_thres":
                    assert abs(got[k] - v) < 1e-6, (name, tag, k, got[k], v)


_PEER_SWEEP_SCRIPT = r"""
import os, sys, json
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
dist.init_process_group("nccl", init_method="file://" + sys.argv[2], rank=0, world_size=1, device_id=dev)
from vitad.parallel import PeerMailbox
from vitad.sweep import build_sweep_models, run_sweep
from vitad.synthetic import make_category

# the mailbox alone: rows written in two pieces, one signal, the wait on a side stream
box = PeerMailbox(40, (1, 8, 8), dev)
g = torch.Generator().manual_seed(3)
res = {"image_scores": torch.rand(24, generator=g).to(dev), "pixel_scores": torch.rand(24, 1, 8, 8, generator=g).to(dev),
       "image_labels": torch.randint(0, 2, (24,), generator=g).to(dev), "pixel_labels": torch.randint(0, 2, (24, 1, 8, 8), generator=g).to(torch.uint8).to(dev)}
box.put(0, 5, res, 0, 10)
box.put(0, 15, res, 10, 24)
box.signal(0, 7)
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    box.wait_all(7)
    got = {k: v.clone() for k, v in box.rows(5, 24).items()}
side.synchronize()
mailbox_ok = all(torch.equal(got[k], res[k]) for k in res)
try:
    box.put(0, 30, res, 0, 24)
    bounds_ok = False
except ValueError:
    bounds_ok = True

# the sweep through each delivery path of the multi-rank code (one rank) against the one-GPU path
v_gmm, v_nf = build_sweep_models(0, 1, dev)
data = {n: make_category(n, k, seed=910 + i) for i, (n, k) in enumerate((("alpha", 37), ("beta", 70), ("gamma", 33)))}
data = {k: (t[0].to(dev), t[1], (t[2] != 0).to(torch.uint8).to(dev)) for k, t in data.items()}
one = run_sweep(v_gmm, v_nf, data, batch_size=32)
peer = run_sweep(v_gmm, v_nf, data, batch_size=32, transport="peer")
peer2 = run_sweep(v_gmm, v_nf, data, batch_size=32, transport="peer")  # buffers and signals are reused
exch = run_sweep(v_gmm, v_nf, data, batch_size=32, transport="exchange")
print(json.dumps({"mailbox_ok": mailbox_ok, "bounds_ok": bounds_ok, "transports": [one["transport"], peer["transport"], exch["transport"]],
                  "peer_equal": peer["metrics"] == one["metrics"], "peer_again_equal": peer2["metrics"] == one["metrics"],
                  "exchange_equal": exch["metrics"] == one["metrics"], "n": len(one["metrics"])}))
dist.destroy_process_group()
"""


def test_sweep_peer_memory_delivery_reproduces_the_one_gpu_metrics(tmp_path):
    """config 5's multi-rank code path on the one GPU of the test box (a single-rank NCCL group in a child process): result
    rows delivered through symmetric memory (parallel.PeerMailbox: copy kernels + stream-ordered signals) and through the
    routed NCCL exchange give exactly the metrics of the one-GPU path; the mailbox returns the rows it was sent."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "peer_sweep.py"
    script.write_text(_PEER_SWEEP_SCRIPT)
    env = dict(os.environ, VITAD_SWEEP_TRANSPORT="")
    env.pop("VITAD_SWEEP_TRANSPORT")
    r = subprocess.run([sys.executable, str(script), os.path.join(root, "vit-ad_b200"), str(tmp_path / "store")],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["mailbox_ok"] and out["bounds_ok"], out
    assert out["transports"] == ["one GPU", "peer", "exchange"], out
    assert out["peer_equal"] and out["peer_again_equal"] and out["exchange_equal"] and out["n"] == 6, out
