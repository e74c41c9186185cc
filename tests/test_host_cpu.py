"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/vitad.h declares
(no compute without a GPU), the product path refuses CPU tensors, state_dict key layouts match the
reference's, the metrics wrapper, and the world_size-2 gloo gather."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from vitad import _lib

    header = open(os.path.join(ROOT, "include", "vitad.h")).read()
    names = sorted(set(re.findall(r"\b(vitad_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 15
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in include/vitad.h but not exported"
    assert _lib.lib.vitad_abi_version() == 1
    assert _lib.gmm_plan(100) == (1, 104, 100) and _lib.gmm_plan(110) == (1, 112, 110) and _lib.gmm_plan(130) == (2, 72, 65)
    with pytest.raises(_lib.VitadError):
        _lib.gmm_plan(500)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_entry_points_fail_loudly_without_gpu():
    import ctypes as C

    from vitad import _lib

    args = _lib.LinearArgs()
    rc = _lib.lib.vitad_linear_f16(C.byref(args), None)
    assert rc == -4 and b"CPU path" in _lib.lib.vitad_last_error()  # VITAD_ERR_ARCH
    # the entry points added for the reconstruction decoders and the input resize behave the same way
    buf = (C.c_uint8 * 64)()
    w = _lib.ResnetDecoderWeights()
    assert _lib.lib.vitad_resnet_decoder_forward(C.byref(w), buf, 1, buf, 64, buf, None) == -4
    assert _lib.lib.vitad_cnn_decoder_forward(C.byref(_lib.CnnDecoderWeights()), buf, 1, buf, 64, buf, None) == -4
    assert _lib.lib.vitad_resize_bilinear_u8(buf, 1, 2, 2, 2, buf, buf, buf, buf, None) == -4


def test_modules_refuse_cpu_tensors():
    from vitad.encoders import EncoderDeit
    from vitad.mdn import GaussianMixtureDensityNetwork

    with pytest.raises(RuntimeError, match="no CPU path"):
        EncoderDeit(224)(torch.rand(1, 3, 224, 224))
    with pytest.raises(RuntimeError, match="no CPU path"):
        GaussianMixtureDensityNetwork(768, 768, 100)(torch.rand(1, 196, 768))
    from vitad.autoencoders import DecoderResNetVariableEmbeddingSize, DecoderVanillaCNN

    with pytest.raises(RuntimeError, match="no CPU path"):
        DecoderResNetVariableEmbeddingSize(768)(torch.rand(1, 768))
    with pytest.raises(RuntimeError, match="no CPU path"):
        DecoderVanillaCNN(z_space=768, first_feature_map_size=7)(torch.rand(1, 768))


def test_state_dict_layouts_match_reference_keys():
    from oracle import weights as W
    from vitad.encoders import EncoderDeit
    from vitad.mdn import GaussianMixtureDensityNetwork

    enc = EncoderDeit(224)
    ref = W.make_deit_state_dict(seed=0)  # key layout proven against the reference in oracle/make_golden.py
    assert set(enc.state_dict().keys()) == set(ref.keys())
    for k, v in enc.state_dict().items():
        assert tuple(v.shape) == tuple(ref[k].shape), k
    enc.load_state_dict(ref, strict=True)
    assert enc.img_size == 224 and enc.patch_size == 16 and enc.size_patch_embedding == 768
    assert enc.num_embedded_patches == 196 and enc.architecture == "transformer_encoder"
    head = GaussianMixtureDensityNetwork(768, 768, 130)
    ref = W.make_mdn_state_dict(seed=0, num_gaussians=130)
    assert {k: tuple(v.shape) for k, v in head.state_dict().items()} == {k: tuple(v.shape) for k, v in ref.items()}
    assert torch.all(head.mu.bias == 0.001)


def test_metrics_match_sklearn_direct_calls():
    from sklearn import metrics as skm

    from vitad.metrics import calc_all_metrics

    rng = np.random.RandomState(0)
    n = 24
    labels = (rng.rand(n) > 0.5).astype(np.int64)
    scores = rng.rand(n) + 0.5 * labels
    pl = (rng.rand(n, 1, 16, 16) > 0.9).astype(np.float32)
    ps = rng.rand(n, 1, 16, 16).astype(np.float32) + 0.3 * pl
    out = calc_all_metrics({"image_scores": scores, "image_labels": labels, "pixel_scores": ps, "pixel_labels": pl},
                           fp_thres=0.3, dataset_name="t")
    assert out["image_auroc_score"] == pytest.approx(skm.roc_auc_score(labels, scores), abs=1e-12)
    assert out["pixel_auroc_score"] == pytest.approx(skm.roc_auc_score(pl.ravel(), ps.ravel()), abs=1e-12)
    assert 0.0 <= out["pro_score_0.3fp"] <= 1.0 and 0.0 <= out["image_prauc_score"] <= 1.0


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, os.path.join(sys.argv[1], "vit-ad_b200"))
import numpy as np, torch
from vitad.parallel import init_from_env, gather_results
from vitad.validators import _BatchSharding
rank, world, _ = init_from_env(backend="gloo")
sizes = [4, 4, 4, 4, 3]                       # five batches, short tail (no drop_last in the reference loader)
starts = np.cumsum([0] + sizes)
shard = _BatchSharding(rank, world)
mine = [b for b in range(len(sizes)) if shard.mine(b)]
ids = np.concatenate([np.arange(starts[b], starts[b + 1]) for b in mine])
res = {"image_scores": ids.astype(np.float32) * 0.5, "pixel_scores": np.tile(ids.astype(np.float32)[:, None, None, None], (1, 1, 4, 4)),
       "image_labels": (ids % 2).astype(np.int64), "pixel_labels": np.zeros((len(ids), 1, 4, 4), np.float32),
       "batch_index": np.asarray(mine), "batch_sizes": np.asarray([sizes[b] for b in mine])}
full = gather_results(res, num_batches=len(sizes), device=torch.device("cpu"))
n = sum(sizes)
assert np.array_equal(full["image_scores"], np.arange(n, dtype=np.float32) * 0.5), full["image_scores"]
assert np.array_equal(full["image_labels"], np.arange(n) % 2)
assert full["pixel_scores"].shape == (n, 1, 4, 4) and np.array_equal(full["pixel_scores"][:, 0, 0, 0], np.arange(n, dtype=np.float32))
# the same rows as torch tensors (valid_loop_*(on_device=True)): gathered without a numpy round trip, tensors come back
rest = {k: (torch.from_numpy(v) if k not in ("batch_index", "batch_sizes") else v) for k, v in res.items()}
rest["pixel_labels"] = rest["pixel_labels"].to(torch.uint8)
pend = gather_results(rest, num_batches=len(sizes), device=torch.device("cpu"), async_op=True)
full_t = pend.result()
assert torch.is_tensor(full_t["image_scores"]) and full_t["pixel_labels"].dtype == torch.uint8
assert np.array_equal(full_t["image_scores"].numpy(), full["image_scores"]) and np.array_equal(full_t["pixel_scores"].numpy(), full["pixel_scores"])
# the layout path: every rank derives the counts from the dealing rule, nothing is exchanged but the payloads; offset 1
shard1 = _BatchSharding(rank, world, offset=1)
mine1 = [b for b in range(len(sizes)) if shard1.mine(b)]
ids1 = np.concatenate([np.arange(starts[b], starts[b + 1]) for b in mine1]) if mine1 else np.zeros(0, np.int64)
res1 = {"image_scores": torch.from_numpy(ids1.astype(np.float32) * 0.5), "pixel_scores": torch.from_numpy(np.tile(ids1.astype(np.float32)[:, None, None, None], (1, 1, 4, 4))),
        "image_labels": torch.from_numpy((ids1 % 2).astype(np.int64)), "pixel_labels": torch.zeros((len(ids1), 1, 4, 4), dtype=torch.uint8),
        "batch_index": np.asarray(mine1, dtype=np.int64), "batch_sizes": np.asarray([sizes[b] for b in mine1], dtype=np.int64)}
layout = {"batch_sizes": sizes, "owners": [shard1.owner(b) for b in range(len(sizes))], "map_shape": (1, 4, 4), "pixel_label_dtype": torch.uint8}
full_l = gather_results(res1, num_batches=len(sizes), device=torch.device("cpu"), async_op=True, layout=layout).result()
assert np.array_equal(full_l["image_scores"].numpy(), full["image_scores"]) and np.array_equal(full_l["pixel_scores"].numpy(), full["pixel_scores"])
assert np.array_equal(full_l["image_labels"].numpy(), full["image_labels"]) and full_l["pixel_labels"].dtype == torch.uint8
# the routed exchange of the sweep: three validations with their own layouts, each wanted in full by one owner rank only
from vitad.parallel import exchange_to_owners
entries, truth = [], []
for p, (szs, off) in enumerate((([4, 4, 3], 0), ([2], 1), ([4, 4, 4, 4, 1], 2))):
    st = np.cumsum([0] + szs)
    sh = _BatchSharding(rank, world, offset=off)
    mine_p = [b for b in range(len(szs)) if sh.mine(b)]
    ids_p = (np.concatenate([np.arange(st[b], st[b + 1]) for b in mine_p]) if mine_p else np.zeros(0, np.int64)) + 1000 * p
    resp = {"image_scores": torch.from_numpy(ids_p.astype(np.float32)), "pixel_scores": torch.from_numpy(np.tile(ids_p.astype(np.float32)[:, None, None, None], (1, 1, 4, 4))),
            "image_labels": torch.from_numpy((ids_p % 2).astype(np.int64)), "pixel_labels": torch.from_numpy((ids_p % 3 == 0).astype(np.uint8)[:, None, None, None].repeat(4, 2).repeat(4, 3))}
    if not mine_p:
        resp = {}
    entries.append({"result": resp, "layout": {"batch_sizes": szs, "owners": [sh.owner(b) for b in range(len(szs))], "map_shape": (1, 4, 4), "pixel_label_dtype": torch.uint8}})
    truth.append(np.arange(sum(szs)) + 1000 * p)
pair_owner = [p % world for p in range(3)]
got = exchange_to_owners(entries, pair_owner, torch.device("cpu"))
assert sorted(got) == [p for p in range(3) if pair_owner[p] == rank], (rank, sorted(got))
for p, r in got.items():
    assert np.array_equal(r["image_scores"].numpy(), truth[p].astype(np.float32)), (p, r["image_scores"])
    assert np.array_equal(r["pixel_scores"][:, 0, 0, 0].numpy(), truth[p].astype(np.float32)) and r["pixel_scores"].shape[1:] == (1, 4, 4)
    assert np.array_equal(r["image_labels"].numpy(), truth[p] % 2) and np.array_equal(r["pixel_labels"][:, 0, 0, 0].numpy(), (truth[p] % 3 == 0).astype(np.uint8))
# fewer batches than ranks: the rank without a batch must take part in the collectives with empty payloads (no hang, no raise)
one = [b for b in range(1) if _BatchSharding(rank, world).mine(b)]
r1 = {"image_scores": np.arange(3, dtype=np.float32), "pixel_scores": np.ones((3, 1, 4, 4), np.float32), "image_labels": np.array([0, 1, 0]),
      "pixel_labels": np.zeros((3, 1, 4, 4), np.float32), "batch_index": np.asarray([0]), "batch_sizes": np.asarray([3])} if one else \
     {"batch_index": np.zeros(0, np.int64), "batch_sizes": np.zeros(0, np.int64)}
f1 = gather_results(r1, num_batches=1, device=torch.device("cpu"))
assert np.array_equal(f1["image_scores"], np.arange(3, dtype=np.float32)) and f1["pixel_scores"].shape == (3, 1, 4, 4)
assert np.array_equal(f1["image_labels"], [0, 1, 0])
torch.distributed.destroy_process_group()
print("rank", rank, "ok")
"""


@pytest.mark.parametrize("world", [2, 3])
def test_batch_sharded_gather_gloo(tmp_path, world):
    """world 2: the N>1 path; world 3: also ranks that hold nothing (one batch on three ranks)."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29531 + world), str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == world


@pytest.mark.parametrize("seed,n,ties", [(0, 500, False), (1, 4000, True), (2, 37, True), (3, 20000, True)])
def test_device_metrics_equal_sklearn(seed, n, ties):
    """vitad.gpu_metrics (sort + cumulative counts, torch ops — run on the CPU here) against the sklearn calls of
    ValidationHelper.calc_all_metrics: AUROC, PR-AUC, the FPR-limited threshold and the thresholded 'PRO' AUROC,
    with heavy ties (quantised scores, as a thresholded anomaly map has)."""
    import torch
    from sklearn import metrics as skm

    from vitad import gpu_metrics as G
    from vitad.metrics import calc_threshold as sk_threshold

    rng = np.random.default_rng(seed)
    y = (rng.random(n) < 0.3).astype(np.float32)
    s = (rng.normal(size=n) + 0.8 * y).astype(np.float32)
    if ties:
        s = np.round(s * 4) / 4
    ts, ty = torch.from_numpy(s), torch.from_numpy(y)
    assert abs(G.roc_auc_score(ts, ty) - skm.roc_auc_score(y, s)) < 1e-9
    p, r, _ = skm.precision_recall_curve(y, s)
    assert abs(G.pr_auc_score(ts, ty) - skm.auc(r, p)) < 1e-9
    for fpr_t in (0.05, 0.3, 0.9):
        thr = sk_threshold(s, y, fpr_t)
        assert G.calc_threshold(ts, ty, fpr_t) == pytest.approx(thr, abs=0), (fpr_t, thr)
        an = np.where(s > thr, s, 0)
        assert abs(G.roc_auc_score(torch.from_numpy(an), ty) - skm.roc_auc_score(y, an)) < 1e-9


@pytest.mark.parametrize("seed,n,levels", [(0, 3000, 6), (1, 3000, 11), (2, 50000, 24), (3, 200, 3), (4, 4000, 64)])
def test_device_threshold_and_pro_with_collinear_ties(seed, n, levels):
    """Tie-heavy NON-NEGATIVE scores (few distinct levels, as anomaly maps quantise) whose per-level positive share is
    constant over runs of levels: roc_curve(drop_intermediate=True) then drops collinear points together with their
    thresholds, and ValidationHelper.calc_threshold (:70-88) must be reproduced on the kept points only.  Also pins
    the one-sort derivation of the thresholded-map AUROC (predict_anomaly 'fluently', :91-104) against sklearn."""
    import torch
    from sklearn import metrics as skm

    from vitad import gpu_metrics as G
    from vitad.metrics import calc_threshold as sk_threshold

    rng = np.random.default_rng(seed)
    # upper half of the levels: the SAME (positives, negatives) count per level -> equal ROC increments -> diff(fps, 2) ==
    # diff(tps, 2) == 0 at the interior levels of the run; lower half: random counts
    per = max(4, n // (2 * levels))
    lev_list, y_list = [], []
    for v in range(levels):
        if v >= levels // 2:
            n_pos, n_neg = per // 4, per - per // 4
        else:
            n_pos, n_neg = int(rng.integers(0, per // 8 + 1)), int(rng.integers(1, per))
        lev_list += [v] * (n_pos + n_neg)
        y_list += [1.0] * n_pos + [0.0] * n_neg
    perm = rng.permutation(len(lev_list))
    lev, y = np.asarray(lev_list)[perm], np.asarray(y_list, np.float32)[perm]
    s = (lev / levels).astype(np.float32)
    ts, ty = torch.from_numpy(s), torch.from_numpy(y)
    fpr, tpr, thr_all = skm.roc_curve(y, s)
    fpr_full, _, _ = skm.roc_curve(y, s, drop_intermediate=False)
    assert len(fpr) < len(fpr_full) or levels <= 3, "the case is meant to make drop_intermediate drop points"
    c = G.Curve(ts, ty)
    for fpr_t in (0.02, 0.1, 0.3, 0.5, 0.75, 1.0):
        thr = sk_threshold(s, y, fpr_t)
        assert G.calc_threshold(ts, ty, fpr_t, c) == pytest.approx(thr, abs=0), (fpr_t, thr)
        an = np.where(s > thr, s, 0)
        ref = skm.roc_auc_score(y, an)
        assert abs(G.thresholded_roc_auc(c, thr) - ref) < 1e-9, (fpr_t, thr)
        assert abs(G.roc_auc_score(torch.from_numpy(an), ty) - ref) < 1e-9
    out = G.calc_all_metrics_device({"image_scores": s[:64], "image_labels": np.r_[y[:62], 0, 1], "pixel_scores": s,
                                     "pixel_labels": y}, fp_thres=0.3, device=torch.device("cpu"))
    from vitad.metrics import calc_all_metrics

    ref = calc_all_metrics({"image_scores": s[:64], "image_labels": np.r_[y[:62], 0, 1], "pixel_scores": s,
                            "pixel_labels": y}, fp_thres=0.3)
    for k, v in ref.items():
        if isinstance(v, float):
            assert abs(out[k] - v) < 1e-9, (k, out[k], v)


def test_esvit_checkpoint_position_encoding_interpolation():
    """EncoderEsVit.load_student_weights = the reference's checkpoint path
    (TransformerEncoder.py:248-263, interpolate_position_encoding :276-350).  A window-7 'student' checkpoint is
    loaded into the window-14 model; the resized tables / indices equal the reference function's output
    (fixture from oracle/make_golden.py case esvit_interpolate)."""
    from oracle import weights as W
    from vitad.encoders import EncoderEsVit, interpolate_position_encoding

    g = np.load(os.path.join(ROOT, "tests", "golden", "esvit_interpolate.npz"))
    enc = EncoderEsVit(224, requires_grad=True)
    delattr(enc.esvit, "head")
    ckpt = W.synthetic_esvit_checkpoint()
    out = interpolate_position_encoding(weights=ckpt, model=enc.esvit)
    assert len(g.files) >= 6
    for key in g.files:
        k = key.replace("__", ".")
        np.testing.assert_allclose(out[k].float().numpy(), g[key], rtol=0, atol=1e-6, err_msg=k)
    enc.load_student_weights(ckpt)  # strict load succeeds with the adapted shapes
    sd = enc.esvit.state_dict()
    assert sd["layers.0.blocks.0.attn.relative_position_bias_table"].shape == (729, 3)
    assert sd["layers.0.blocks.0.attn.relative_position_index"].dtype == torch.int64


def test_delivery_plan_tiles_the_owner_buffer_in_global_batch_order():
    """vitad.sweep.delivery_plan (peer-memory transport of config 5): the pieces all ranks write cover the validation's rows
    exactly once, in global batch order, whatever the dealing offset, ragged last batches and ranks without a batch."""
    from vitad.sweep import delivery_plan
    from vitad.validators import _BatchSharding

    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        for trial in range(20):
            nb = int(rng.integers(1, 12))
            sizes = [32] * (nb - 1) + [int(rng.integers(1, 33))]
            offset = int(rng.integers(0, world))
            owners = [_BatchSharding(0, world, offset).owner(b) for b in range(nb)]
            base = int(rng.integers(0, 100))
            total = sum(sizes)
            buf = np.full(base + total + 5, -1, dtype=np.int64)
            starts = np.concatenate(([0], np.cumsum(sizes)))
            for rank in range(world):
                local = np.concatenate([np.arange(starts[b], starts[b + 1]) for b in range(nb) if owners[b] == rank] or [np.zeros(0, np.int64)])
                covered = 0
                for row0, lo, hi in delivery_plan(sizes, owners, rank, base):
                    assert (buf[row0:row0 + hi - lo] == -1).all()  # nobody else wrote here
                    buf[row0:row0 + hi - lo] = local[lo:hi]
                    covered += hi - lo
                assert covered == local.size
            assert (buf[base:base + total] == np.arange(total)).all() and (buf[:base] == -1).all() and (buf[base + total:] == -1).all()
