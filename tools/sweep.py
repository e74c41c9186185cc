#!/usr/bin/env python
"""Config 5 of BASELINE.json: the full 15-category MVTecAD-sized synthetic validation sweep (DeiT + GMM head and
DeiT + NF head), batch 32, batches dealt round-robin over the ranks, NCCL all-gather of scores/maps/labels,
AUROC / PR-AUC on rank 0.

    python tools/sweep.py                                        # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sweep.py

Prints one JSON line (rank 0): images/s of the scoring path over the whole sweep (device time, max over ranks,
H2D of every batch included), the gather time and the per-category metrics."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from vitad import synth_weights as W  # noqa: E402  (seeded synthetic weights only)
from vitad.encoders import EncoderDeit  # noqa: E402
from vitad.mdn import GaussianMixtureDensityNetwork  # noqa: E402
from vitad.gpu_metrics import calc_all_metrics_device  # noqa: E402
from vitad.metrics import calc_all_metrics  # noqa: E402
from vitad.nf import NormalizingFlow  # noqa: E402
from vitad.parallel import gather_results, init_from_env  # noqa: E402
from vitad.synthetic import MVTEC_TEST_SIZES, batches, make_category  # noqa: E402
from vitad.validators import ValidatorMdn, ValidatorNF  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--categories", type=int, default=len(MVTEC_TEST_SIZES))
    ap.add_argument("--pixel-metrics", action="store_true", help="also pixel AUROC / PRO over all pixels")
    ap.add_argument("--sklearn", action="store_true", help="metrics through sklearn on the host instead of vitad.gpu_metrics")
    args = ap.parse_args()
    rank, world, local = init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    K = 100
    enc = EncoderDeit(224)
    enc.load_state_dict(W.make_deit_state_dict(seed=11, stress=True))
    head = GaussianMixtureDensityNetwork(768, 768, K)
    np.random.seed(0)
    nf = NormalizingFlow(768, 224, 196, hidden_ratio=0.16, flow_steps=20)
    props = {"dataset": "synthetic_mvtec", "dataclass": "", "num_gaussians": K, "fp_thres": 0.3}
    gseed = torch.Generator(device=dev).manual_seed(1234)
    gum = lambda bi, shape: -torch.empty(shape, device=dev).exponential_(generator=gseed).log()
    v_gmm = ValidatorMdn([head], enc, None, props, weights_object=[W.make_mdn_state_dict(21, K, stress=True)],
                         rank=rank, world_size=world, gumbel=gum)
    v_nf = ValidatorNF([nf], enc, None, props, weights_object=[W.make_nf_state_dict(31, stress=True)], rank=rank,
                       world_size=world)
    cats = list(MVTEC_TEST_SIZES.items())[: args.categories]
    data = {name: make_category(name, n, seed=500 + i) for i, (name, n) in enumerate(cats)}
    # pinned host memory, as a loader with pin_memory=True would deliver it
    data = {k: (t[0].pin_memory(),) + tuple(t[1:]) for k, t in data.items()}
    # warm-up (packs weights, builds workspaces)
    name0 = cats[0][0]
    wb = batches(*data[name0])  # one whole category: full and short batches, workspaces, pinned buffers
    v_gmm.shard.world_size, v_nf.shard.world_size = 1, 1
    v_gmm.shard.rank, v_nf.shard.rank = 0, 0
    v_gmm.valid_loop_transformer(wb), v_nf.valid_loop_transformer_nf(wb)
    v_gmm.shard.world_size, v_nf.shard.world_size = world, world
    v_gmm.shard.rank, v_nf.shard.rank = rank, rank

    results, n_images, t_gather = {}, 0, 0.0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    per_cat, loop_ms = {}, []
    for name, n in cats:
        bl = batches(*data[name])
        t0 = time.perf_counter()
        rg = v_gmm.valid_loop_transformer(bl, keep_origs=False)
        t1 = time.perf_counter()
        rn = v_nf.valid_loop_transformer_nf(bl, keep_origs=False)
        t2 = time.perf_counter()
        loop_ms.append((name, n, round((t1 - t0) * 1e3, 1), round((t2 - t1) * 1e3, 1)))
        per_cat[name] = (rg, rn, len(bl))
        n_images += n
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    t0 = time.perf_counter()
    for name, (rg, rn, nb) in per_cat.items():
        results[name] = (gather_results(rg, nb, dev), gather_results(rn, nb, dev))
    if world > 1:
        dist.barrier()
    t_gather = time.perf_counter() - t0
    if rank == 0:
        t_m0 = time.perf_counter()
        metrics = {}
        for name, (rg, rn) in results.items():
            for tag, r in (("gmm", rg), ("nf", rn)):
                if not args.pixel_metrics:
                    r = {k: v for k, v in r.items() if not k.startswith("pixel")}
                    r["pixel_labels"], r["pixel_scores"] = np.zeros(1), np.zeros(1)
                m = (calc_all_metrics if args.sklearn else calc_all_metrics_device)(r, fp_thres=0.3, dataset_name=name)
                metrics[f"{name}/{tag}"] = {k: round(v, 4) for k, v in m.items() if isinstance(v, float)}
        print(json.dumps({
            "workload": "15-category MVTecAD-sized synthetic validation sweep, DeiT + GMM(100) and DeiT + NF(20 steps), batch 32",
            "n_gpus": world, "images": n_images, "heads_per_image": 2,
            "ms_scoring": float(ms.item()), "images_per_s": 2 * n_images / (float(ms.item()) * 1e-3),
            "gather_s": t_gather, "metrics_s": time.perf_counter() - t_m0,
            "metrics_impl": "sklearn" if args.sklearn else "vitad.gpu_metrics", "loop_ms_gmm_nf": loop_ms, "metrics": metrics}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
