#!/usr/bin/env python
"""Config 5 of BASELINE.json from the command line: the full 15-category MVTecAD-sized synthetic validation sweep (DeiT +
GMM head and DeiT + NF head), batch 32, batches dealt round-robin over the ranks, NCCL all-gather of scores/maps/labels,
AUROC / PR-AUC on the device of the rank that owns each validation (vitad.sweep.run_sweep; `bench.py --workload sweep` times the same function).

    python tools/sweep.py                                        # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sweep.py

Prints one JSON line (rank 0) with every category's metrics: a sharded run must print the same metrics as N=1."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from vitad.parallel import init_from_env, warm_up  # noqa: E402
from vitad.sweep import build_sweep_models, make_sweep_data, run_sweep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--categories", type=int, default=0)
    ap.add_argument("--no-pixel-metrics", action="store_true")
    args = ap.parse_args()
    rank, world, local = init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    v_gmm, v_nf = build_sweep_models(rank, world, dev)
    data = make_sweep_data(args.categories or None)
    warm_up(dev)
    first = dict(list(data.items())[:1])
    run_sweep(v_gmm, v_nf, first, rank, world)  # packs weights, builds workspaces
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = run_sweep(v_gmm, v_nf, data, rank, world, pixel_metrics=not args.no_pixel_metrics)
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps({
            "workload": "15-category MVTecAD-sized synthetic validation sweep, DeiT + GMM(100) and DeiT + NF(20 steps), batch 32",
            "n_gpus": world, "images": out["images"], "heads_per_image": 2, "seconds_end_to_end": dt,
            "images_per_s": 2 * out["images"] / dt,
            "metrics": {k: {m: round(v, 4) for m, v in d.items()} for k, d in out["metrics"].items()}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
