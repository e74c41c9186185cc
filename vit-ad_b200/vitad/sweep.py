"""Config 5 of BASELINE.json: the 15-category MVTecAD-sized validation sweep (DeiT + GMM head and DeiT + NF head) over
1/2/4/8 GPUs — what `validation_loop.py` does category by category (validate_mdn :35-84, validate_nf :160-207), with the
reference's per-category tail (ValidatorMDN.py:170-183 result rows → ValidationHelper.calc_all_metrics :131-211).

Per category: every rank scores its batches (batch i → rank i % W, validators._BatchSharding) with the result rows left
on the device, the rows are all-gathered over NCCL (parallel.gather_results, asynchronous: NVLink moves category c while
the SMs score category c+1), and the rank that owns the category (c % W) computes its metrics on the device
(gpu_metrics) — so the AUROC work is spread over the ranks as well.  The only device→host traffic is the metric values.
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.distributed as dist

from .gpu_metrics import calc_all_metrics_device
from .parallel import gather_results
from .synthetic import batches


def run_sweep(v_gmm, v_nf, data: dict, rank: int = 0, world: int = 1, batch_size: int = 32, pixel_metrics: bool = True,
              fp_thres: float = 0.3, gmm_seed: int = 1234, lag: int | None = None) -> dict:
    """`data`: {category: (images [n,3,S,S], image_labels [n], pixel_labels [n,1,S,S])}, host (pinned) or device tensors.
    → {"metrics": {category/head: {...}} (complete on rank 0), "images": n, "timing": {...}}.  Enqueues everything on the
    current stream; returns after the last metric value has reached the host."""
    dev = v_gmm.device
    t0 = time.perf_counter()
    main = torch.cuda.current_stream(dev)
    # Metrics run on their own stream: their host-side reads (sorted-curve sizes, metric values) then wait only for the
    # gather of THEIR category, not for the scoring kernels of the next one already queued on the main stream — the SMs
    # never wait for the host.  Everything a metric touches stays referenced until the final synchronize.
    side = getattr(v_gmm, "_metrics_stream", None)
    if side is None:
        side = v_gmm._metrics_stream = torch.cuda.Stream(dev)
    pending, local_metrics, n_images = [], {}, 0
    keep_alive = []
    if lag is None:
        lag = len(data)  # enqueue every category's scoring and gathers first: ~0.9 GB of gathered rows stay alive per rank

    def finish(entry):
        ci, name, pend, ev = entry
        if ci % world != rank:  # not the owner of this category's metrics: only keep the buffers alive
            keep_alive.append(pend)
            return
        with torch.cuda.stream(side):
            side.wait_event(ev)
            for tag, p in pend.items():
                r = p.result()  # waits for the collectives on the metrics stream, stitches the global order
                keep_alive.append(r)
                if not pixel_metrics:
                    r = {k: v for k, v in r.items() if not k.startswith("pixel")}
                    r["pixel_labels"], r["pixel_scores"] = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
                m = calc_all_metrics_device(r, fp_thres=fp_thres, dataset_name=name, device=dev)
                local_metrics[f"{name}/{tag}"] = {k: v for k, v in m.items() if isinstance(v, float)}

    dealt = 0  # batches dealt so far: the round-robin continues across categories (61 batches over W ranks, not 15 x "rank 0 first")
    for ci, (name, (images, labels, masks)) in enumerate(data.items()):
        bl = batches(images, labels, masks, batch_size=batch_size)
        v_gmm.gumbel_seed = gmm_seed + ci  # noise field per category, keyed inside by the global batch index
        v_gmm.shard.offset = v_nf.shard.offset = dealt % world
        rg = v_gmm.valid_loop_transformer(bl, keep_origs=False, on_device=True)
        rn = v_nf.valid_loop_transformer_nf(bl, keep_origs=False, on_device=True)
        # every rank knows how the batches were dealt: no metadata exchange, no host/device synchronisation in the gather
        layout = {"batch_sizes": [int(b[0].shape[0]) for b in bl], "owners": [v_gmm.shard.owner(i) for i in range(len(bl))],
                  "map_shape": (1, int(images.shape[-2]), int(images.shape[-1])), "pixel_label_dtype": torch.uint8}
        dealt += len(bl)
        pend = {"gmm": gather_results(rg, len(bl), dev, async_op=True, layout=layout),
                "nf": gather_results(rn, len(bl), dev, async_op=True, layout=layout)}
        keep_alive.append((rg, rn))
        ev = torch.cuda.Event()
        ev.record(main)
        n_images += int(images.shape[0])
        pending.append((ci, name, pend, ev))
        # Metrics trail the scoring by `lag` categories: their host-side reads block this thread until the category's
        # gather and sort have run, and the main stream must hold enough queued scoring to cover that wait (at 8 GPUs a
        # category is ~3 ms of scoring per rank, a metric evaluation ~10 ms of latency).
        if len(pending) > lag:
            finish(pending.pop(0))
    while pending:
        finish(pending.pop(0))
    main.wait_stream(side)
    torch.cuda.synchronize(dev)
    t_local = time.perf_counter() - t0

    metrics = local_metrics
    if world > 1:  # metric values (a few floats per category) to rank 0
        parts = [None] * world
        dist.all_gather_object(parts, local_metrics)
        metrics = {}
        for p in parts:
            metrics.update(p)
    ordered = {}
    for name in data:
        for tag in ("gmm", "nf"):
            if f"{name}/{tag}" in metrics:
                ordered[f"{name}/{tag}"] = metrics[f"{name}/{tag}"]
    keep_alive.clear()
    return {"metrics": ordered, "images": n_images, "heads_per_image": 2, "host_s": t_local}


def build_sweep_models(rank: int, world: int, device, gaussians: int = 100):
    """The sweep's seeded synthetic-weight models (no network for checkpoints): DeiT-B + GMM(K) + NF(20 steps, 0.16)."""
    from . import synth_weights as W
    from .encoders import EncoderDeit
    from .mdn import GaussianMixtureDensityNetwork
    from .nf import NormalizingFlow
    from .validators import ValidatorMdn, ValidatorNF

    enc = EncoderDeit(224)
    enc.load_state_dict(W.make_deit_state_dict(seed=11, stress=True))
    head = GaussianMixtureDensityNetwork(768, 768, gaussians)
    np.random.seed(0)
    nf = NormalizingFlow(768, 224, 196, hidden_ratio=0.16, flow_steps=20)
    props = {"dataset": "synthetic_mvtec", "dataclass": "", "num_gaussians": gaussians, "fp_thres": 0.3}
    v_gmm = ValidatorMdn([head], enc, None, props, weights_object=[W.make_mdn_state_dict(21, gaussians, stress=True)],
                         rank=rank, world_size=world, gumbel_seed=1234)
    v_nf = ValidatorNF([nf], enc, None, props, weights_object=[W.make_nf_state_dict(31, stress=True)], rank=rank,
                       world_size=world)
    for m in (enc, head, nf):
        m.to(device).eval()
    return v_gmm, v_nf


def make_sweep_data(categories: int | None = None, pin: bool = True) -> dict:
    """Seeded MVTecAD-sized synthetic categories (synthetic.MVTEC_TEST_SIZES: 1725 images), pinned host memory as a
    DataLoader(pin_memory=True) would deliver them."""
    from .synthetic import MVTEC_TEST_SIZES, make_category

    cats = list(MVTEC_TEST_SIZES.items())[: categories or len(MVTEC_TEST_SIZES)]
    data = {}
    for i, (name, n) in enumerate(cats):
        images, labels, masks = make_category(name, n, seed=500 + i)
        data[name] = (images.pin_memory() if pin else images, labels, masks)
    return data
