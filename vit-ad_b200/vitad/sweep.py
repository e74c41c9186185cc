"""Config 5 of BASELINE.json: the 15-category MVTecAD-sized validation sweep (DeiT + GMM head and DeiT + NF head) over
1/2/4/8 GPUs — what `validation_loop.py` does category by category (validate_mdn :35-84, validate_nf :160-207), with the
reference's per-category tail (ValidatorMDN.py:170-183 result rows → ValidationHelper.calc_all_metrics :131-211).

Every rank scores its batches (dealt round-robin over the whole sweep, validators._BatchSharding) with the result rows left
on the device; one routed NCCL exchange (parallel.exchange_to_owners) then sends each (category, head) validation to the
rank that evaluates its metrics on the device (gpu_metrics) — so the AUROC sorts are spread over the ranks as well.  The
only device→host traffic is the metric values.
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.distributed as dist

from .gpu_metrics import calc_all_metrics_device
from .parallel import exchange_to_owners
from .synthetic import batches


METRIC_KEYS = ("image_auroc_score", "image_prauc_score", "pixel_auroc_score", "pro_score")


def run_sweep(v_gmm, v_nf, data: dict, rank: int = 0, world: int = 1, batch_size: int = 32, pixel_metrics: bool = True,
              fp_thres: float = 0.3, gmm_seed: int = 1234, transport: str | None = None) -> dict:
    """`data`: {category: (images [n,3,S,S], image_labels [n], pixel_labels [n,1,S,S])}, host (pinned) or device tensors.
    → {"metrics": {category/head: {...}} (complete on every rank), "images": n}.  Returns after the last metric value has
    reached the host.

    One GPU: the metrics of category c run on a side stream while the main stream scores the following categories (their
    host-side reads wait only for their own inputs).  Several GPUs: every rank scores its batches of the whole sweep (the
    round-robin continues across categories: 61 batches over W ranks), then ONE routed exchange sends each validation's
    rows to the rank that evaluates it (parallel.exchange_to_owners; validation p = (category, head) → rank p % W, so the
    sorts are spread over the ranks too), then the metric values are summed into place on every rank.
    `transport` (default: env VITAD_SWEEP_TRANSPORT, else "peer"): "peer" delivers the rows through symmetric memory as they
    are scored (parallel.PeerMailbox: NVLink writes + signals, no collective; the owners' metrics run under the scoring),
    "exchange" is the single routed NCCL exchange after the scoring; "peer" falls back to it where symmetric memory cannot
    be set up (every rank takes the same decision).  A transport named explicitly is also honoured by a single-rank
    process group: the multi-rank code path with one rank, which is how the one-GPU test box checks it against the path
    above."""
    import os

    explicit = transport or os.environ.get("VITAD_SWEEP_TRANSPORT")
    if explicit not in (None, "peer", "exchange"):
        raise ValueError(f"run_sweep: transport {explicit!r} (peer, exchange)")
    transport = explicit or "peer"
    one_gpu = world == 1 and not (explicit and dist.is_available() and dist.is_initialized())
    dev = v_gmm.device
    main = torch.cuda.current_stream(dev)
    heads = ("gmm", "nf")
    names = list(data)

    def evaluate(r, name):
        if not pixel_metrics:
            r = {k: v for k, v in r.items() if not k.startswith("pixel")}
            r["pixel_labels"], r["pixel_scores"] = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
        m = calc_all_metrics_device(r, fp_thres=fp_thres, dataset_name=name, device=dev)
        return {k: v for k, v in m.items() if isinstance(v, float) and k != "fp_thres"}

    # Host-resident data: both heads read the same images, so this rank's batches of a category cross PCIe ONCE, on a copy
    # stream, one category ahead of the scoring (the validators then see device tensors and copy nothing).
    copy_stream = getattr(v_gmm, "_sweep_copy_stream", None)
    if copy_stream is None:
        copy_stream = v_gmm._sweep_copy_stream = torch.cuda.Stream(dev)
    staged, uploads = {}, []  # uploads: keeps the staged tensors referenced until the final synchronize
    deal, acc = [], 0  # first global batch number of every category (the round-robin continues across categories)
    for name in names:
        deal.append(acc)
        acc += -(-int(data[name][0].shape[0]) // batch_size)

    def stage(ci):
        if ci >= len(names) or ci in staged:
            return
        images, labels, masks = data[names[ci]]
        bl = batches(images, labels, masks, batch_size=batch_size)
        ev = None
        if not (torch.is_tensor(images) and images.device == dev):
            off = deal[ci] % world
            with torch.cuda.stream(copy_stream):
                bl = [((b[0].to(dev, non_blocking=True), torch.as_tensor(b[1]).to(dev, non_blocking=True), b[2])
                       if (i + off) % world == rank else b) for i, b in enumerate(bl)]
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            uploads.append(bl)
        staged[ci] = (bl, ev)

    def score(ci, name, dealt):
        images = data[name][0]
        stage(ci)
        bl, ev = staged.pop(ci)
        stage(ci + 1)  # the next category's upload runs under this category's kernels
        if ev is not None:
            main.wait_event(ev)
        v_gmm.gumbel_seed = gmm_seed + ci  # noise field per category, keyed inside by the global batch index
        v_gmm.shard.offset = v_nf.shard.offset = dealt % world
        rg = v_gmm.valid_loop_transformer(bl, keep_origs=False, on_device=True)
        rn = v_nf.valid_loop_transformer_nf(bl, keep_origs=False, on_device=True)
        layout = {"batch_sizes": [int(b[0].shape[0]) for b in bl], "owners": [v_gmm.shard.owner(i) for i in range(len(bl))],
                  "map_shape": (1, int(images.shape[-2]), int(images.shape[-1])), "pixel_label_dtype": torch.uint8}
        return rg, rn, layout, len(bl)

    metrics, n_images, dealt = {}, 0, 0
    if one_gpu:
        side = getattr(v_gmm, "_metrics_stream", None)
        if side is None:
            side = v_gmm._metrics_stream = torch.cuda.Stream(dev)
        pending, keep_alive = [], []  # everything a metric touches stays referenced until the final synchronize

        def finish(entry):
            name, rg, rn, ev = entry
            with torch.cuda.stream(side):
                side.wait_event(ev)
                for tag, r in zip(heads, (rg, rn)):
                    metrics[f"{name}/{tag}"] = evaluate(r, name)

        for ci, name in enumerate(names):
            rg, rn, _layout, nb = score(ci, name, dealt)
            dealt += nb
            n_images += int(data[name][0].shape[0])
            ev = torch.cuda.Event()
            ev.record(main)
            keep_alive.append((rg, rn))
            if pending:  # metrics of the PREVIOUS category: this category's kernels are already queued behind its inputs
                finish(pending.pop(0))
            pending.append((name, rg, rn, ev))
        while pending:
            finish(pending.pop(0))
        main.wait_stream(side)
        torch.cuda.synchronize(dev)
        keep_alive.clear()
    else:
        n_pairs = 2 * len(names)
        table = torch.full((n_pairs, len(METRIC_KEYS)), float("nan"), dtype=torch.float64)

        def fill(p, r):
            m = evaluate(r, names[p // 2])
            for j, key in enumerate(METRIC_KEYS):
                hit = [v for k, v in m.items() if k.startswith(key)]
                if hit:
                    table[p, j] = hit[0]

        pair_owner = [p % world for p in range(n_pairs)]  # validation p = (category, head) → rank p mod W
        box = _mailbox(v_gmm, data, names, pair_owner, world, dev) if transport != "exchange" else None
        if box is not None:
            # Peer-memory delivery: rows go straight into the owner's buffer as soon as they are scored (NVLink writes, no
            # collective), the owner evaluates a validation `lag` categories later on a side stream — by then its rows
            # have arrived and the main stream has the categories in between queued — so only the last ones are exposed.
            side = getattr(v_gmm, "_metrics_stream", None)
            if side is None:
                side = v_gmm._metrics_stream = torch.cuda.Stream(dev)
            sizes = [int(data[n][0].shape[0]) for n in names]
            base, used = [0] * n_pairs, [0] * world
            for p in range(n_pairs):
                base[p] = used[pair_owner[p]]
                used[pair_owner[p]] += sizes[p // 2]
            # how far the metrics trail the scoring: the host blocks on a validation's metric values, so it needs queued
            # kernels ahead of them; 2 * lag validations (at most one per rank) are left for after the last category
            lag = int(os.environ.get("VITAD_SWEEP_LAG", 0)) or max(1, world // 2)
            keep_alive = []

            def deliver(p, res, layout):
                for row0, lo, hi in delivery_plan(layout["batch_sizes"], layout["owners"], rank, base[p]):
                    box.put(pair_owner[p], row0, res, lo, hi)
                box.signal(pair_owner[p], p)  # every rank signals every validation, rows or not

            def finish_category(ci):
                for p in (2 * ci, 2 * ci + 1):
                    if pair_owner[p] == rank:
                        with torch.cuda.stream(side):
                            box.wait_all(p)
                            fill(p, box.rows(base[p], sizes[ci]))

            for ci, name in enumerate(names):
                rg, rn, layout, nb = score(ci, name, dealt)
                dealt += nb
                n_images += sizes[ci]
                deliver(2 * ci, rg, layout)
                deliver(2 * ci + 1, rn, layout)
                keep_alive.append((rg, rn))
                if ci >= lag:
                    finish_category(ci - lag)
            for ci in range(max(0, len(names) - lag), len(names)):
                finish_category(ci)
            main.wait_stream(side)
        else:
            entries = []
            for ci, name in enumerate(names):
                rg, rn, layout, nb = score(ci, name, dealt)
                dealt += nb
                n_images += int(data[name][0].shape[0])
                entries += [{"result": rg, "layout": layout}, {"result": rn, "layout": layout}]
            mine = exchange_to_owners(entries, pair_owner, dev)
            for p, r in mine.items():
                fill(p, r)
        # metric values: one [pairs, keys] table, every rank fills the rows it owns, a sum puts them everywhere (and no
        # rank starts the next sweep — whose rows land in the same buffers — before every owner is done with this one)
        present = (~torch.isnan(table)).to(torch.float64)
        packed = torch.stack((torch.nan_to_num(table, nan=0.0), present)).to(dev)
        dist.all_reduce(packed, op=dist.ReduceOp.SUM)
        packed = packed.cpu()
        torch.cuda.synchronize(dev)
        pro_key = f"pro_score_{fp_thres}fp"
        for p in range(n_pairs):
            name, tag = names[p // 2], heads[p % 2]
            metrics[f"{name}/{tag}"] = {(pro_key if key == "pro_score" else key): float(packed[0, p, j])
                                        for j, key in enumerate(METRIC_KEYS) if packed[1, p, j] > 0}
    used = "one GPU" if one_gpu else ("peer" if box is not None else "exchange")
    return {"metrics": metrics, "images": n_images, "heads_per_image": 2, "transport": used}


def delivery_plan(batch_sizes, owners, rank: int, base: int = 0) -> list:
    """Where the rows a rank scored go in the evaluating rank's buffer.  A validation's rows are stored in global batch
    order from row `base`; `rank` holds the rows of its own batches (owners[b] == rank) back to back in ascending batch
    order.  → [(destination row, first local row, end local row)] per batch of `rank`: both sides derive it from the
    dealing rule, nothing is negotiated."""
    plan, lo, start = [], 0, int(base)
    for size, owner in zip(batch_sizes, owners):
        size = int(size)
        if int(owner) == rank and size > 0:
            plan.append((start, lo, lo + size))
            lo += size
        start += size
    return plan


def _mailbox(holder, data: dict, names: list, pair_owner: list, world: int, dev):
    """The sweep's PeerMailbox, sized for the busiest owner and kept on the validator between sweeps; None (on every rank)
    when symmetric memory is not available."""
    from .parallel import PeerMailbox

    rows = [0] * world
    for p, o in enumerate(pair_owner):
        rows[o] += int(data[names[p // 2]][0].shape[0])
    images = data[names[0]][0]
    key = (max(rows), int(images.shape[-2]), int(images.shape[-1]), world)
    cached = getattr(holder, "_sweep_mailbox", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    ok = torch.ones(1, device=dev)
    box = None
    try:
        box = PeerMailbox(key[0], (1, key[1], key[2]), dev)
        if box.channels < len(pair_owner):
            raise RuntimeError(f"signal pad holds {box.channels} channels, the sweep needs {len(pair_owner)}")
    except Exception as exc:  # noqa: BLE001 — any set-up failure means: use the NCCL exchange, on every rank
        print(f"vitad.sweep: symmetric memory unavailable ({type(exc).__name__}: {exc}); using the NCCL exchange", flush=True)
        ok.zero_()
        box = None
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if float(ok.item()) < 1.0:
        box = None
    holder._sweep_mailbox = (key, box)
    return box


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
