"""Multi-GPU plumbing of the scoring path: one process per GPU (torchrun), a weight replica per rank,
whole batches dealt round-robin (validators.py), and ONE exchange step — gathering per-image scores, maps
and labels on every rank for AUROC / PR-AUC.  The reference is single-process; this is new.
Works with the nccl backend on GPUs and with gloo on CPU (used by the world_size-2/3 tests).

Every rank takes part in the same sequence of collectives whatever it holds: a rank that scored nothing (fewer
batches than ranks — a small category on 8 GPUs) contributes empty payloads instead of raising while its peers wait.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

PAYLOAD = (("image_scores", torch.float32), ("pixel_scores", torch.float32), ("image_labels", torch.int64),
           ("pixel_labels", None))  # pixel_labels keeps the dtype the ranks hold (fp32 from the loader, uint8 on-device rows)


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """→ (rank, world_size, local_rank); initialises torch.distributed when WORLD_SIZE > 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def warm_up(device: torch.device | None = None) -> None:
    """One tiny all_gather so that communicator set-up (NCCL: hundreds of ms) does not land in a timed gather."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    device = _default_device() if device is None else device
    t = torch.zeros(8, device=device)
    out = torch.empty(8 * dist.get_world_size(), device=device)
    dist.all_gather_into_tensor(out, t)
    if device.type == "cuda":
        torch.cuda.synchronize(device)


def _default_device() -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def _as_tensor(a, device, dtype=None) -> torch.Tensor:
    t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
    return t.to(device=device, dtype=dtype if dtype is not None else t.dtype, non_blocking=True)


class PendingGather:
    """Handle of a gather in flight (gather_results(..., async_op=True)): .result() waits for the collectives on the
    calling stream and stitches the global batch order back."""

    def __init__(self, finish):
        self._finish = finish

    def result(self) -> dict:
        return self._finish()


def gather_results(result: dict, num_batches: int, device: torch.device | None = None, async_op: bool = False,
                   as_numpy: bool | None = None, layout: dict | None = None):
    """All-gather the per-rank validator dicts (keys image_scores, pixel_scores, image_labels, pixel_labels,
    batch_index, batch_sizes) and restore the original batch order.  Ranks may hold different numbers of images
    (short tail batches) or none at all: payloads are padded to the per-rank maximum for equal-count
    all_gather_into_tensor.

    Values may be numpy arrays (valid_loop_*) or torch tensors on `device` (valid_loop_*(on_device=True)); tensors that
    already live on the device are gathered in place — no host round trip.  Returns numpy arrays when the input held
    numpy arrays, device tensors otherwise (`as_numpy` overrides).  async_op=True returns a PendingGather after
    enqueueing the payload collectives, so the caller can score the next category while NVLink moves this one.

    `layout` = {"batch_sizes": images of EVERY batch in loader order, "owners": rank of every batch, "map_shape": (1,S,S),
    "pixel_label_dtype": torch dtype}: when the caller knows how the batches were dealt (it always does: the sharding rule
    is deterministic), no metadata has to be exchanged and nothing here synchronises the host with the device."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return PendingGather(lambda: result) if async_op else result
    world = dist.get_world_size()
    if device is None:
        device = _default_device()
    if as_numpy is None:
        as_numpy = not any(torch.is_tensor(result.get(k)) for k, _ in PAYLOAD)
    if "batch_sizes" not in result or "batch_index" not in result:
        raise ValueError("gather_results needs result['batch_index'] and result['batch_sizes'] (images per scored batch)")
    counts_local = np.asarray(result["batch_sizes"], dtype=np.int64).reshape(-1)
    bidx_local = np.asarray(result["batch_index"], dtype=np.int64).reshape(-1)
    n_local, b_local = int(counts_local.sum()), int(counts_local.size)

    if layout is not None:
        return _gather_with_layout(result, num_batches, device, async_op, as_numpy, layout, n_local)
    # -- metadata: image / batch counts, trailing map shape and label dtype of every rank (empty ranks send zeros)
    ps = result.get("pixel_scores")
    map_shape = tuple(int(v) for v in ps.shape[1:]) if ps is not None and n_local > 0 else (0, 0, 0)
    if len(map_shape) != 3:
        raise ValueError(f"pixel_scores must be [n, 1, S, S], got trailing shape {map_shape}")
    pl = result.get("pixel_labels")
    pl_is_u8 = int(n_local > 0 and pl is not None and (pl.dtype in (torch.uint8, np.uint8, torch.bool, np.bool_)))
    meta = torch.tensor([n_local, b_local, *map_shape, pl_is_u8], dtype=torch.int64).to(device)
    metas = torch.empty(world * meta.numel(), device=device, dtype=torch.int64)
    dist.all_gather_into_tensor(metas, meta)
    metas = metas.view(world, -1).cpu().numpy()
    n_max, b_max = int(metas[:, 0].max()), int(metas[:, 1].max())
    if n_max == 0:
        raise ValueError("gather_results: no rank holds any result")
    holder = int(np.argmax(metas[:, 0] > 0))
    map_shape = tuple(int(v) for v in metas[holder, 2:5])
    pl_dtype = torch.uint8 if int(metas[holder, 5]) else torch.float32
    trailing = {"image_scores": (), "pixel_scores": map_shape, "image_labels": (), "pixel_labels": map_shape}

    # -- payloads: padded to n_max rows, gathered on the device
    works, gathered = [], {}

    def gather(name, local, pad_to, dtype, trail):
        t = torch.zeros((pad_to,) + trail, device=device, dtype=dtype)
        if local is not None and pad_to > 0 and (local.shape[0] if hasattr(local, "shape") else len(local)) > 0:
            src = _as_tensor(local, device, dtype)
            t[: src.shape[0]] = src.reshape((src.shape[0],) + trail)
        out = torch.empty((world * pad_to,) + trail, device=device, dtype=dtype)
        w = dist.all_gather_into_tensor(out, t, async_op=True)
        works.append(w)
        gathered[name] = out.view((world, pad_to) + trail)

    for key, dtype in PAYLOAD:
        gather(key, result.get(key) if n_local > 0 else None, n_max, dtype if dtype is not None else pl_dtype, trailing[key])
    gather("batch_index", bidx_local, b_max, torch.int64, ())
    gather("batch_sizes", counts_local, b_max, torch.int64, ())

    def finish():
        for w in works:
            w.wait()
        bidx = gathered["batch_index"].cpu().numpy()
        bsz = gathered["batch_sizes"].cpu().numpy()
        # row of the flattened [world * n_max] gather buffers for every image, in global batch order
        rows = [None] * num_batches
        for r in range(world):
            off = 0
            for j in range(int(metas[r, 1])):
                b, n = int(bidx[r, j]), int(bsz[r, j])
                if not (0 <= b < num_batches) or rows[b] is not None:
                    raise ValueError(f"gather_results: batch {b} reported twice or outside [0, {num_batches})")
                rows[b] = np.arange(r * n_max + off, r * n_max + off + n, dtype=np.int64)
                off += n
        present = [b for b in range(num_batches) if rows[b] is not None]
        order = torch.from_numpy(np.concatenate([rows[b] for b in present])).to(device)
        merged = {}
        for key, _ in PAYLOAD:
            flat = gathered[key].reshape((world * n_max,) + trailing[key])
            t = flat.index_select(0, order)
            merged[key] = t.cpu().numpy() if as_numpy else t
        merged["batch_index"] = np.asarray(present, dtype=np.int64)
        merged["batch_sizes"] = np.asarray([len(rows[b]) for b in present], dtype=np.int64)
        return merged

    return PendingGather(finish) if async_op else finish()


def _gather_with_layout(result, num_batches, device, async_op, as_numpy, layout, n_local):
    """gather_results without the metadata exchange: every rank derives all counts from `layout`."""
    world = dist.get_world_size()
    sizes = np.asarray(layout["batch_sizes"], dtype=np.int64)
    owners = np.asarray(layout["owners"], dtype=np.int64)
    if sizes.size != num_batches or owners.size != num_batches:
        raise ValueError("gather_results: layout must describe every batch of the loader")
    per_rank = np.zeros(world, dtype=np.int64)
    rows = []
    for b in range(num_batches):  # row of image k of batch b inside the [world * n_max] gather buffer
        rows.append((int(owners[b]), int(per_rank[owners[b]]), int(sizes[b])))
        per_rank[owners[b]] += sizes[b]
    if n_local != int(per_rank[dist.get_rank()]):
        raise ValueError(f"gather_results: this rank holds {n_local} images, the layout says {int(per_rank[dist.get_rank()])}")
    n_max = int(per_rank.max())
    map_shape = tuple(int(v) for v in layout["map_shape"])
    pl_dtype = layout.get("pixel_label_dtype", torch.float32)
    trailing = {"image_scores": (), "pixel_scores": map_shape, "image_labels": (), "pixel_labels": map_shape}
    works, gathered = [], {}
    for key, dtype in PAYLOAD:
        dtype = dtype if dtype is not None else pl_dtype
        t = torch.zeros((n_max,) + trailing[key], device=device, dtype=dtype)
        local = result.get(key) if n_local > 0 else None
        if local is not None:
            src = _as_tensor(local, device, dtype)
            t[: src.shape[0]] = src.reshape((src.shape[0],) + trailing[key])
        out = torch.empty((world * n_max,) + trailing[key], device=device, dtype=dtype)
        works.append(dist.all_gather_into_tensor(out, t, async_op=True))
        gathered[key] = out
    order = np.concatenate([np.arange(r * n_max + off, r * n_max + off + n, dtype=np.int64) for r, off, n in rows])
    order_dev = torch.from_numpy(order).to(device, non_blocking=True)

    def finish():
        for w in works:
            w.wait()
        merged = {}
        for key, _ in PAYLOAD:
            t = gathered[key].index_select(0, order_dev)
            merged[key] = t.cpu().numpy() if as_numpy else t
        merged["batch_index"] = np.arange(num_batches, dtype=np.int64)
        merged["batch_sizes"] = sizes.copy()
        return merged

    return PendingGather(finish) if async_op else finish()


def exchange_to_owners(entries: list, pair_owner: list, device: torch.device | None = None) -> dict:
    """One routed exchange for many validation results at once: entry p = {"result": this rank's rows of validation p
    (valid_loop_*(on_device=True); may be empty), "layout": as for gather_results} is needed in full only by rank
    pair_owner[p] (the rank that evaluates its metrics).  Every rank sends each owner exactly the rows it scored — four
    all_to_all_single calls (scores, maps, image labels, pixel labels) for the whole list, with split sizes every rank
    derives from the layouts (no metadata exchange, no host/device synchronisation) — and the owners stitch the global
    batch order back.  → {p: merged result dict (device tensors)} for the entries this rank owns.

    Why not an all_gather per validation as it completes: a collective kernel waits on the device for its peers, and while
    it sits on an SM the persistent 1-CTA-per-SM GEMM kernels of the scoring path cannot be fully resident — every early
    gather turned into a barrier across ranks.  One exchange after the scoring has no such coupling and moves 1/W of
    the bytes."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    me = dist.get_rank() if world > 1 else 0
    if device is None:
        device = _default_device() if world > 1 else torch.device("cpu")
    P = len(entries)
    n_loc = np.zeros((P, world), dtype=np.int64)  # rows of validation p scored by rank r
    for p, e in enumerate(entries):
        sizes = np.asarray(e["layout"]["batch_sizes"], dtype=np.int64)
        owners = np.asarray(e["layout"]["owners"], dtype=np.int64)
        np.add.at(n_loc[p], owners, sizes)
    map_shape = tuple(int(v) for v in entries[0]["layout"]["map_shape"])
    pl_dtype = entries[0]["layout"].get("pixel_label_dtype", torch.float32)
    trailing = {"image_scores": (), "pixel_scores": map_shape, "image_labels": (), "pixel_labels": map_shape}
    owned_by = [[p for p in range(P) if pair_owner[p] == d] for d in range(world)]
    in_split = [int(sum(n_loc[p][me] for p in owned_by[d])) for d in range(world)]
    out_split = [int(sum(n_loc[p][s] for p in owned_by[me])) for s in range(world)]
    received = {}
    for key, dtype in PAYLOAD:
        dtype = dtype if dtype is not None else pl_dtype
        parts = []
        for d in range(world):
            for p in owned_by[d]:
                if n_loc[p][me] > 0:
                    src = _as_tensor(entries[p]["result"][key], device, dtype)
                    parts.append(src.reshape((src.shape[0],) + trailing[key]))
        inp = torch.cat(parts) if parts else torch.empty((0,) + trailing[key], device=device, dtype=dtype)
        if inp.shape[0] != sum(in_split):
            raise ValueError(f"exchange_to_owners: {key}: this rank holds {inp.shape[0]} rows, the layouts say {sum(in_split)}")
        out = torch.empty((sum(out_split),) + trailing[key], device=device, dtype=dtype)
        if world > 1:
            dist.all_to_all_single(out, inp.contiguous(), out_split, in_split)
        else:
            out = inp
        received[key] = out
    # stitch: inside the chunk of source s, my validations follow each other in list order, each with s's batches ascending
    src_base = np.concatenate(([0], np.cumsum(out_split)))[:-1]
    merged = {}
    consumed = np.zeros(world, dtype=np.int64)
    for p in owned_by[me]:
        sizes = np.asarray(entries[p]["layout"]["batch_sizes"], dtype=np.int64)
        owners = np.asarray(entries[p]["layout"]["owners"], dtype=np.int64)
        within = np.zeros(world, dtype=np.int64)
        idx = []
        for b in range(sizes.size):
            s = int(owners[b])
            start = src_base[s] + consumed[s] + within[s]
            idx.append(np.arange(start, start + sizes[b], dtype=np.int64))
            within[s] += sizes[b]
        consumed += n_loc[p]
        order = torch.from_numpy(np.concatenate(idx) if idx else np.zeros(0, np.int64)).to(device, non_blocking=True)
        res = {key: received[key].index_select(0, order) for key, _ in PAYLOAD}
        res["batch_index"] = np.arange(sizes.size, dtype=np.int64)
        res["batch_sizes"] = sizes.copy()
        merged[p] = res
    return merged


class PeerMailbox:
    """Receive buffers for validation results in symmetric memory (torch.distributed._symmetric_memory, CUDA backend:
    cuMem allocations of every rank of the node mapped into each other's address space over NVLink).

    A rank that has scored rows of a validation writes them straight into the buffer of the rank that evaluates it — plain
    copy kernels on its own stream, 700 GB/s per direction through NVSwitch — and then raises a stream-ordered signal for
    that validation; the owner waits for one signal per rank on a side stream.  There is no collective and no rendezvous
    with the receiver: a collective kernel waits on the device for its peers and, while it sits on an SM, the persistent
    one-CTA-per-SM kernels of the scoring path cannot be resident (exchange_to_owners pays that once, after the scoring;
    here the rows arrive while both sides keep scoring, so the owner's metrics can run under the scoring again).

    Every rank allocates the same capacity (rows of the busiest owner).  put / signal / wait_all are stream-ordered and do
    not block the host; a signal raised twice without a wait in between, or a wait without its signals, ends in the
    device-side time-out of the signal kernels (60 s) instead of a hang."""

    SIGNAL_TIMEOUT_MS = 60000

    def __init__(self, rows: int, map_shape: tuple, device: torch.device, pixel_label_dtype=torch.uint8, group=None):
        import torch.distributed._symmetric_memory as symm

        self.rows_cap, self.map_shape, self.device = int(rows), tuple(int(v) for v in map_shape), device
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        group = dist.group.WORLD if group is None else group
        per_map = int(np.prod(self.map_shape))
        self._spec = {"image_scores": ((), torch.float32), "pixel_scores": (self.map_shape, torch.float32),
                      "image_labels": ((), torch.int64), "pixel_labels": (self.map_shape, pixel_label_dtype)}
        self._local, self._hdl, self._peer = {}, {}, {}
        for key, (trail, dtype) in self._spec.items():
            n = self.rows_cap * (per_map if trail else 1)
            t = symm.empty(max(n, 1), dtype=dtype, device=device)
            self._local[key] = t
            self._hdl[key] = symm.rendezvous(t, group)  # collective, once
            self._peer[key] = [self._hdl[key].get_buffer(r, (self.rows_cap,) + trail, dtype, 0) if n else None
                               for r in range(self.world)]
        self._sig = self._hdl["image_scores"]  # one signal pad serves all four payloads
        self.channels = self._sig.signal_pad_size // 4 // self.world

    def put(self, owner: int, row0: int, result: dict, lo: int, hi: int) -> None:
        """rows [lo, hi) of this rank's result → rows [row0, row0 + hi - lo) of `owner`'s buffers (current stream)."""
        if hi <= lo:
            return
        if row0 < 0 or row0 + (hi - lo) > self.rows_cap:
            raise ValueError(f"PeerMailbox.put: rows [{row0}, {row0 + hi - lo}) outside the capacity {self.rows_cap}")
        for key, (trail, dtype) in self._spec.items():
            src = result[key][lo:hi]
            self._peer[key][owner].narrow(0, row0, hi - lo).copy_(src.reshape((hi - lo,) + trail), non_blocking=True)

    def signal(self, owner: int, channel: int) -> None:
        """After everything this stream has written so far: tell `owner` that this rank is done with `channel`."""
        self._sig.put_signal(owner, channel, self.SIGNAL_TIMEOUT_MS)

    def wait_all(self, channel: int) -> None:
        """Current stream waits until every rank (this one included) has signalled `channel`; the signals are consumed."""
        for src in range(self.world):
            self._sig.wait_signal(src, channel, self.SIGNAL_TIMEOUT_MS)

    def rows(self, row0: int, n: int) -> dict:
        """This rank's received rows [row0, row0 + n) as a result dictionary (views of the buffers)."""
        out = {}
        for key, (trail, _dtype) in self._spec.items():
            out[key] = self._local[key][: self.rows_cap * (int(np.prod(trail)) if trail else 1)].view(
                (self.rows_cap,) + trail).narrow(0, row0, n)
        return out
