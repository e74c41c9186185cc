"""Multi-GPU plumbing of the scoring path: one process per GPU (torchrun), a weight replica per rank,
whole batches dealt round-robin (validators.py), and ONE exchange step — gathering per-image scores, maps
and labels on every rank for AUROC / PR-AUC.  The reference is single-process; this is new.
Works with the nccl backend on GPUs and with gloo on CPU (used by the world_size-2 tests).
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist


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


def gather_results(result: dict, num_batches: int, device: torch.device | None = None) -> dict:
    """All-gather the per-rank validator dicts (keys image_scores, pixel_scores, image_labels, pixel_labels,
    batch_index [, origs]) and restore the original batch order.  Ranks may hold different numbers of images
    (short tail batches): payloads are padded to the per-rank maximum for equal-count all_gather_into_tensor.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return result
    world = dist.get_world_size()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    counts_local = np.asarray(result.get("batch_sizes", []), dtype=np.int64)
    if counts_local.size == 0:  # derive from batch_index: caller did not record sizes -> equal split unknown
        raise ValueError("gather_results needs result['batch_sizes'] (images per scored batch)")
    n_local = int(counts_local.sum())
    meta = torch.tensor([n_local, len(counts_local)], device=device, dtype=torch.int64)
    metas = torch.empty(world * 2, device=device, dtype=torch.int64)
    dist.all_gather_into_tensor(metas, meta)
    metas = metas.view(world, 2).cpu().numpy()
    n_max, b_max = int(metas[:, 0].max()), int(metas[:, 1].max())

    def gather(arr: np.ndarray, pad_to: int, dtype: torch.dtype) -> list[np.ndarray]:
        t = torch.zeros((pad_to,) + arr.shape[1:], device=device, dtype=dtype)
        t[: arr.shape[0]] = torch.as_tensor(arr, dtype=dtype).to(device)
        out = torch.empty((world * pad_to,) + arr.shape[1:], device=device, dtype=dtype)
        dist.all_gather_into_tensor(out, t.contiguous())
        return list(out.view((world, pad_to) + arr.shape[1:]).cpu().numpy())

    per_rank = {}
    for key, dtype in (("image_scores", torch.float32), ("pixel_scores", torch.float32),
                       ("image_labels", torch.int64), ("pixel_labels", torch.float32)):
        per_rank[key] = gather(np.asarray(result[key]), n_max, dtype)
    bidx = gather(np.asarray(result["batch_index"], dtype=np.int64), b_max, torch.int64)
    bsz = gather(counts_local, b_max, torch.int64)

    # stitch back in global batch order
    pieces = {k: [None] * num_batches for k in per_rank}
    for r in range(world):
        off = 0
        for j in range(int(metas[r, 1])):
            b, n = int(bidx[r][j]), int(bsz[r][j])
            for k in per_rank:
                pieces[k][b] = per_rank[k][r][off : off + n]
            off += n
    merged = {k: np.concatenate([p for p in v if p is not None], axis=0) for k, v in pieces.items()}
    merged["batch_index"] = np.arange(num_batches)
    return merged
