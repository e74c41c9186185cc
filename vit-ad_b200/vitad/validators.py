"""Validators of the scoring path under the reference's names (src/pipeline/Validator{MDN,NF,Recon}.py).

Same constructors and the same ``valid_loop_*`` result dictionaries (``image_scores, pixel_scores,
image_labels, pixel_labels, origs[, recons]`` as fp32 numpy), but each batch is one H2D copy, a handful of
CUDA launches and one D2H copy of [scores | maps] instead of the reference's 33 device→host syncs per batch
(ValidatorMDN.py:133-168).  The reference shards nothing; here a validator can be given ``rank``/``world_size``
and then scores only every world_size-th batch (the GMM score is coupled across a batch through the
batch-global max, MixtureDensityNetwork.py:90-92, so the shard unit is a whole batch) — see ``parallel.py``.
"""
from __future__ import annotations

import os
from typing import Callable, Iterable

import numpy as np
import torch

from . import ops


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("vitad validators need a CUDA device (B200): this implementation has no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _load_weights(models, weights_object, weights_base_path, weights_name):
    """ValidatorMDN.py:61-69 / ValidatorNF.py:56-64: state dicts given directly or read from .pth files."""
    if weights_object is not None:
        for i, model in enumerate(models):
            model.load_state_dict(weights_object[i])
    elif weights_name:
        for i, model in enumerate(models):
            path = os.path.join(weights_base_path, weights_name[i])
            model.load_state_dict(torch.load(path, map_location=torch.device("cpu")))


class _BatchSharding:
    """Batch-granular round-robin sharding: batch i belongs to rank i % world_size."""

    def __init__(self, rank: int = 0, world_size: int = 1):
        if not (0 <= rank < world_size):
            raise ValueError(f"rank {rank} outside world of size {world_size}")
        self.rank, self.world_size = rank, world_size

    def mine(self, batch_index: int) -> bool:
        return batch_index % self.world_size == self.rank


class ValidatorMdn:
    """Drop-in for src/pipeline/ValidatorMDN.py:27-183 (transformer encoders)."""

    def __init__(self, gmm_model: list, feature_extractor, dataloader, props: dict, weights_object: list | None = None,
                 weights_base_path: str = "", weights_name: list | str = "", rank: int = 0, world_size: int = 1,
                 gumbel: Callable[[int, tuple], torch.Tensor] | None = None):
        self.gmm_model = gmm_model
        self.feature_extractor = feature_extractor
        self.dataloader = dataloader
        self.dataset_name = f"{props['dataset']}_{props['dataclass']}"
        self.run_name = f"gmm_{props['num_gaussians']}"
        self.props = props
        self.device = _require_cuda()
        self.shard = _BatchSharding(rank, world_size)
        self.gumbel = gumbel  # (batch_index, shape) -> noise tensor; None = draw on the device like the reference
        _load_weights(self.gmm_model, weights_object, weights_base_path, weights_name)

    # -- one batch: host images in, host scores/maps out -------------------------------------------
    def score_batch(self, images: torch.Tensor, batch_index: int = 0):
        """→ (image_scores [B] , pixel_scores [B,1,S,S]) as device tensors for one batch of images
        (host or device, fp32 NCHW in [0,1]).  Body of ValidatorMDN.py:123-162."""
        model = self.gmm_model[0]
        fe = self.feature_extractor
        images = images.to(self.device, non_blocking=True)
        features = fe(images)
        x = features.patch_embedding
        g = None
        if self.gumbel is not None:
            g = self.gumbel(batch_index, (x.shape[0], x.shape[1], model.num_gaussians)).to(self.device, non_blocking=True)
        prob, image_scores = model.score(x, g)
        grid = int(fe.img_size / fe.patch_size)
        pixel_scores, _ = ops.bilinear_up(prob.view(-1, grid, grid), fe.img_size, align_corners=True,
                                          post_one_minus=True)
        return image_scores, pixel_scores

    def valid_loop_transformer(self, dataloader: Iterable) -> dict:
        model = self.gmm_model[0]
        model.to(self.device).eval()
        self.feature_extractor.to(self.device).eval()
        out_s, out_p, lab_i, lab_p, origs, index, sizes = [], [], [], [], [], [], []
        with torch.no_grad():
            for bi, (images, pixel_labels, image_labels) in enumerate(dataloader):
                if not self.shard.mine(bi):
                    continue
                s, p = self.score_batch(images, bi)
                out_s.append(s.cpu().numpy())
                out_p.append(p.cpu().numpy())
                lab_i.append(np.asarray(image_labels))
                lab_p.append(np.asarray(pixel_labels))
                origs.append(np.asarray(images.cpu() if torch.is_tensor(images) else images))
                index.append(bi)
                sizes.append(int(s.shape[0]))
        return {
            "image_scores": np.concatenate(out_s, axis=0),
            "pixel_scores": np.concatenate(out_p, axis=0),
            "image_labels": np.concatenate(lab_i, axis=0),
            "pixel_labels": np.concatenate(lab_p, axis=0),
            "origs": np.concatenate(origs, axis=0),
            "batch_index": np.asarray(index),
            "batch_sizes": np.asarray(sizes),
        }

    def calc_all_metrics(self, centering: bool = False, new_wandb_run: bool = True) -> dict:
        """ValidatorMDN.py:71-102 without the W&B / matplotlib side effects: returns the metric dict."""
        from .metrics import calc_all_metrics

        loader = self.dataloader.get_dataloader(centering=centering)
        result = self.valid_loop_transformer(loader)
        return calc_all_metrics(result, fp_thres=self.props.get("fp_thres", 0.3), dataset_name=self.dataset_name)


def _collect(loop_body, dataloader, shard, with_recons=False):
    """Shared batch loop: `loop_body(images, batch_index)` → (scores, maps[, recons]) device tensors."""
    acc = {k: [] for k in ("image_scores", "pixel_scores", "image_labels", "pixel_labels", "origs", "recons")}
    index, sizes = [], []
    with torch.no_grad():
        for bi, (images, pixel_labels, image_labels) in enumerate(dataloader):
            if not shard.mine(bi):
                continue
            out = loop_body(images, bi)
            acc["image_scores"].append(out[0].cpu().numpy())
            acc["pixel_scores"].append(out[1].cpu().numpy())
            if with_recons:
                acc["recons"].append(out[2].cpu().numpy())
            acc["image_labels"].append(np.asarray(image_labels))
            acc["pixel_labels"].append(np.asarray(pixel_labels))
            acc["origs"].append(np.asarray(images.cpu() if torch.is_tensor(images) else images))
            index.append(bi)
            sizes.append(int(out[0].shape[0]))
    res = {k: np.concatenate(v, axis=0) for k, v in acc.items() if v}
    res["batch_index"], res["batch_sizes"] = np.asarray(index), np.asarray(sizes)
    return res


BLOCK_INDEX_DEIT = 0  # src/pipeline/ValidatorNF.py:23


class ValidatorNF:
    """Drop-in for src/pipeline/ValidatorNF.py:26-164 (transformer encoders)."""

    def __init__(self, nf_model: list, feature_extractor, dataloader, props: dict, weights_object: list | None = None,
                 weights_base_path: str = "", weights_name: list | str = "", rank: int = 0, world_size: int = 1):
        self.nf_model = nf_model
        self.dataloader = dataloader
        self.feature_extractor = feature_extractor
        self.dataset_name = f"{props['dataset']}_{props['dataclass']}"
        self.run_name = "nf"
        self.props = props
        self.device = _require_cuda()
        self.shard = _BatchSharding(rank, world_size)
        _load_weights(self.nf_model, weights_object, weights_base_path, weights_name)

    def score_batch(self, images: torch.Tensor, batch_index: int = 0):
        """ValidatorNF.py:123-142 → (image_scores [B], anomaly maps [B,1,S,S])."""
        images = images.to(self.device, non_blocking=True)
        embedding = self.feature_extractor(images, block_index=BLOCK_INDEX_DEIT).patch_embedding
        result = self.nf_model[0].forward_tokens(embedding)
        return result.image_max, result.anomaly_score_map

    def valid_loop_transformer_nf(self, dataloader: Iterable) -> dict:
        self.nf_model[0].to(self.device).eval()
        self.feature_extractor.to(self.device).eval()
        return _collect(self.score_batch, dataloader, self.shard)

    def calc_all_metrics(self, centering: bool = False, new_wandb_run: bool = True) -> dict:
        from .metrics import calc_all_metrics

        result = self.valid_loop_transformer_nf(self.dataloader.get_dataloader(centering=centering))
        return calc_all_metrics(result, fp_thres=self.props.get("fp_thres", 0.3), dataset_name=self.dataset_name)


class ValidatorRecon:
    """Drop-in for src/pipeline/ValidatorRecon.py:21-136: reconstruction → per-pixel L2 map → amax."""

    def __init__(self, model, dataloader, props: dict, weights_object: dict | None = None,
                 weights_base_path: str = "", weights_name: str = "", rank: int = 0, world_size: int = 1):
        self.model = model
        self.dataloader = dataloader
        self.dataset_name = f"{props['dataset']}_{props['dataclass']}"
        self.run_name = f"recon_{type(model.decoder).__name__}"
        self.props = props
        self.device = _require_cuda()
        self.shard = _BatchSharding(rank, world_size)
        if weights_object is not None:
            model.load_state_dict(weights_object)
        elif weights_name:
            model.load_state_dict(torch.load(os.path.join(weights_base_path, weights_name),
                                             map_location=torch.device("cpu")))

    def score_batch(self, images: torch.Tensor, batch_index: int = 0):
        """ValidatorRecon.py:107-116 → (image_scores [B], pixel_scores [B,1,S,S], reconstruction)."""
        images = images.to(self.device, non_blocking=True).to(torch.float32)
        output = self.model(images)
        amap, score = self.model.anomaly_map_and_score(output.reconstruction, images)
        return score, amap, output.reconstruction

    def valid_loop_mse(self, dataloader: Iterable) -> dict:
        self.model.to(self.device).eval()
        return _collect(self.score_batch, dataloader, self.shard, with_recons=True)

    def calc_all_metrics(self, centering: bool = False, new_wandb_run: bool = True) -> dict:
        from .metrics import calc_all_metrics

        result = self.valid_loop_mse(self.dataloader.get_dataloader(centering=centering))
        return calc_all_metrics(result, fp_thres=self.props.get("fp_thres", 0.3), dataset_name=self.dataset_name)
