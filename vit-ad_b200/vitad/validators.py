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


def _place(module, device):
    """module.to(device).eval() without touching a module that already lives there: nn.Module.to() always walks
    _apply(), which drops the packed fp16 weight copies the CUDA modules keep (and re-packing costs ms per call)."""
    p = next(module.parameters(), None)
    if p is None or p.device != device:
        module.to(device)
    return module.eval()


def _metrics(result: dict, fp_thres: float, dataset_name: str, on_device: bool) -> dict:
    if on_device:
        from .gpu_metrics import calc_all_metrics_device

        return calc_all_metrics_device(result, fp_thres=fp_thres, dataset_name=dataset_name)
    from .metrics import calc_all_metrics

    return calc_all_metrics(result, fp_thres=fp_thres, dataset_name=dataset_name)


def _shared_seed() -> int:
    """One fresh 62-bit seed from torch's CPU generator (follows torch.manual_seed); with torch.distributed
    initialised every rank takes rank 0's, so all shards of one validation draw the same noise field."""
    seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64)
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
            t = seed.to(dev)
            dist.broadcast(t, src=0)
            seed = t.cpu()
    except ImportError:  # pragma: no cover
        pass
    return int(seed.item())


class _BatchSharding:
    """Batch-granular round-robin sharding: batch i belongs to rank (i + offset) % world_size.  `offset` lets a caller that
    validates many small data sets in a row (the 15-category sweep) continue the round-robin across them instead of
    handing batch 0 of every set to rank 0."""

    def __init__(self, rank: int = 0, world_size: int = 1, offset: int = 0):
        if not (0 <= rank < world_size):
            raise ValueError(f"rank {rank} outside world of size {world_size}")
        self.rank, self.world_size, self.offset = rank, world_size, offset

    def owner(self, batch_index: int) -> int:
        return (batch_index + self.offset) % self.world_size

    def mine(self, batch_index: int) -> bool:
        return self.owner(batch_index) == self.rank


class _Pipelined:
    def _img_size(self) -> int:
        fe = getattr(self, "feature_extractor", None)
        return int((self.model if fe is None else fe).img_size)

    """Three-stream batch pipeline shared by the validators: the H2D copy of batch k+1 (copy-in stream) and the D2H
    copy of batch k-1's scores/maps into pinned buffers (copy-out stream) overlap the kernels of batch k (current
    stream).  The reference moves one batch at a time and synchronises 33 times per batch (ValidatorMDN.py:123-168).
    Host tensors should be pinned (DataLoader(pin_memory=True)); pageable input still works, its copy just blocks."""

    DEPTH = 2  # batches in flight

    def _stream_batches(self, batches: Iterable, compute: Callable, copy: bool = True, to_host: bool = True):
        """`batches` yields (batch_index, host_or_device_images, extra).  `compute(images_dev, batch_index)` returns a
        tuple of device tensors.  Yields (batch_index, tuple of host numpy arrays, extra) in input order.  With
        copy=False the arrays are views of the pinned staging buffers, valid until the next item is requested.
        to_host=False: no D2H at all — yields the device tensors `compute` returned, without synchronising (the H2D
        prefetch of the next batch still overlaps); for callers that keep working on the device (gather + metrics)."""
        dev = self.device
        main = torch.cuda.current_stream(dev)
        if not hasattr(self, "_s_in"):
            self._s_in, self._s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            self._pinned = {}
        s_in, s_out = self._s_in, self._s_out
        inflight = []  # (batch_index, extra, [pinned host tensors], done event)

        def stage_in(item):
            bi, images, extra = item
            if torch.is_tensor(images) and images.device == dev:
                return bi, images, extra, None
            with torch.cuda.stream(s_in):  # allocated from the copy stream's pool; record_stream() guards its reuse
                d = torch.as_tensor(images).to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s_in)
            d.record_stream(main)
            return bi, d, extra, ev

        def drain(n_keep):
            while len(inflight) > n_keep:
                bi, extra, host, ev, slot = inflight.pop(0)
                ev.synchronize()
                yield bi, tuple((h.numpy().copy() if copy else h.numpy()) for h in host), extra
                self._free_slots.append(slot)

        self._free_slots = list(range(self.DEPTH + 1))
        it = iter(batches)
        nxt = next(it, None)
        staged = stage_in(nxt) if nxt is not None else None
        while staged is not None:
            bi, d_img, extra, ev_in = staged
            if ev_in is not None:
                main.wait_event(ev_in)
            outs = compute(d_img, bi)
            ev_c = torch.cuda.Event()
            ev_c.record(main)
            nxt = next(it, None)  # enqueue the next batch's H2D while this batch computes
            staged = stage_in(nxt) if nxt is not None else None
            if not to_host:
                yield bi, tuple(outs), extra
                continue
            slot = self._free_slots.pop(0)
            s_out.wait_event(ev_c)
            host = []
            with torch.cuda.stream(s_out):
                for j, o in enumerate(outs):
                    # one pinned buffer per (slot, output), grown to the largest batch seen: a short last batch
                    # (the reference's loaders do not drop it) reuses a view instead of a fresh cudaHostAlloc
                    key = (slot, j, tuple(o.shape[1:]), o.dtype)
                    buf = self._pinned.get(key)
                    if buf is None or buf.shape[0] < o.shape[0]:
                        buf = self._pinned[key] = torch.empty(o.shape, dtype=o.dtype).pin_memory()
                    view = buf[: o.shape[0]]
                    view.copy_(o, non_blocking=True)
                    o.record_stream(s_out)
                    host.append(view)
                ev_o = torch.cuda.Event()
                ev_o.record(s_out)
            inflight.append((bi, extra, host, ev_o, slot))
            yield from drain(self.DEPTH - 1)
        yield from drain(0)

    def iter_scores(self, dataloader: Iterable):
        """Streaming form of the valid loops: yields (batch_index, image_scores [B], pixel_scores [B,1,S,S]) as fp32
        numpy per batch of `dataloader` ((images, pixel_labels, image_labels) tuples), pipelined as above.  The arrays
        are views of pinned staging buffers: consume (or copy) them before asking for the next batch."""
        def mine():
            for bi, (images, _pl, _il) in enumerate(dataloader):
                if self.shard.mine(bi):
                    yield bi, images, None

        with torch.no_grad():
            for bi, out, _ in self._stream_batches(mine(), self.score_batch, copy=False):
                yield bi, out[0], out[1]


class ValidatorMdn(_Pipelined):
    """Drop-in for src/pipeline/ValidatorMDN.py:27-183 (transformer encoders)."""

    def __init__(self, gmm_model: list, feature_extractor, dataloader, props: dict, weights_object: list | None = None,
                 weights_base_path: str = "", weights_name: list | str = "", rank: int = 0, world_size: int = 1,
                 gumbel: Callable[[int, tuple], torch.Tensor] | None = None, gumbel_seed: int | None = None):
        self.gmm_model = gmm_model
        self.feature_extractor = feature_extractor
        self.dataloader = dataloader
        self.dataset_name = f"{props['dataset']}_{props['dataclass']}"
        self.run_name = f"gmm_{props['num_gaussians']}"
        self.props = props
        self.device = _require_cuda()
        self.shard = _BatchSharding(rank, world_size)
        # Gumbel noise of the mixing weights (the reference draws it from torch's global generator in every call,
        # MixtureDensityNetwork.py:62).  `gumbel`: (batch_index, shape) -> explicit noise tensor (parity tests).  Otherwise the
        # kernel generates it from (gumbel_seed, GLOBAL batch index, token, mixture): batch i scores the same on whichever
        # rank it lands, so a sharded sweep reproduces the unsharded scores bit for bit.  gumbel_seed=None draws one seed
        # per validator from torch's CPU generator (rank 0's, when torch.distributed is up).
        self.gumbel = gumbel
        self.gumbel_seed = _shared_seed() if gumbel_seed is None else int(gumbel_seed)
        _load_weights(self.gmm_model, weights_object, weights_base_path, weights_name)

    # -- one batch: host images in, host scores/maps out -------------------------------------------
    def score_batch(self, images: torch.Tensor, batch_index: int = 0):
        """→ (image_scores [B] , pixel_scores [B,1,S,S]) as device tensors for one batch of images
        (host or device, fp32 NCHW in [0,1]).  Body of ValidatorMDN.py:123-162."""
        model = self.gmm_model[0]
        fe = self.feature_extractor
        images = ops.prepare_images(images.to(self.device, non_blocking=True), self._img_size())
        features = fe(images)
        x = features.patch_embedding
        g = None
        if self.gumbel is not None:
            g = self.gumbel(batch_index, (x.shape[0], x.shape[1], model.num_gaussians)).to(self.device, non_blocking=True)
        prob, image_scores = model.score(x, g, seed=self.gumbel_seed, batch_index=batch_index)
        grid = int(fe.img_size / fe.patch_size)
        pixel_scores, _ = ops.bilinear_up(prob.view(-1, grid, grid), fe.img_size, align_corners=True,
                                          post_one_minus=True)
        return image_scores, pixel_scores

    def valid_loop_transformer(self, dataloader: Iterable, keep_origs: bool = True, on_device: bool = False) -> dict:
        """ValidatorMDN.py:104-183.  `keep_origs=False` drops the copy of the input images from the result (the
        reference returns them for its plots).  `on_device=True`: the result rows stay on the GPU as torch tensors
        (no per-batch D2H; feed them to parallel.gather_results / gpu_metrics.calc_all_metrics_device)."""
        _reject_cnn_encoder(self.feature_extractor, "ValidatorMdn", "valid_loop_resnet (ValidatorMDN.py:185-273)")
        _place(self.gmm_model[0], self.device)
        _place(self.feature_extractor, self.device)
        return _collect(self, self.score_batch, dataloader, self.shard, keep_origs=keep_origs, on_device=on_device)

    def valid_loop_resnet(self, dataloader: Iterable, *a, **kw) -> dict:
        """ValidatorMDN.py:185-273 serves the CNN encoders (ResNet / EfficientNet feature pyramids), which are outside this
        implementation's scope (transformer encoders only)."""
        raise NotImplementedError(_CNN_MSG.format(cls="ValidatorMdn", what="valid_loop_resnet (ValidatorMDN.py:185-273)"))

    def calc_all_metrics(self, centering: bool = False, new_wandb_run: bool = True, on_device: bool = True) -> dict:
        """ValidatorMDN.py:71-102 without the W&B / matplotlib side effects: returns the metric dict
        (`on_device`: sort-based metrics on the GPU, vitad.gpu_metrics; False: the reference's sklearn calls)."""
        loader = self.dataloader.get_dataloader(centering=centering)
        result = self.valid_loop_transformer(loader, keep_origs=False, on_device=on_device)
        return _metrics(result, self.props.get("fp_thres", 0.3), self.dataset_name, on_device)


class _Rows:
    """Result rows appended batch by batch into one geometrically grown array: one copy per batch (from the pinned
    staging view) instead of a per-batch copy plus a final np.concatenate."""

    def __init__(self):
        self.buf, self.n = None, 0

    def append(self, a):
        a = np.asarray(a)
        k = a.shape[0]
        if self.buf is None:
            self.buf = np.empty((max(4 * k, 64),) + a.shape[1:], dtype=a.dtype)
        elif self.n + k > self.buf.shape[0]:
            grown = np.empty((max(2 * self.buf.shape[0], self.n + k),) + self.buf.shape[1:], dtype=self.buf.dtype)
            grown[: self.n] = self.buf[: self.n]
            self.buf = grown
        self.buf[self.n: self.n + k] = a
        self.n += k

    def get(self):
        return self.buf[: self.n]


_CNN_MSG = ("{cls}: {what} needs a CNN feature extractor (ResNet / EfficientNet), which the B200 scoring path does not "
            "provide — it covers the transformer encoders (enc_deit, enc_vit, enc_esvit; SURVEY.md §8).  Use the "
            "reference's PyTorch validator for CNN encoders.")


def _reject_cnn_encoder(fe, cls: str, what: str) -> None:
    """The reference dispatches on the encoder type (ValidatorMDN.py:80-92, ValidatorNF.py:73-84); an encoder that is not
    one of this package's transformer encoders gets a clear error instead of an AttributeError deep in the loop."""
    from .encoders import TransformerEncoder

    if not isinstance(fe, TransformerEncoder):
        raise NotImplementedError(_CNN_MSG.format(cls=cls, what=what) + f"  (got {type(fe).__name__})")


def _collect(validator, loop_body, dataloader, shard, with_recons=False, keep_origs=True, on_device=False):
    """Shared batch loop: `loop_body(images, batch_index)` → (scores, maps[, recons]) device tensors, run through the
    validator's copy/compute pipeline; returns the reference's result dictionary (fp32 numpy; torch tensors on the
    validator's device with on_device=True — labels as int64 / uint8)."""
    index, sizes = [], []

    def mine():
        for bi, (images, pixel_labels, image_labels) in enumerate(dataloader):
            if shard.mine(bi):
                yield bi, images, (images if keep_origs else None, pixel_labels, image_labels)

    if on_device:
        dev = validator.device
        rows = {k: [] for k in ("image_scores", "pixel_scores", "image_labels", "pixel_labels", "origs", "recons")}
        with torch.no_grad():
            for bi, out, (images, pixel_labels, image_labels) in validator._stream_batches(mine(), loop_body, to_host=False):
                rows["image_scores"].append(out[0])
                rows["pixel_scores"].append(out[1])
                if with_recons:
                    rows["recons"].append(out[2])
                rows["image_labels"].append(torch.as_tensor(image_labels).reshape(-1).to(torch.int64))
                rows["pixel_labels"].append(torch.as_tensor(pixel_labels))
                if keep_origs:
                    rows["origs"].append(torch.as_tensor(images))
                index.append(bi)
                sizes.append(int(out[0].shape[0]))
            res = {}
            for k, v in rows.items():
                if not v:
                    continue
                if k == "pixel_labels":  # masks are 0/1: one byte per pixel crosses PCIe / NVLink instead of four
                    t = torch.cat([(p != 0).to(torch.uint8) for p in v])
                else:
                    t = torch.cat(v)
                res[k] = t.to(dev, non_blocking=True)
        res["batch_index"], res["batch_sizes"] = np.asarray(index, dtype=np.int64), np.asarray(sizes, dtype=np.int64)
        return res

    acc = {k: _Rows() for k in ("image_scores", "pixel_scores", "image_labels", "pixel_labels", "origs", "recons")}
    with torch.no_grad():
        # copy=False: `out` are views of the pinned staging buffers, consumed (copied into the result rows) right here
        for bi, out, (images, pixel_labels, image_labels) in validator._stream_batches(mine(), loop_body, copy=False):
            acc["image_scores"].append(out[0])
            acc["pixel_scores"].append(out[1])
            if with_recons:
                acc["recons"].append(out[2])
            acc["image_labels"].append(image_labels)
            acc["pixel_labels"].append(pixel_labels)
            if keep_origs:
                acc["origs"].append(images.cpu() if torch.is_tensor(images) else images)
            index.append(bi)
            sizes.append(int(out[0].shape[0]))
    res = {k: v.get() for k, v in acc.items() if v.n}
    res["batch_index"], res["batch_sizes"] = np.asarray(index, dtype=np.int64), np.asarray(sizes, dtype=np.int64)
    return res


BLOCK_INDEX_DEIT = 0  # src/pipeline/ValidatorNF.py:23


class ValidatorNF(_Pipelined):
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
        images = ops.prepare_images(images.to(self.device, non_blocking=True), self._img_size())
        embedding = self.feature_extractor(images, block_index=BLOCK_INDEX_DEIT).patch_embedding
        result = self.nf_model[0].forward_tokens(embedding)
        return result.image_max, result.anomaly_score_map

    def valid_loop_transformer_nf(self, dataloader: Iterable, keep_origs: bool = True, on_device: bool = False) -> dict:
        _reject_cnn_encoder(self.feature_extractor, "ValidatorNF", "valid_loop_cnn_nf (ValidatorNF.py:166-219)")
        _place(self.nf_model[0], self.device)
        _place(self.feature_extractor, self.device)
        return _collect(self, self.score_batch, dataloader, self.shard, keep_origs=keep_origs, on_device=on_device)

    def valid_loop_cnn_nf(self, dataloader: Iterable, *a, **kw) -> dict:
        """ValidatorNF.py:166-219 serves the CNN encoders; outside this implementation's scope."""
        raise NotImplementedError(_CNN_MSG.format(cls="ValidatorNF", what="valid_loop_cnn_nf (ValidatorNF.py:166-219)"))

    def calc_all_metrics(self, centering: bool = False, new_wandb_run: bool = True, on_device: bool = True) -> dict:
        result = self.valid_loop_transformer_nf(self.dataloader.get_dataloader(centering=centering), keep_origs=False,
                                                on_device=on_device)
        return _metrics(result, self.props.get("fp_thres", 0.3), self.dataset_name, on_device)


class ValidatorRecon(_Pipelined):
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
        images = ops.prepare_images(images.to(self.device, non_blocking=True), self._img_size())
        # uint8 pixels -> the fp32 [0,1] tensor ToTensor produces (the L2 map compares against it)
        images = images.to(torch.float32).div_(255.0) if images.dtype == torch.uint8 else images.to(torch.float32)
        output = self.model(images)
        amap, score = self.model.anomaly_map_and_score(output.reconstruction, images)
        return score, amap, output.reconstruction

    def valid_loop_mse(self, dataloader: Iterable, keep_origs: bool = True, on_device: bool = False) -> dict:
        _place(self.model, self.device)
        return _collect(self, self.score_batch, dataloader, self.shard, with_recons=True, keep_origs=keep_origs,
                        on_device=on_device)

    def calc_all_metrics(self, centering: bool = False, new_wandb_run: bool = True, on_device: bool = True) -> dict:
        result = self.valid_loop_mse(self.dataloader.get_dataloader(centering=centering), keep_origs=False,
                                     on_device=on_device)
        return _metrics(result, self.props.get("fp_thres", 0.3), self.dataset_name, on_device)
