"""torch custom-op registration of the C-ABI entry points: ``torch.ops.vitad.*`` (north_star: "exposed as torch custom
ops behind a thin C-ABI layer").  Every op has a schema, a CUDA implementation (ctypes call into
libvitad.so on the current stream) and a fake (meta) implementation, so the dispatcher, FakeTensorMode and
``torch.compile``d callers see shapes and dtypes without running a kernel.  There is NO CPU implementation: calling an
op with CPU tensors fails in the dispatcher ("no kernel for CPU") — the product has no CPU path.

Model weights are not op arguments: a module registers itself here and passes an integer handle (its packed fp16
weights, C structs and workspaces are per-module state the library reads through raw pointers).  Handles are weak — a
collected module frees its slot.
"""
from __future__ import annotations

import itertools
import weakref

import torch
from torch import Tensor

_HANDLES: "weakref.WeakValueDictionary[int, object]" = weakref.WeakValueDictionary()
_next = itertools.count(1)


def register_module(obj) -> int:
    h = next(_next)
    _HANDLES[h] = obj
    return h


def module_of(handle: int):
    obj = _HANDLES.get(int(handle))
    if obj is None:
        raise RuntimeError(f"vitad custom op: module handle {handle} is not alive")
    return obj


LIB = torch.library.Library("vitad", "DEF")

_SCHEMAS = {
    # encoders: (patch tokens [B,P,C] fp32, latent [B,C] fp32, fp16 GEMM operand of the GMM head [B*P,784])
    "deit_forward": "(Tensor images, int handle, int block_index) -> (Tensor, Tensor, Tensor)",
    "swin_forward": "(Tensor images, int handle) -> (Tensor, Tensor, Tensor)",
    # GMM head: per-patch mean log-likelihood L [B,P]; (prob [B,P], scores [B])
    "gmm_patch_loglik": "(Tensor x, Tensor? xaug, Tensor? gumbel, int handle, int seed, int batch_index) -> Tensor",
    "gmm_finish": "(Tensor L) -> (Tensor, Tensor)",
    # normalizing flow: (1 - p map [B,g,g], per-image loss terms [B])
    "nf_forward": "(Tensor tokens, int handle) -> (Tensor, Tensor)",
    # score maps
    "bilinear_up": "(Tensor x, int size, bool align_corners, bool pre_one_minus, bool post_one_minus, bool want_max)"
                   " -> (Tensor, Tensor)",
    "l2_map_score": "(Tensor recon, Tensor images) -> (Tensor, Tensor)",
    # reconstruction decoders: latent [B,Z] -> image [B,3,S,S]
    "decoder_forward": "(Tensor latent, int handle) -> Tensor",
    # loader-side resize: uint8 HWC -> uint8 planar
    "resize_u8": "(Tensor images_hwc, int size) -> Tensor",
}
for _name, _schema in _SCHEMAS.items():
    LIB.define(_name + _schema)


def _impl(name):
    def deco(fn):
        LIB.impl(name, fn, "CUDA")
        return fn

    return deco


def _fake(name):
    return torch.library.register_fake("vitad::" + name)


# ------------------------------------------------------------------------------------------- CUDA
@_impl("deit_forward")
def _deit_forward(images: Tensor, handle: int, block_index: int):
    return module_of(handle)._run(images, block_index)


@_impl("swin_forward")
def _swin_forward(images: Tensor, handle: int):
    return module_of(handle)._run(images)


@_impl("gmm_patch_loglik")
def _gmm_patch_loglik(x, xaug, gumbel, handle: int, seed: int, batch_index: int):
    return module_of(handle)._run(x, xaug, gumbel, seed, batch_index)


@_impl("gmm_finish")
def _gmm_finish(L: Tensor):
    from . import mdn

    return mdn._finish(L)


@_impl("nf_forward")
def _nf_forward(tokens: Tensor, handle: int):
    return module_of(handle)._run(tokens)


@_impl("bilinear_up")
def _bilinear_up(x, size: int, align_corners: bool, pre_one_minus: bool, post_one_minus: bool, want_max: bool):
    from . import ops

    return ops._bilinear_up(x, size, align_corners, pre_one_minus, post_one_minus, want_max)


@_impl("l2_map_score")
def _l2_map_score(recon: Tensor, images: Tensor):
    from . import ops

    return ops._l2_map_score(recon, images)


@_impl("decoder_forward")
def _decoder_forward(latent: Tensor, handle: int):
    return module_of(handle)._run(latent)


@_impl("resize_u8")
def _resize_u8(images_hwc: Tensor, size: int):
    from . import ops

    return ops._resize_u8(images_hwc, size)


# ------------------------------------------------------------------------------------------- fake
MDN_KA = 784


@_fake("deit_forward")
def _(images, handle, block_index):
    m = module_of(handle)
    B, P, Cd = images.shape[0], m.num_embedded_patches, m.size_patch_embedding
    return (images.new_empty((B, P, Cd), dtype=torch.float32), images.new_empty((B, Cd), dtype=torch.float32),
            images.new_empty((B * P, MDN_KA), dtype=torch.float16))


@_fake("swin_forward")
def _(images, handle):
    m = module_of(handle)
    B, P, Cd = images.shape[0], m.num_embedded_patches, m.size_patch_embedding
    return (images.new_empty((B, P, Cd), dtype=torch.float32), images.new_empty((B, Cd), dtype=torch.float32),
            images.new_empty((B * P, MDN_KA), dtype=torch.float16))


@_fake("gmm_patch_loglik")
def _(x, xaug, gumbel, handle, seed, batch_index):
    return x.new_empty((x.shape[0], x.shape[1]), dtype=torch.float32)


@_fake("gmm_finish")
def _(L):
    return L.new_empty(L.shape, dtype=torch.float32), L.new_empty((L.shape[0],), dtype=torch.float32)


@_fake("nf_forward")
def _(tokens, handle):
    m = module_of(handle)
    B = tokens.shape[0]
    return tokens.new_empty((B, m.grid, m.grid), dtype=torch.float32), tokens.new_empty((B,), dtype=torch.float32)


@_fake("bilinear_up")
def _(x, size, align_corners, pre_one_minus, post_one_minus, want_max):
    n = x.shape[0]
    return x.new_empty((n, 1, size, size), dtype=torch.float32), x.new_empty((n if want_max else 0,), dtype=torch.float32)


@_fake("l2_map_score")
def _(recon, images):
    n, _c, h, w = images.shape
    return images.new_empty((n, 1, h, w), dtype=torch.float32), images.new_empty((n,), dtype=torch.float32)


@_fake("decoder_forward")
def _(latent, handle):
    s = module_of(handle).out_size
    return latent.new_empty((latent.shape[0], 3, s, s), dtype=torch.float32)


@_fake("resize_u8")
def _(images_hwc, size):
    return images_hwc.new_empty((images_hwc.shape[0], 3, size, size), dtype=torch.uint8)


OP_NAMES = tuple(_SCHEMAS)
