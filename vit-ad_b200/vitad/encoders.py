"""Encoders of the scoring path under the reference's class names
(src/classes/transformer/TransformerEncoder.py).  The nn.Modules below hold the parameters with the
reference's ``state_dict`` key layout (timm names) so reference checkpoints load unchanged; the forward
pass is one C-ABI call into the CUDA library.  There is no torch fallback: non-CUDA input raises.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import torch
from torch import nn

from . import _lib
from ._lib import check, lib


@dataclass
class TransformerEncoderOutput:
    """Same fields as the reference dataclass (TransformerEncoder.py:15-20)."""

    patch_embedding: torch.Tensor
    latent_space: torch.Tensor | None = None


class TransformerEncoder(nn.Module):
    """Base class (TransformerEncoder.py:23-43): attributes read by the validators."""

    def __init__(self, img_size: int) -> None:
        super().__init__()
        self.img_size = img_size
        self.architecture = "transformer_encoder"
        self.size_patch_embedding = 0
        self.patch_size = 1
        self.num_embedded_patches = 0

    def calc_num_embedded_patches(self):
        return int((self.img_size / self.patch_size) ** 2)


# --------------------------------------------------------------------------------------------------
# Parameter containers with timm's attribute tree (never called; they only own the tensors).
# --------------------------------------------------------------------------------------------------
class _Attn(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attn(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, hidden)


class _PatchEmbed(nn.Module):
    def __init__(self, dim, patch):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)


class _DeitParams(nn.Module):
    """Parameters of timm's deit_base_distilled_patch16_224 with its random init (pretrained=False)."""

    def __init__(self, img=224, patch=16, dim=768, depth=12, heads=12, hidden=3072, num_classes=1000):
        super().__init__()
        self.geometry = dict(img=img, patch=patch, dim=dim, depth=depth, heads=heads, hidden=hidden)
        n_patches = (img // patch) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n_patches + 2, dim))
        self.dist_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.patch_embed = _PatchEmbed(dim, patch)
        self.blocks = nn.ModuleList([_Block(dim, hidden) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.head = nn.Linear(dim, num_classes)
        self.head_dist = nn.Linear(dim, num_classes)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.dist_token, std=0.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)


class EncoderDeit(TransformerEncoder):
    """Drop-in for the reference EncoderDeit (TransformerEncoder.py:116-173).

    ``requires_grad=False`` in the reference downloads ImageNet weights through timm; there is no network
    here, so both settings start from timm's random init and the caller loads a checkpoint with
    ``load_state_dict`` (keys ``deit.*`` as in the reference).  Parameters are always frozen: this class
    implements the inference/scoring path only.
    """

    def __init__(self, img_size: int, requires_grad: bool = False) -> None:
        super().__init__(img_size=img_size)
        if img_size != 224:
            raise ValueError("EncoderDeit: image size has to be 224 (position embedding), as in the reference")
        self.deit = _DeitParams(img=img_size)
        self.size_patch_embedding = 768
        self.patch_size = 16
        self.num_embedded_patches = self.calc_num_embedded_patches()
        for p in self.deit.parameters():
            p.requires_grad = False
        self._packed = None  # device-side fp16 copies + C structs, rebuilt when parameters change

    # -- parameter packing -------------------------------------------------------------------------
    def _apply(self, fn, recurse=True):
        self._packed = None
        return super()._apply(fn, recurse)

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._packed = None
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def _pack(self, device):
        d = self.deit
        geo = d.geometry
        keep = []  # keep device tensors alive

        def f32(t):
            t = t.detach().to(device=device, dtype=torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()

        def f16(t):
            t = t.detach().to(device=device, dtype=torch.float16).contiguous()
            keep.append(t)
            return t.data_ptr()

        layers = (_lib.DeitLayer * geo["depth"])()
        for i, b in enumerate(d.blocks):
            L = layers[i]
            L.ln1_w, L.ln1_b = f32(b.norm1.weight), f32(b.norm1.bias)
            L.qkv_w, L.qkv_b = f16(b.attn.qkv.weight), f32(b.attn.qkv.bias)
            L.proj_w, L.proj_b = f16(b.attn.proj.weight), f32(b.attn.proj.bias)
            L.ln2_w, L.ln2_b = f32(b.norm2.weight), f32(b.norm2.bias)
            L.fc1_w, L.fc1_b = f16(b.mlp.fc1.weight), f32(b.mlp.fc1.bias)
            L.fc2_w, L.fc2_b = f16(b.mlp.fc2.weight), f32(b.mlp.fc2.bias)
        w = _lib.DeitWeights()
        w.img, w.patch, w.dim, w.heads = geo["img"], geo["patch"], geo["dim"], geo["heads"]
        w.hidden, w.depth = geo["hidden"], geo["depth"]
        w.tokens, w.prefix = d.pos_embed.shape[1], 2
        w.patch_w = f16(d.patch_embed.proj.weight.reshape(geo["dim"], -1))
        w.patch_b = f32(d.patch_embed.proj.bias)
        w.prefix_tokens = f32(torch.cat((d.cls_token, d.dist_token), dim=1).reshape(2, geo["dim"]))
        w.pos = f32(d.pos_embed.reshape(-1, geo["dim"]))
        w.norm_w, w.norm_b = f32(d.norm.weight), f32(d.norm.bias)
        w.layers = C.cast(layers, C.POINTER(_lib.DeitLayer))
        self._packed = dict(w=w, layers=layers, keep=keep, device=device, ws=None, ws_batch=0)

    def _workspace(self, batch, device):
        pk = self._packed
        if pk["ws"] is None or pk["ws_batch"] < batch:
            nbytes = lib.vitad_deit_workspace_bytes(C.byref(pk["w"]), batch)
            pk["ws"] = torch.empty(nbytes, device=device, dtype=torch.uint8)
            pk["ws_batch"] = batch
        return pk["ws"]

    # -- forward -----------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, block_index: int = 0) -> TransformerEncoderOutput:
        if not x.is_cuda:
            raise RuntimeError("EncoderDeit (vitad): CUDA input required — this implementation has no CPU path")
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != self.img_size or x.shape[3] != self.img_size:
            raise ValueError(f"expected [B,3,{self.img_size},{self.img_size}] input, got {tuple(x.shape)}")
        if self._packed is None or self._packed["device"] != x.device:
            self._pack(x.device)
        pk = self._packed
        x = x.to(torch.float32).contiguous()
        B = x.shape[0]
        P, Cdim = self.num_embedded_patches, self.size_patch_embedding
        ws = self._workspace(B, x.device)
        tokens = torch.empty((B, P, Cdim), device=x.device, dtype=torch.float32)
        cls = torch.empty((B, Cdim), device=x.device, dtype=torch.float32)
        xaug = torch.empty((B * P, _lib.MDN_KA), device=x.device, dtype=torch.float16)
        check(lib.vitad_deit_forward(C.byref(pk["w"]), x.data_ptr(), B, int(block_index), ws.data_ptr(), ws.numel(),
                                     tokens.data_ptr(), cls.data_ptr(), xaug.data_ptr(), _lib.MDN_KA,
                                     torch.cuda.current_stream().cuda_stream))
        tokens._vitad_xaug = xaug  # fp16 GEMM operand for the MDN head (saves one conversion pass)
        return TransformerEncoderOutput(patch_embedding=tokens, latent_space=cls)
