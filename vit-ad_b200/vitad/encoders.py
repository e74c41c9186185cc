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

from . import _lib, custom_ops
from ._lib import check, lib


def _param_key(module: nn.Module, device) -> tuple:
    """Identity of the parameter values a packed fp16 copy was made from: device + the sum of the tensors' in-place
    version counters (optimizer.step(), copy_(), init functions bump them; load_state_dict and .to() reset the cache
    explicitly).  Writes through `.data` bypass the counter — call `module._packed = None` after such an edit."""
    v = 0
    for t in module.parameters():
        v += t._version
    for t in module.buffers():
        v += t._version
    return (device, v)


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

    def __init__(self, img=224, patch=16, dim=768, depth=12, heads=12, hidden=3072, num_classes=1000,
                 distilled=True):
        super().__init__()
        self.geometry = dict(img=img, patch=patch, dim=dim, depth=depth, heads=heads, hidden=hidden)
        n_patches = (img // patch) ** 2
        self.num_prefix = 2 if distilled else 1
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n_patches + self.num_prefix, dim))
        if distilled:
            self.dist_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.patch_embed = _PatchEmbed(dim, patch)
        self.blocks = nn.ModuleList([_Block(dim, hidden) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.head = nn.Linear(dim, num_classes)
        if distilled:
            self.head_dist = nn.Linear(dim, num_classes)
            nn.init.trunc_normal_(self.dist_token, std=0.02)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)


class _TimmVitEncoder(TransformerEncoder):
    """Shared body of EncoderDeit / EncoderVit: timm's 16/224 base transformer with 2 or 1 prefix tokens, parameters
    under the attribute the reference uses (``deit`` / ``vit``), forward = one C-ABI call (vitad_deit_forward)."""

    _ATTR = "deit"
    _DISTILLED = True

    def __init__(self, img_size: int, requires_grad: bool = False) -> None:
        super().__init__(img_size=img_size)
        if img_size != 224:
            raise ValueError(f"{type(self).__name__}: image size has to be 224 (position embedding), as in the reference")
        setattr(self, self._ATTR, _DeitParams(img=img_size, distilled=self._DISTILLED))
        self.size_patch_embedding = 768
        self.patch_size = 16
        self.num_embedded_patches = self.calc_num_embedded_patches()
        for p in self._params().parameters():
            p.requires_grad = False
        self._packed = None  # device-side fp16 copies + C structs, rebuilt when parameters change
        self._handle = custom_ops.register_module(self)

    def _params(self) -> _DeitParams:
        return getattr(self, self._ATTR)

    # -- parameter packing -------------------------------------------------------------------------
    def _apply(self, fn, recurse=True):
        self._packed = None
        return super()._apply(fn, recurse)

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._packed = None
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def _pack(self, device):
        d = self._params()
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
        w.tokens, w.prefix = d.pos_embed.shape[1], d.num_prefix
        w.patch_w = f16(d.patch_embed.proj.weight.reshape(geo["dim"], -1))
        w.patch_b = f32(d.patch_embed.proj.bias)
        prefix = (d.cls_token, d.dist_token) if d.num_prefix == 2 else (d.cls_token,)
        w.prefix_tokens = f32(torch.cat(prefix, dim=1).reshape(d.num_prefix, geo["dim"]))
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
            raise RuntimeError(f"{type(self).__name__} (vitad): CUDA input required — this implementation has no CPU path")
        block_index = self._block_index(int(block_index))
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != self.img_size or x.shape[3] != self.img_size:
            raise ValueError(f"expected [B,3,{self.img_size},{self.img_size}] input, got {tuple(x.shape)}")
        tokens, cls, xaug = torch.ops.vitad.deit_forward(x, self._handle, int(block_index))
        # fp16 GEMM operand for the MDN head (saves one conversion pass); valid while `tokens` is not modified in place
        tokens._vitad_xaug = (xaug, tokens._version)
        return TransformerEncoderOutput(patch_embedding=tokens, latent_space=cls)

    def _run(self, x: torch.Tensor, block_index: int):
        """CUDA implementation of torch.ops.vitad.deit_forward for this module's weights."""
        key = _param_key(self, x.device)
        if self._packed is None or self._packed.get("key") != key:
            self._pack(x.device)
            self._packed["key"] = key
        pk = self._packed
        # uint8 images (the dataset's native pixels) are taken as they are: the /255 of ToTensor happens in the patch
        # gather; anything else is the reference's fp32 [0,1] tensor
        u8 = x.dtype == torch.uint8
        x = x.contiguous() if u8 else x.to(torch.float32).contiguous()
        B = x.shape[0]
        P, Cdim = self.num_embedded_patches, self.size_patch_embedding
        ws = self._workspace(B, x.device)
        tokens = torch.empty((B, P, Cdim), device=x.device, dtype=torch.float32)
        cls = torch.empty((B, Cdim), device=x.device, dtype=torch.float32)
        xaug = torch.empty((B * P, _lib.MDN_KA), device=x.device, dtype=torch.float16)
        fwd = lib.vitad_deit_forward_u8 if u8 else lib.vitad_deit_forward
        check(fwd(C.byref(pk["w"]), x.data_ptr(), B, int(block_index), ws.data_ptr(), ws.numel(), tokens.data_ptr(),
                  cls.data_ptr(), xaug.data_ptr(), _lib.MDN_KA, torch.cuda.current_stream().cuda_stream))
        return tokens, cls, xaug


    def _block_index(self, block_index: int) -> int:
        return block_index


class EncoderDeit(_TimmVitEncoder):
    """Drop-in for the reference EncoderDeit (TransformerEncoder.py:116-173).

    ``requires_grad=False`` in the reference downloads ImageNet weights through timm; there is no network
    here, so both settings start from timm's random init and the caller loads a checkpoint with
    ``load_state_dict`` (keys ``deit.*`` as in the reference).  Parameters are always frozen: this class
    implements the inference/scoring path only.
    """

    _ATTR = "deit"
    _DISTILLED = True


class EncoderVit(_TimmVitEncoder):
    """Drop-in for the reference EncoderVit (TransformerEncoder.py:176-208): timm vit_base_patch16_224, one prefix
    token (197 tokens), keys ``vit.*``.  The reference's forward ignores ``block_index`` (:196-208); so does this."""

    _ATTR = "vit"
    _DISTILLED = False

    def _block_index(self, block_index: int) -> int:
        return 0


# --------------------------------------------------------------------------------------------------
# EsViT Swin-T (window 14)
# --------------------------------------------------------------------------------------------------
SWIN_DEPTHS = (2, 2, 6, 2)
SWIN_HEADS = (3, 6, 12, 24)
SWIN_EMBED = 96
SWIN_WINDOW = 14
ESVIT_CHECKPOINT = "pretrained_vit_weights/esvit-T/checkpoint_best.pth"  # TransformerEncoder.py:243-246


def _relative_position_index(ws: int) -> torch.Tensor:
    coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def interpolate_position_encoding(weights: dict, model: nn.Module) -> dict:
    """TransformerEncoder.py:276-350 (from microsoft/esvit): adapt a checkpoint whose window / image size differs
    from the model's — relative-position bias tables are resized bicubically over their (2w-1)x(2w-1) grid, the
    relative-position index buffers by nearest neighbour, an absolute position embedding bicubically over its
    token grid.  Entries with matching shapes (and everything else) pass through; a head-count mismatch is
    reported and passed through unchanged, as in the reference."""
    model_dict = model.state_dict()
    out = {}
    for k, v in weights.items():
        cur = model_dict.get(k)
        if cur is not None and v.size() != cur.size():
            if "relative_position_bias_table" in k:
                (l1, nh1), (l2, nh2) = v.size(), cur.size()
                if nh1 != nh2:
                    print(f"Error in loading {k}, passing")
                elif l1 != l2:
                    s1, s2 = int(l1 ** 0.5), int(l2 ** 0.5)
                    r = torch.nn.functional.interpolate(v.permute(1, 0).view(1, nh1, s1, s1), size=(s2, s2), mode="bicubic")
                    v = r.view(nh2, l2).permute(1, 0)
            elif "relative_position_index" in k:
                (l1, h1), (l2, h2) = v.size(), cur.size()
                v = torch.nn.functional.interpolate(v.view(1, 1, l1, h1).float(), size=(l2, h2), mode="nearest").view(l2, h2)
            elif "absolute_pos_embed" in k:
                (_, l1, c1), (_, l2, c2) = v.size(), cur.size()
                if c1 != c2:
                    print(f"Error in loading {k}, passing")
                elif l1 != l2:
                    s1, s2 = int(l1 ** 0.5), int(l2 ** 0.5)
                    r = torch.nn.functional.interpolate(v.reshape(-1, s1, s1, c1).permute(0, 3, 1, 2), size=(s2, s2), mode="bicubic")
                    v = r.permute(0, 2, 3, 1).flatten(1, 2)
        out[k] = v
    return out


class _WinAttn(nn.Module):
    def __init__(self, dim, ws, heads):
        super().__init__()
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) ** 2, heads))
        self.register_buffer("relative_position_index", _relative_position_index(ws))
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)


class _SwinBlock(nn.Module):
    def __init__(self, dim, res, heads, window, shift):
        super().__init__()
        if res <= window:  # SwinTransformerModule.py:262-265
            shift, window = 0, res
        self.window_size, self.shift_size = window, shift
        self.norm1 = nn.LayerNorm(dim)
        self.attn = _WinAttn(dim, window, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _Mlp(dim, 4 * dim)


class _PatchMerging(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = nn.LayerNorm(4 * dim)


class _SwinLayer(nn.Module):
    def __init__(self, dim, res, depth, heads, window, downsample):
        super().__init__()
        self.blocks = nn.ModuleList(
            [_SwinBlock(dim, res, heads, window, 0 if i % 2 == 0 else window // 2) for i in range(depth)])
        if downsample:
            self.downsample = _PatchMerging(dim)
        else:
            self.downsample = None


class _SwinPatchEmbed(nn.Module):
    def __init__(self, embed):
        super().__init__()
        self.proj = nn.Conv2d(3, embed, kernel_size=4, stride=4)
        self.norm = nn.LayerNorm(embed)


class _SwinParams(nn.Module):
    """Parameter tree of the vendored SwinTransformer as EncoderEsVit builds it (185 state_dict keys)."""

    def __init__(self, img=224, num_classes=3):
        super().__init__()
        self.img = img
        self.patch_embed = _SwinPatchEmbed(SWIN_EMBED)
        res = img // 4
        self.layers = nn.ModuleList()
        for s, (depth, heads) in enumerate(zip(SWIN_DEPTHS, SWIN_HEADS)):
            self.layers.append(_SwinLayer(SWIN_EMBED * 2**s, res, depth, heads, SWIN_WINDOW, s < 3))
            if s < 3:
                res //= 2
        self.norm = nn.LayerNorm(SWIN_EMBED * 8)
        self.head = nn.Linear(SWIN_EMBED * 8, num_classes)
        for m in self.modules():  # SwinTransformer._init_weights (:800-808)
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)


def _window_maps(H: int, ws: int, shift: int):
    """token -> window*T + pos for the (cyclically shifted) window partition, its inverse, and the region label
    of every window position (create_attn_mask, SwinTransformerModule.py:316-347)."""
    T = ws * ws
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(H), indexing="ij")
    y2, x2 = (ys - shift) % H, (xs - shift) % H  # torch.roll(x, -shift): token y lands at (y - shift) mod H
    tok2win = ((y2 // ws) * (H // ws) + x2 // ws) * T + (y2 % ws) * ws + x2 % ws
    tok2win = tok2win.reshape(-1).to(torch.int32)
    win2tok = torch.empty_like(tok2win)
    win2tok[tok2win.long()] = torch.arange(H * H, dtype=torch.int32)
    region = None
    if shift > 0:
        def band(v):
            return (v >= H - ws).long() + (v >= H - shift).long()
        ids = (3 * band(y2) + band(x2)).reshape(-1)  # label per token (in shifted-frame coordinates)
        region = torch.empty(H * H, dtype=torch.int8)
        region[tok2win.long()] = ids.to(torch.int8)
    return tok2win, win2tok, region


class EncoderEsVit(TransformerEncoder):
    """Drop-in for the reference EncoderEsVit (TransformerEncoder.py:211-273): vendored Swin-T, window 14.

    `requires_grad=False` loads the EsViT student checkpoint from the reference's relative path when that file
    exists, through `interpolate_position_encoding` (:276-350) like the reference; without the file the random
    init is kept and a notice is printed (the reference would raise).  Forward runs
    in inference mode: the reference leaves DropPath(0.1) active during MDN/NF validation, which makes its
    features random; that accident is deliberately not reproduced.
    """

    def __init__(self, img_size: int, requires_grad: bool = False) -> None:
        super().__init__(img_size=img_size)
        if img_size != 224:
            raise ValueError("EncoderEsVit (vitad): image size has to be 224")
        self.size_patch_embedding = 768
        self.patch_size = 32
        self.num_embedded_patches = self.calc_num_embedded_patches()
        self.esvit = _SwinParams(img=img_size)
        if not requires_grad:
            import os

            if os.path.exists(ESVIT_CHECKPOINT):
                student = torch.load(ESVIT_CHECKPOINT, map_location="cpu")["student"]
                weights = {k[7:]: v for k, v in student.items() if not k.startswith("module.head")}
                delattr(self.esvit, "head")
                self.load_student_weights(weights)
            else:
                print(f"EncoderEsVit (vitad): {ESVIT_CHECKPOINT} not found, keeping the random initialisation")
        for p in self.esvit.parameters():
            p.requires_grad = False
        self._packed = None
        self._handle = custom_ops.register_module(self)

    def _apply(self, fn, recurse=True):
        self._packed = None
        return super()._apply(fn, recurse)

    def load_student_weights(self, weights: dict) -> None:
        """The reference's checkpoint path (TransformerEncoder.py:248-263): `student` weights with the `module.`
        prefix stripped and the head dropped, adapted by interpolate_position_encoding, loaded into `self.esvit`."""
        adapted = interpolate_position_encoding(weights=weights, model=self.esvit)
        fixed = {}
        for k, v in adapted.items():  # the nearest-neighbour resize yields float indices; buffers are int64
            fixed[k] = v.long() if "relative_position_index" in k else v
        self.esvit.load_state_dict(fixed)
        self._packed = None

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._packed = None
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def _pack(self, device):
        e = self.esvit
        keep = []

        def dev(t, dtype):
            t = t.detach().to(device=device, dtype=dtype).contiguous()
            keep.append(t)
            return t.data_ptr()

        f32 = lambda t: dev(t, torch.float32)
        f16 = lambda t: dev(t, torch.float16)
        stages = (_lib.SwinStage * len(e.layers))()
        block_arrays = []
        res = self.img_size // 4
        for s, layer in enumerate(e.layers):
            st = stages[s]
            dim = SWIN_EMBED * 2**s
            ws = layer.blocks[0].window_size
            st.dim, st.heads, st.res, st.window, st.depth = dim, SWIN_HEADS[s], res, ws, len(layer.blocks)
            T = ws * ws
            blocks = (_lib.SwinBlock * len(layer.blocks))()
            block_arrays.append(blocks)
            for i, blk in enumerate(layer.blocks):
                B = blocks[i]
                B.ln1_w, B.ln1_b = f32(blk.norm1.weight), f32(blk.norm1.bias)
                B.qkv_w, B.qkv_b = f16(blk.attn.qkv.weight), f32(blk.attn.qkv.bias)
                table, index = blk.attn.relative_position_bias_table.detach().float(), blk.attn.relative_position_index
                # dense bias[h][key][query] (key-major: the kernel's query lanes read contiguous memory)
                B.attn_bias = f32(table[index.reshape(-1).long()].view(T, T, -1).permute(2, 1, 0))
                B.proj_w, B.proj_b = f16(blk.attn.proj.weight), f32(blk.attn.proj.bias)
                B.ln2_w, B.ln2_b = f32(blk.norm2.weight), f32(blk.norm2.bias)
                B.fc1_w, B.fc1_b = f16(blk.mlp.fc1.weight), f32(blk.mlp.fc1.bias)
                B.fc2_w, B.fc2_b = f16(blk.mlp.fc2.weight), f32(blk.mlp.fc2.bias)
                B.shift = blk.shift_size
            st.blocks = C.cast(blocks, C.POINTER(_lib.SwinBlock))
            if res > ws:
                t0, w0, _ = _window_maps(res, ws, 0)
                st.tok2win[0], st.win2tok[0] = dev(t0, torch.int32), dev(w0, torch.int32)
                shift = max(b.shift_size for b in layer.blocks)
                if shift > 0:
                    t1, w1, reg = _window_maps(res, ws, shift)
                    st.tok2win[1], st.win2tok[1] = dev(t1, torch.int32), dev(w1, torch.int32)
                    st.region = dev(reg, torch.int8)
            if layer.downsample is not None:
                st.merge_ln_w, st.merge_ln_b = f32(layer.downsample.norm.weight), f32(layer.downsample.norm.bias)
                st.merge_w = f16(layer.downsample.reduction.weight)
                res //= 2
        w = _lib.SwinWeights()
        w.img, w.patch, w.embed, w.stages = self.img_size, 4, SWIN_EMBED, len(e.layers)
        w.patch_w = f16(e.patch_embed.proj.weight.reshape(SWIN_EMBED, -1))
        w.patch_b = f32(e.patch_embed.proj.bias)
        w.patch_ln_w, w.patch_ln_b = f32(e.patch_embed.norm.weight), f32(e.patch_embed.norm.bias)
        w.norm_w, w.norm_b = f32(e.norm.weight), f32(e.norm.bias)
        w.stage = C.cast(stages, C.POINTER(_lib.SwinStage))
        self._packed = dict(w=w, stages=stages, blocks=block_arrays, keep=keep, device=device, ws=None, ws_batch=0)

    def forward(self, x: torch.Tensor, block_index: int = 0) -> TransformerEncoderOutput:
        if not x.is_cuda:
            raise RuntimeError("EncoderEsVit (vitad): CUDA input required — this implementation has no CPU path")
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != self.img_size or x.shape[3] != self.img_size:
            raise ValueError(f"expected [B,3,{self.img_size},{self.img_size}] input, got {tuple(x.shape)}")
        tokens, latent, xaug = torch.ops.vitad.swin_forward(x, self._handle)
        tokens._vitad_xaug = (xaug, tokens._version)
        return TransformerEncoderOutput(latent_space=latent, patch_embedding=tokens)

    def _run(self, x: torch.Tensor):
        """CUDA implementation of torch.ops.vitad.swin_forward for this module's weights."""
        key = _param_key(self, x.device)
        if self._packed is None or self._packed.get("key") != key:
            self._pack(x.device)
            self._packed["key"] = key
        pk = self._packed
        if x.dtype == torch.uint8:  # native pixels: the /255 of ToTensor (GeneralDataset.py:46-53)
            x = x.to(torch.float32).div_(255.0)
        x = x.to(torch.float32).contiguous()
        B = x.shape[0]
        P, Cdim = self.num_embedded_patches, self.size_patch_embedding
        if pk["ws"] is None or pk["ws_batch"] < B:
            nbytes = lib.vitad_swin_workspace_bytes(C.byref(pk["w"]), B)
            pk["ws"], pk["ws_batch"] = torch.empty(nbytes, device=x.device, dtype=torch.uint8), B
        tokens = torch.empty((B, P, Cdim), device=x.device, dtype=torch.float32)
        latent = torch.empty((B, Cdim), device=x.device, dtype=torch.float32)
        xaug = torch.empty((B * P, _lib.MDN_KA), device=x.device, dtype=torch.float16)
        check(lib.vitad_swin_forward(C.byref(pk["w"]), x.data_ptr(), B, pk["ws"].data_ptr(), pk["ws"].numel(),
                                     tokens.data_ptr(), latent.data_ptr(), xaug.data_ptr(), _lib.MDN_KA,
                                     torch.cuda.current_stream().cuda_stream))
        return tokens, latent, xaug
