"""Tensor-level wrappers over the C ABI: torch is used for device memory and streams only."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from . import custom_ops  # noqa: F401  (registers torch.ops.vitad.*)
from ._lib import check, lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("vitad ops run on CUDA tensors only (no CPU path)")


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def linear(
    a: torch.Tensor,
    w: torch.Tensor,
    bias: torch.Tensor | None,
    epilogue: int = _lib.EPI_BIAS_F16,
    out: torch.Tensor | None = None,
    resid: torch.Tensor | None = None,
    block_n: int = 0,
) -> torch.Tensor:
    """out = epilogue(a @ w.T + bias).  a: fp16 [M,K]; w: fp16 [N,K]; bias: fp32 [N]."""
    _need_cuda(a, w, bias, out, resid)
    assert a.dtype == torch.float16 and w.dtype == torch.float16
    assert a.dim() == 2 and w.dim() == 2 and a.stride(1) == 1 and w.stride(1) == 1
    m, k = a.shape
    n = w.shape[0]
    if out is None:
        dt = torch.float32 if epilogue in (_lib.EPI_RESIDUAL_F32, _lib.EPI_F32) else torch.float16
        out = torch.empty((m, n), device=a.device, dtype=dt)
    args = _lib.LinearArgs()
    args.a, args.w, args.bias = a.data_ptr(), w.data_ptr(), _ptr(bias)
    args.m, args.n, args.k = m, n, k
    args.lda, args.ldw = a.stride(0), w.stride(0)
    args.epilogue, args.block_n = epilogue, block_n
    args.out, args.ldo = out.data_ptr(), out.stride(0)
    args.resid = _ptr(resid)
    check(lib.vitad_linear_f16(C.byref(args), _stream()))
    return out


def pad_pixels(x: torch.Tensor, batch: int, grid: int) -> torch.Tensor:
    """Plain NHWC pixel rows fp16 [B*g*g, C] → the zero-bordered layout [B*(g+2)*(g+2), C] the implicit convolution reads."""
    c = x.shape[1]
    out = torch.zeros((batch, grid + 2, grid + 2, c), device=x.device, dtype=x.dtype)
    out[:, 1:-1, 1:-1, :] = x.view(batch, grid, grid, c)
    return out.view(-1, c)


def conv3x3(xp: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, batch: int, grid: int, relu: bool = False,
            out: torch.Tensor | None = None) -> torch.Tensor:
    """3x3 convolution (stride 1, zero padding 1) as one tcgen05 GEMM without an im2col pass.  xp: zero-bordered NHWC
    fp16 [B*(g+2)*(g+2), C] (C % 64 == 0); w: fp16 [N, 9*C], columns (ty, tx, c) = kernel element [n, c, ty, tx] of an
    nn.Conv2d; returns plain pixel rows fp16 [B*g*g, N]."""
    _need_cuda(xp, w, bias, out)
    assert xp.dtype == torch.float16 and w.dtype == torch.float16 and xp.stride(1) == 1 and w.stride(1) == 1
    c = xp.shape[1]
    n = w.shape[0]
    assert xp.shape[0] == batch * (grid + 2) ** 2 and w.shape[1] == 9 * c
    if out is None:
        out = torch.empty((batch * grid * grid, n), device=xp.device, dtype=torch.float16)
    args = _lib.LinearArgs()
    args.a, args.w, args.bias = xp.data_ptr(), w.data_ptr(), bias.data_ptr()
    args.m, args.n, args.k = xp.shape[0], n, 9 * c
    args.lda, args.ldw = xp.stride(0), w.stride(0)
    args.epilogue = _lib.EPI_BIAS_RELU_F16 if relu else _lib.EPI_BIAS_F16
    args.out, args.ldo = out.data_ptr(), out.stride(0)
    args.conv_grid = grid
    check(lib.vitad_linear_f16(C.byref(args), _stream()))
    return out


def resize_plan(in_size: int, out_size: int) -> torch.Tensor:
    """Host plan of one axis of Pillow's BILINEAR resize (int32: first source pixel | tap count | coefficients)."""
    ks = lib.vitad_resize_ksize(in_size, out_size)
    plan = torch.empty(2 * out_size + out_size * ks, dtype=torch.int32)
    check(lib.vitad_resize_plan(in_size, out_size, plan.data_ptr()))
    return plan


_resize_plans: dict = {}


def resize_u8(images_hwc: torch.Tensor, size: int) -> torch.Tensor:
    """transforms.Resize((size, size)) of the reference's loader (GeneralDataset.py:38-59) on the device, bit-identical
    to Pillow: uint8 [B, H, W, 3] (decoded images, HWC) → uint8 [B, 3, size, size] (planar; feed it to the encoders'
    uint8 path, which folds ToTensor's /255 into the patch gather).  = torch.ops.vitad.resize_u8."""
    _need_cuda(images_hwc)
    return torch.ops.vitad.resize_u8(images_hwc, size)


def _resize_u8(images_hwc: torch.Tensor, size: int) -> torch.Tensor:
    assert images_hwc.dtype == torch.uint8 and images_hwc.dim() == 4 and images_hwc.shape[3] == 3
    images_hwc = images_hwc.contiguous()
    b, h, w, _ = images_hwc.shape
    key = (h, w, size, images_hwc.device)
    if key not in _resize_plans:
        _resize_plans[key] = (resize_plan(w, size).to(images_hwc.device), resize_plan(h, size).to(images_hwc.device))
    plan_h, plan_v = _resize_plans[key]
    tmp = torch.empty((b, h, size, 3), device=images_hwc.device, dtype=torch.uint8)
    out = torch.empty((b, 3, size, size), device=images_hwc.device, dtype=torch.uint8)
    check(lib.vitad_resize_bilinear_u8(images_hwc.data_ptr(), b, h, w, size, plan_h.data_ptr(), plan_v.data_ptr(),
                                       tmp.data_ptr(), out.data_ptr(), _stream()))
    return out


def prepare_images(images: torch.Tensor, size: int) -> torch.Tensor:
    """Device-side half of the loader's image transform: a uint8 HWC batch of decoded images [B, H, W, 3] (any H, W) is
    resized like transforms.Resize((size, size)) does on the CPU and returned as uint8 NCHW; NCHW batches (fp32 in
    [0,1] or uint8) pass through."""
    if images.dtype == torch.uint8 and images.dim() == 4 and images.shape[-1] == 3 and images.shape[1] != 3:
        return resize_u8(images, size)
    return images


def linear_qkv(
    a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, batch: int, tokens: int, heads: int, tokens_pad: int,
    q: torch.Tensor, k: torch.Tensor, vt: torch.Tensor, q_scale: float, block_n: int = 0, head_dim: int = 0,
    windows: int = 0, win_tokens: int = 0, tok2win: torch.Tensor | None = None, v_natural: bool = False,
) -> None:
    """Fused qkv projection writing q/k as [BW,H,T,hd] and v transposed as [BW,H,hd,Tpad] (fp16); with
    `windows`/`tok2win` the rows are scattered into (shifted) attention windows."""
    _need_cuda(a, w, bias, q, k, vt, tok2win)
    m, kk = a.shape
    assert m == batch * tokens
    args = _lib.LinearArgs()
    args.a, args.w, args.bias = a.data_ptr(), w.data_ptr(), bias.data_ptr()
    args.m, args.n, args.k = m, w.shape[0], kk
    args.lda, args.ldw = a.stride(0), w.stride(0)
    args.epilogue, args.block_n = _lib.EPI_QKV, block_n
    args.q, args.kmat, args.vt = q.data_ptr(), k.data_ptr(), vt.data_ptr()
    args.tokens, args.tokens_pad, args.heads, args.q_scale = tokens, tokens_pad, heads, q_scale
    args.head_dim, args.windows, args.win_tokens, args.tok2win = head_dim, windows, win_tokens, _ptr(tok2win)
    args.v_natural = int(v_natural)  # vt then receives V as [BW,H,T,hd] (the layout of k)
    check(lib.vitad_linear_f16(C.byref(args), _stream()))


def linear_patch_embed(
    a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, pos: torch.Tensor, out: torch.Tensor, patches: int,
    prefix: int, block_n: int = 0,
) -> None:
    """x[b, prefix+p, :] = a[b*P+p] @ w.T + bias + pos[prefix+p]  (fp32 residual stream)."""
    _need_cuda(a, w, bias, pos, out)
    args = _lib.LinearArgs()
    args.a, args.w, args.bias = a.data_ptr(), w.data_ptr(), bias.data_ptr()
    args.m, args.n, args.k = a.shape[0], w.shape[0], a.shape[1]
    args.lda, args.ldw = a.stride(0), w.stride(0)
    args.epilogue, args.block_n = _lib.EPI_PATCH_EMBED, block_n
    args.out, args.ldo = out.data_ptr(), w.shape[0]
    args.pos, args.patches, args.prefix = pos.data_ptr(), patches, prefix
    check(lib.vitad_linear_f16(C.byref(args), _stream()))


# ------------------------------------------------------------------------------- encoder kernels
def layernorm(x, weight, bias, eps, out_f16=None, out_f32=None, in_tokens=None, out_tokens=None, skip=0, aug_ones=0):
    """LayerNorm over the last dim of fp32 x [rows_in, C]; see include/vitad.h for the row remap."""
    _need_cuda(x, weight, bias, out_f16, out_f32)
    rows_in, c = x.shape
    if in_tokens is None:
        in_tokens = out_tokens = rows_in
    rows = (rows_in // in_tokens) * out_tokens
    check(lib.vitad_layernorm(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), _ptr(out_f16), _ptr(out_f32), rows, c,
                              x.stride(0), out_f16.stride(0) if out_f16 is not None else 0,
                              out_f32.stride(0) if out_f32 is not None else 0, in_tokens, out_tokens, skip, eps,
                              aug_ones, _stream()))


def patchify(images, patch):
    _need_cuda(images)
    b, c, s, _ = images.shape
    g = s // patch
    out = torch.empty((b * g * g, c * patch * patch), device=images.device, dtype=torch.float16)
    check(lib.vitad_patchify(images.data_ptr(), out.data_ptr(), b, c, s, patch, _stream()))
    return out


def attention(q, k, vt, tokens, windows=1, bias=None, region=None, win2tok=None, v=None):
    """q,k fp16 [BW,H,T,hd] (q pre-scaled); vt fp16 [BW,H,hd,Tpad] zero-padded -> fp16 [BW*T, H*hd] in original
    token order.  Swin: bias fp32 [H,T_key,T_query] (key-major), region int8 [windows,T], win2tok int32 [windows*T].
    Instead of vt: v fp16 [BW,H,T,64] in its natural layout (the MN-major operand path of the kernel)."""
    _need_cuda(q, k, vt, bias, region, win2tok, v)
    bw, h, t, hd = q.shape
    out = torch.empty((bw * t, h * hd), device=q.device, dtype=torch.float16)
    a = _lib.AttentionArgs()
    a.q, a.k, a.vt, a.out, a.v = q.data_ptr(), k.data_ptr(), _ptr(vt), out.data_ptr(), _ptr(v)
    a.batch_windows, a.heads, a.tokens, a.head_dim, a.windows = bw, h, tokens, hd, windows
    a.tokens_pad = vt.shape[-1] if vt is not None else 0
    a.bias, a.region, a.win2tok = _ptr(bias), _ptr(region), _ptr(win2tok)
    check(lib.vitad_attention_f16(C.byref(a), _stream()))
    return out


# ------------------------------------------------------------------------------------- score maps
def bilinear_up(x, size, align_corners, pre_one_minus=False, post_one_minus=False, want_max=False):
    """x fp32 [N,g,g] -> ([N,1,size,size], per-image max or None).  = torch.ops.vitad.bilinear_up."""
    _need_cuda(x)
    out, mx = torch.ops.vitad.bilinear_up(x, int(size), bool(align_corners), bool(pre_one_minus), bool(post_one_minus),
                                          bool(want_max))
    return out, (mx if want_max else None)


def _bilinear_up(x, size, align_corners, pre_one_minus, post_one_minus, want_max):
    x = x.to(torch.float32).contiguous()
    n, g, _ = x.shape
    out = torch.empty((n, 1, size, size), device=x.device, dtype=torch.float32)
    mx = torch.empty((n,), device=x.device, dtype=torch.float32) if want_max else None
    check(lib.vitad_bilinear_up(x.data_ptr(), out.data_ptr(), _ptr(mx), n, g, size, int(align_corners),
                                int(pre_one_minus), int(post_one_minus), _stream()))
    return out, (mx if mx is not None else torch.empty((0,), device=x.device, dtype=torch.float32))


def l2_map_score(recon, x):
    """mean_c (recon - x)^2 -> ([N,1,H,W] map, [N] per-image max).  = torch.ops.vitad.l2_map_score."""
    _need_cuda(recon, x)
    return torch.ops.vitad.l2_map_score(recon, x)


def _l2_map_score(recon, x):
    n, c, h, w = x.shape
    recon, x = recon.contiguous(), x.contiguous()
    amap = torch.empty((n, 1, h, w), device=x.device, dtype=torch.float32)
    mx = torch.empty((n,), device=x.device, dtype=torch.float32)
    check(lib.vitad_l2_map_score(recon.data_ptr(), x.data_ptr(), amap.data_ptr(), mx.data_ptr(), n, c, h * w,
                                 _stream()))
    return amap, mx
