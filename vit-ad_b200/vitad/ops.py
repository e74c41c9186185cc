"""Tensor-level wrappers over the C ABI: torch is used for device memory and streams only."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
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
    epilogue: int = _lib.EPI_BIAS_BF16,
    out: torch.Tensor | None = None,
    resid: torch.Tensor | None = None,
    block_n: int = 0,
) -> torch.Tensor:
    """out = epilogue(a @ w.T + bias).  a: bf16 [M,K]; w: bf16 [N,K]; bias: fp32 [N]."""
    _need_cuda(a, w, bias, out, resid)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert a.dim() == 2 and w.dim() == 2 and a.stride(1) == 1 and w.stride(1) == 1
    m, k = a.shape
    n = w.shape[0]
    if out is None:
        dt = torch.float32 if epilogue in (_lib.EPI_RESIDUAL_F32, _lib.EPI_F32) else torch.bfloat16
        out = torch.empty((m, n), device=a.device, dtype=dt)
    args = _lib.LinearArgs()
    args.a, args.w, args.bias = a.data_ptr(), w.data_ptr(), _ptr(bias)
    args.m, args.n, args.k = m, n, k
    args.lda, args.ldw = a.stride(0), w.stride(0)
    args.epilogue, args.block_n = epilogue, block_n
    args.out, args.ldo = out.data_ptr(), out.stride(0)
    args.resid = _ptr(resid)
    check(lib.vitad_linear_bf16(C.byref(args), _stream()))
    return out


def linear_qkv(
    a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, batch: int, tokens: int, heads: int, tokens_pad: int,
    q: torch.Tensor, k: torch.Tensor, vt: torch.Tensor, q_scale: float, block_n: int = 0,
) -> None:
    """Fused qkv projection writing q/k as [B,H,T,64] and v transposed as [B,H,64,Tpad] (bf16)."""
    _need_cuda(a, w, bias, q, k, vt)
    m, kk = a.shape
    assert m == batch * tokens
    args = _lib.LinearArgs()
    args.a, args.w, args.bias = a.data_ptr(), w.data_ptr(), bias.data_ptr()
    args.m, args.n, args.k = m, w.shape[0], kk
    args.lda, args.ldw = a.stride(0), w.stride(0)
    args.epilogue, args.block_n = _lib.EPI_QKV, block_n
    args.q, args.kmat, args.vt = q.data_ptr(), k.data_ptr(), vt.data_ptr()
    args.tokens, args.tokens_pad, args.heads, args.q_scale = tokens, tokens_pad, heads, q_scale
    check(lib.vitad_linear_bf16(C.byref(args), _stream()))


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
    check(lib.vitad_linear_bf16(C.byref(args), _stream()))
