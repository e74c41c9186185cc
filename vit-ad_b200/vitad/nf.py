"""Normalizing-flow head under the reference's names (src/classes/NormalizingFlow.py).

`NormalizingFlow` keeps the reference's constructor and `state_dict` layout — `layer_norm.*` (constructed but
never applied, :43-45,124-125) and `fast_flow_decoder.module_list.{i}.{global_scale, global_offset, w_perm,
w_perm_inv, subnet.0.*, subnet.2.*}` as FrEIA's AllInOneBlock registers them — and runs the whole flow as one
C-ABI call.  The 0/1 permutation matrices are turned into index vectors at pack time.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np
import torch
from torch import nn

from . import _lib, custom_ops, ops
from ._lib import check, lib

HIDDEN_PAD = 64
TILE_HALF = 48


@dataclass
class NormalizingFlowReturn:
    """Same fields as NormalizingFlow.py:14-19."""

    loss: torch.Tensor
    anomaly_score_map: torch.Tensor


class _Step(nn.Module):
    """Parameter container of one AllInOneBlock (FrEIA 0.2 names and default initialisation)."""

    def __init__(self, channels: int, hidden: int, ksize: int):
        super().__init__()
        c2 = channels // 2
        c1 = channels - c2
        gs0 = float(2.0 * np.log(np.exp(0.5 * 10.0 * 1.0) - 1))
        self.global_scale = nn.Parameter(torch.ones(1, channels, 1, 1) * gs0)
        self.global_offset = nn.Parameter(torch.zeros(1, channels, 1, 1))
        w = np.zeros((channels, channels), dtype=np.float32)
        for i, j in enumerate(np.random.permutation(channels)):
            w[i, j] = 1.0
        self.w_perm = nn.Parameter(torch.from_numpy(w).view(channels, channels, 1, 1), requires_grad=False)
        self.w_perm_inv = nn.Parameter(torch.from_numpy(w.T.copy()).view(channels, channels, 1, 1), requires_grad=False)
        self.subnet = nn.Sequential(
            nn.Conv2d(c1, hidden, kernel_size=ksize, padding="same"),
            nn.ReLU(inplace=False),
            nn.Conv2d(hidden, 2 * c2, kernel_size=ksize, padding="same"),
        )


class _Sequence(nn.Module):
    def __init__(self):
        super().__init__()
        self.module_list = nn.ModuleList()


class NormalizingFlow(nn.Module):
    """Drop-in for NormalizingFlow.py:22-145."""

    def __init__(self, num_channels: int, img_size: int, num_patches: int, hidden_ratio: float = 1.0,
                 flow_steps: int = 8) -> None:
        super().__init__()
        if num_channels != 768:
            raise ValueError("vitad NormalizingFlow supports 768-channel features (DeiT / EsViT)")
        self.norms = nn.ModuleList()
        self.img_size = img_size
        g = int(math.sqrt(num_patches))
        self.layer_norm = nn.LayerNorm((num_channels, g, g))
        self.num_channels = num_channels
        self.flow_type = "AllInOneBlock"
        self.grid = g
        self.hidden = int((num_channels - num_channels // 2) * hidden_ratio)
        if self.hidden > HIDDEN_PAD:
            raise ValueError(f"hidden channels {self.hidden} > {HIDDEN_PAD} unsupported (hidden_ratio <= 0.16)")
        self.fast_flow_decoder = _Sequence()
        for i in range(flow_steps):
            self.fast_flow_decoder.module_list.append(_Step(num_channels, self.hidden, 1 if i % 2 == 1 else 3))
        self._packed = None
        self._handle = custom_ops.register_module(self)

    def _apply(self, fn, recurse=True):
        self._packed = None
        return super()._apply(fn, recurse)

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._packed = None
        return super().load_state_dict(state_dict, strict=strict, **kw)

    # -- weight packing ----------------------------------------------------------------------------
    def _pack(self, device):
        Cn = self.num_channels
        c2 = Cn // 2
        c1 = Cn - c2
        keep = []
        steps = (_lib.NfStep * len(self.fast_flow_decoder.module_list))()
        logdet_const = 0.0
        # interleave: tile n holds s-channels n*48.. then t-channels c2 + n*48..
        order = []
        for n in range(c2 // TILE_HALF):
            order += list(range(n * TILE_HALF, (n + 1) * TILE_HALF))
            order += list(range(c2 + n * TILE_HALF, c2 + (n + 1) * TILE_HALF))
        order = torch.tensor(order, dtype=torch.long)

        def dev(t, dtype):
            t = t.detach().to(device=device, dtype=dtype).contiguous()
            keep.append(t)
            return t.data_ptr()

        for i, st in enumerate(self.fast_flow_decoder.module_list):
            conv0, conv2 = st.subnet[0], st.subnet[2]
            k = conv0.kernel_size[0]
            w0 = conv0.weight.detach().float().cpu()  # [hidden, c1, k, k]
            w0p = torch.zeros(HIDDEN_PAD, k * k, c1)
            w0p[: self.hidden] = w0.permute(0, 2, 3, 1).reshape(self.hidden, k * k, c1)
            b0p = torch.zeros(HIDDEN_PAD)
            b0p[: self.hidden] = conv0.bias.detach().float().cpu()
            w2 = conv2.weight.detach().float().cpu()  # [2*c2, hidden, k, k]
            w2p = torch.zeros(Cn, k * k, HIDDEN_PAD)
            w2p[:, :, : self.hidden] = w2.permute(0, 2, 3, 1).reshape(Cn, k * k, self.hidden)
            w2p = (0.1 * w2p)[order]
            b2p = (0.1 * conv2.bias.detach().float().cpu())[order]
            scale = 0.1 * torch.nn.functional.softplus(st.global_scale.detach().float().cpu().reshape(Cn), beta=0.5)
            offset = st.global_offset.detach().float().cpu().reshape(Cn)
            perm = st.w_perm.detach().float().cpu().reshape(Cn, Cn).argmax(dim=1)  # out[i] = y[perm[i]]
            inv_perm = torch.empty(Cn, dtype=torch.int32)
            inv_perm[perm] = torch.arange(Cn, dtype=torch.int32)
            logdet_const += float(self.grid * self.grid * torch.log(scale).sum())
            s = steps[i]
            s.w0p, s.b0p = dev(w0p.reshape(HIDDEN_PAD, -1), torch.float16), dev(b0p, torch.float32)
            s.w2p, s.b2p = dev(w2p.reshape(Cn, -1), torch.float16), dev(b2p, torch.float32)
            s.scale, s.offset = dev(scale, torch.float32), dev(offset, torch.float32)
            s.inv_perm = dev(inv_perm, torch.int32)
            s.ksize = k
        w = _lib.NfWeights()
        w.channels, w.grid, w.hidden_pad, w.steps = Cn, self.grid, HIDDEN_PAD, len(steps)
        w.clamp, w.logdet_const = 2.0, logdet_const
        w.step = C.cast(steps, C.POINTER(_lib.NfStep))
        self._packed = dict(w=w, steps=steps, keep=keep, device=device, ws=None, ws_batch=0)

    # -- forward -----------------------------------------------------------------------------------
    def forward_tokens(self, tokens: torch.Tensor) -> NormalizingFlowReturn:
        """tokens: fp32 [B, P, C] patch embedding (the layout the encoder produces)."""
        if not tokens.is_cuda:
            raise RuntimeError("NormalizingFlow (vitad): CUDA input required — no CPU path")
        B, P, Cn = tokens.shape
        if P != self.grid * self.grid or Cn != self.num_channels:
            raise ValueError(f"expected [B,{self.grid * self.grid},{self.num_channels}] tokens, got {tuple(tokens.shape)}")
        omp, loss_terms = torch.ops.vitad.nf_forward(tokens, self._handle)
        amap, amax = ops.bilinear_up(omp, self.img_size, align_corners=False, want_max=True)
        out = NormalizingFlowReturn(loss=loss_terms.mean(), anomaly_score_map=amap)
        out.image_max = amax  # amax(anomaly_score_map, (1,2,3)) computed by the upsample kernel (ValidatorNF.py:137-142)
        return out

    def _run(self, tokens: torch.Tensor):
        """CUDA implementation of torch.ops.vitad.nf_forward for this module's weights."""
        from .encoders import _param_key

        key = _param_key(self, tokens.device)
        if self._packed is None or self._packed.get("key") != key:
            self._pack(tokens.device)
            self._packed["key"] = key
        pk = self._packed
        B = tokens.shape[0]
        tokens = tokens.to(torch.float32).contiguous()
        if pk["ws"] is None or pk["ws_batch"] < B:
            nbytes = lib.vitad_nf_workspace_bytes(C.byref(pk["w"]), B)
            pk["ws"], pk["ws_batch"] = torch.empty(nbytes, device=tokens.device, dtype=torch.uint8), B
        omp = torch.empty((B, self.grid, self.grid), device=tokens.device, dtype=torch.float32)
        loss_terms = torch.empty((B,), device=tokens.device, dtype=torch.float32)
        check(lib.vitad_nf_forward(C.byref(pk["w"]), tokens.data_ptr(), B, pk["ws"].data_ptr(), pk["ws"].numel(),
                                   omp.data_ptr(), loss_terms.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return omp, loss_terms

    def forward(self, x: torch.Tensor) -> NormalizingFlowReturn:
        """x: [B, C, h, w] as in the reference (NormalizingFlow.py:118-123)."""
        if not x.is_cuda:
            raise RuntimeError("NormalizingFlow (vitad): CUDA input required — no CPU path")
        B, Cn, h, w = x.shape
        return self.forward_tokens(x.reshape(B, Cn, h * w).transpose(1, 2))
