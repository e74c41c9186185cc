"""Reconstruction models of the scoring path under the reference's names
(src/classes/transformer/TransformerAutoEncoder.py:152-194, src/classes/CnnAutoEncoder.py:18-74,
src/classes/CnnDecoder.py:16-117).

The encoder is the CUDA DeiT, the small CNN decoder one C-ABI call (`vitad_cnn_decoder_forward`: tcgen05 GEMMs for the two
Linear layers and for each stride-2 transposed convolution, BatchNorm folded, NHWC fp16 activations), the per-pixel L2
map + per-image max the CUDA kernel `vitad_l2_map_score`.  The nn.Modules only hold the parameters under the
reference's names, so reference checkpoints load unchanged.  Only the small CNN decoder (`ae_deit_small`,
decoder="cnn") is provided; the reverse-ResNet decoder raises.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
from torch import Tensor, nn

import ctypes as C

from . import _lib, ops
from ._lib import check, lib
from .encoders import EncoderDeit

BIAS_FILL = 0.001  # src/util/HelperFunctions.py:7


@dataclass
class AutoEncoderOutput:
    """Same fields as CnnAutoEncoder.py:18-24."""

    latent_space: Tensor
    reconstruction: Tensor
    patch_embedding: Tensor = None


def _init(m):
    """init_weights (HelperFunctions.py:19-23): xavier-normal weights, bias 0.001."""
    if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear)):
        nn.init.xavier_normal_(m.weight)
        m.bias.data.fill_(BIAS_FILL)


class DecoderVanillaCNN(nn.Module):
    """2 x Linear (z -> 2z -> 768*f*f) then 5 x (ConvTranspose2d k3 s2, BatchNorm, ReLU), Tanh after the last BN.
    Attribute names follow CnnDecoder.py:16-117 (`decoder_lin`, `recon_conv1..5`, `decoder_cnn`)."""

    def __init__(self, z_space: int = 0, first_feature_map_size: int = 0) -> None:
        super().__init__()
        self.use_linear = z_space != 0
        if self.use_linear:
            f = first_feature_map_size
            self.decoder_lin = nn.Sequential(nn.Linear(z_space, 2 * z_space), nn.ReLU(inplace=True),
                                             nn.Linear(2 * z_space, 768 * f * f), nn.ReLU(inplace=True))
            self.unflatten = nn.Unflatten(dim=1, unflattened_size=(768, f, f))
            self.decoder_lin.apply(_init)
        chans = [768, 384, 192, 96, 48, 3]
        layers = []
        for i in range(5):
            conv = nn.ConvTranspose2d(chans[i], chans[i + 1], kernel_size=3, stride=2, padding=1, output_padding=1)
            setattr(self, f"recon_conv{i + 1}", conv)
            layers += [conv, nn.BatchNorm2d(chans[i + 1]), nn.Tanh() if i == 4 else nn.ReLU(inplace=True)]
        self.decoder_cnn = nn.Sequential(*layers)
        self.decoder_cnn.apply(_init)

    # -- CUDA path ---------------------------------------------------------------------------------
    _packed = None

    def _apply(self, fn, recurse=True):
        self._packed = None
        return super()._apply(fn, recurse)

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._packed = None
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def _pack(self, device):
        """Fold BatchNorm (running statistics: this is the inference path) into per-phase GEMM weights, see
        csrc/decoder.cu.  ConvTranspose2d weight layout is [C_in, C_out, ky, kx] (CnnDecoder.py:47-87)."""
        keep = []

        def dev(t, dtype):
            t = t.detach().to(device=device, dtype=dtype).contiguous()
            keep.append(t)
            return t.data_ptr()

        lin1, lin2 = self.decoder_lin[0], self.decoder_lin[2]
        f = self.unflatten.unflattened_size[1]
        chans = [768, 384, 192, 96, 48]
        pitch = [768, 384, 192, 96, 64]
        w = _lib.CnnDecoderWeights()
        w.latent, w.hidden, w.grid0, w.last_cin = lin1.in_features, lin1.out_features, f, 48
        for i in range(5):
            w.chan[i] = pitch[i]
        w.lin1_w, w.lin1_b = dev(lin1.weight, torch.float16), dev(lin1.bias, torch.float32)
        # rows of the second Linear reordered from (c, h, w) (nn.Unflatten, :42-45) to (h, w, c): NHWC output
        w2 = lin2.weight.detach().float().view(768, f * f, -1).permute(1, 0, 2).reshape(768 * f * f, -1)
        b2 = lin2.bias.detach().float().view(768, f * f).t().reshape(-1)
        w.lin2_w, w.lin2_b = dev(w2, torch.float16), dev(b2, torch.float32)

        def folded(i):
            conv, bn = getattr(self, f"recon_conv{i + 1}"), self.decoder_cnn[3 * i + 1]
            s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
            bias = (conv.bias.detach().float() - bn.running_mean.detach().float()) * s + bn.bias.detach().float()
            return conv.weight.detach().float() * s.view(1, -1, 1, 1), bias  # [C_in, C_out, 3, 3] scaled per C_out

        for l in range(4):
            wt, bias = folded(l)
            cin, cout, cin_p, cout_p = chans[l], chans[l + 1], pitch[l], pitch[l + 1]
            wg = torch.zeros(4 * cout_p, 4 * cin_p)
            bg = torch.zeros(4 * cout_p)
            for a in range(2):
                for c in range(2):
                    r0 = (a * 2 + c) * cout_p
                    bg[r0:r0 + cout] = bias
                    for di in range(2):
                        for dj in range(2):
                            ky, kx = a + 1 - 2 * di, c + 1 - 2 * dj
                            if 0 <= ky <= 2 and 0 <= kx <= 2:
                                k0 = (di * 2 + dj) * cin_p
                                wg[r0:r0 + cout, k0:k0 + cin] = wt[:, :, ky, kx].t()
            w.conv_w[l], w.conv_b[l] = dev(wg, torch.float16), dev(bg, torch.float32)
        wt, bias = folded(4)  # [48, 3, 3, 3] -> [ky][kx][ci][co]
        w.last_w, w.last_b = dev(wt.permute(2, 3, 0, 1), torch.float32), dev(bias, torch.float32)
        self._packed = dict(w=w, keep=keep, device=device, ws=None, ws_batch=0)

    def forward(self, x):
        """latent [B, z_space] → reconstruction fp32 [B, 3, S, S] (tanh range).  BatchNorm uses its running statistics."""
        if not x.is_cuda:
            raise RuntimeError("DecoderVanillaCNN (vitad): CUDA input required — this implementation has no CPU path")
        if not self.use_linear:
            raise NotImplementedError("vitad DecoderVanillaCNN: only the latent (z_space) form used by AutoEncoderDeit is provided")
        if self._packed is None or self._packed["device"] != x.device:
            self._pack(x.device)
        pk = self._packed
        x = x.to(torch.float32).contiguous()
        B = x.shape[0]
        if pk["ws"] is None or pk["ws_batch"] < B:
            nbytes = lib.vitad_cnn_decoder_workspace_bytes(C.byref(pk["w"]), B)
            pk["ws"], pk["ws_batch"] = torch.empty(nbytes, device=x.device, dtype=torch.uint8), B
        size = 32 * pk["w"].grid0
        recon = torch.empty((B, 3, size, size), device=x.device, dtype=torch.float32)
        check(lib.vitad_cnn_decoder_forward(C.byref(pk["w"]), x.data_ptr(), B, pk["ws"].data_ptr(), pk["ws"].numel(),
                                            recon.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return recon


class AutoEncoderDeit(nn.Module):
    """Drop-in for TransformerAutoEncoder.py:152-194 (`ae_deit_small`): EncoderDeit → cls token → decoder."""

    def __init__(self, img_size: int, requires_grad: bool = False, red_mse="mean", red_ssim="elementwise_mean",
                 decoder="resnet") -> None:
        super().__init__()
        if decoder != "cnn":
            raise NotImplementedError(
                "vitad AutoEncoderDeit: only the small CNN decoder (get_model('ae_deit_small'), decoder='cnn') is "
                "provided; the reverse-ResNet decoder stack is outside this round's scope (DESIGN.md §8)")
        self.img_size = img_size
        self.red_mse = red_mse
        self.mse = nn.MSELoss(reduction=red_mse)
        self.encoder = EncoderDeit(img_size=img_size, requires_grad=requires_grad)
        self.z_space = self.encoder.size_patch_embedding
        self.feature_map_size = math.ceil(img_size / (2**5))
        self.size_patch_embedding = self.encoder.size_patch_embedding
        self.num_embedded_patches = self.encoder.num_embedded_patches
        self.decoder = DecoderVanillaCNN(z_space=self.z_space, first_feature_map_size=self.feature_map_size)
        self.architecture = "transformer"

    def forward(self, x) -> AutoEncoderOutput:
        output = self.encoder(x)
        x_recon = self.decoder(output.latent_space)
        return AutoEncoderOutput(latent_space=output.latent_space, reconstruction=x_recon,
                                 patch_embedding=output.patch_embedding)

    def MSELoss(self, output: Tensor, x: Tensor):
        """CnnAutoEncoder.py:68-74.  reduction='none' (forced by get_model) → elementwise squared error."""
        return self.mse(output, x)

    def anomaly_map_and_score(self, reconstruction: Tensor, x: Tensor):
        """mean_c (recon - x)^2 and its per-image max in one CUDA kernel (ValidatorRecon.py:109-116)."""
        return ops.l2_map_score(reconstruction, x)
