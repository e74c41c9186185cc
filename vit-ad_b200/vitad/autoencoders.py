"""Reconstruction models of the scoring path under the reference's names
(src/classes/transformer/TransformerAutoEncoder.py:152-194, src/classes/CnnAutoEncoder.py:18-74,
src/classes/CnnDecoder.py:16-117).

The encoder is the CUDA DeiT; the per-pixel L2 map + per-image max (the scoring tail) is the CUDA kernel
`vitad_l2_map_score`.  The decoder convolution stack is the step *before* that tail and is listed as the next
widening step in DESIGN.md: it runs through torch/cuDNN here, with the reference's parameter names so
reference checkpoints load unchanged.  Only the small CNN decoder (`ae_deit_small`, decoder="cnn") is provided;
the reverse-ResNet decoder raises.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
from torch import Tensor, nn

from . import ops
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

    def forward(self, x):
        if self.use_linear:
            x = self.unflatten(self.decoder_lin(x))
        return self.decoder_cnn(x)


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
