"""Seeded synthetic state_dicts and inputs with the reference's key layout, for benchmarks, diagnostics and tests.

There is no network for checkpoints, so tests and the bench use random-init weights "of that
architecture".  Distributions follow the reference constructors; `stress=True`
additionally perturbs biases / LayerNorm affines / attention logit scale so that code paths a default
init leaves trivial (zero biases, uniform softmax) are exercised by the parity tests.
Key layouts: SURVEY.md §8b.  ``oracle/make_golden.py`` proves the layouts by load_state_dict(strict=True)
into the reference's own classes.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def _tn(g, shape, std):
    # trunc_normal_(std=.02, a=-2, b=2) truncates at +-100 sigma: plain normal in practice.
    return torch.randn(shape, generator=g) * std


def _kaiming_uniform(g, shape, fan_in):
    # nn.Linear / nn.Conv2d default: kaiming_uniform_(a=sqrt(5)) -> U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    bound = 1.0 / math.sqrt(fan_in)
    return (torch.rand(shape, generator=g) * 2 - 1) * bound


def _xavier_normal(g, shape):
    fan_out, fan_in = shape[0], shape[1]
    rf = 1
    for s in shape[2:]:
        rf *= s
    std = math.sqrt(2.0 / ((fan_in + fan_out) * rf))
    return torch.randn(shape, generator=g) * std


def make_deit_state_dict(seed: int = 0, stress: bool = False, depth: int = 12, prefix: str = "deit.",
                         distilled: bool = True) -> dict:
    """timm deit_base_distilled_patch16_224(pretrained=False) init (TransformerEncoder.py:134-136); with
    distilled=False the same for vit_base_patch16_224 (EncoderVit, :193): one prefix token, no dist_token/head_dist."""
    g = torch.Generator().manual_seed(seed)
    C, H = 768, 3072
    sd = {}
    sd["cls_token"] = torch.randn(1, 1, C, generator=g) * (0.02 if stress else 1e-6)
    sd["pos_embed"] = _tn(g, (1, 198 if distilled else 197, C), 0.02)
    if distilled:
        sd["dist_token"] = _tn(g, (1, 1, C), 0.02)
    sd["patch_embed.proj.weight"] = _kaiming_uniform(g, (C, 3, 16, 16), 3 * 256)
    sd["patch_embed.proj.bias"] = _kaiming_uniform(g, (C,), 3 * 256)

    def ln(name):
        if stress:
            sd[name + ".weight"] = 1 + 0.1 * torch.randn(C, generator=g)
            sd[name + ".bias"] = 0.1 * torch.randn(C, generator=g)
        else:
            sd[name + ".weight"] = torch.ones(C)
            sd[name + ".bias"] = torch.zeros(C)

    def lin(name, out_f, in_f, scale=1.0):
        sd[name + ".weight"] = _tn(g, (out_f, in_f), 0.02) * scale
        sd[name + ".bias"] = 0.02 * torch.randn(out_f, generator=g) if stress else torch.zeros(out_f)

    for i in range(depth):
        b = f"blocks.{i}."
        ln(b + "norm1")
        lin(b + "attn.qkv", 3 * C, C)
        if stress:  # make attention logits O(1) so softmax is far from uniform
            sd[b + "attn.qkv.weight"][: 2 * C] *= 2.5
        lin(b + "attn.proj", C, C)
        ln(b + "norm2")
        lin(b + "mlp.fc1", H, C)
        lin(b + "mlp.fc2", C, H)
    ln("norm")
    lin("head", 1000, C)
    if distilled:
        lin("head_dist", 1000, C)
    return {prefix + k: v for k, v in sd.items()}


def make_vit_state_dict(seed: int = 0, stress: bool = False) -> dict:
    """EncoderVit's state_dict (keys `vit.*`, timm vit_base_patch16_224)."""
    return make_deit_state_dict(seed=seed, stress=stress, prefix="vit.", distilled=False)


def make_esvit_state_dict(seed: int = 0, stress: bool = False, prefix: str = "esvit.") -> dict:
    """Vendored SwinTransformer(embed 96, depths 2/2/6/2, heads 3/6/12/24, window 14, num_classes 3) as EncoderEsVit
    builds it (TransformerEncoder.py:228-240): Linear trunc-normal(.02)/zero bias, LayerNorm 1/0, bias tables
    trunc-normal(.02) (SwinTransformerModule.py:132,800-808); 185 keys incl. the relative_position_index buffers."""
    from .encoders import SWIN_DEPTHS, SWIN_EMBED, SWIN_HEADS, SWIN_WINDOW
    from .encoders import _relative_position_index as swin_relative_position_index

    g = torch.Generator().manual_seed(seed)
    sd = {}

    def ln(name, c):
        sd[name + ".weight"] = 1 + 0.1 * torch.randn(c, generator=g) if stress else torch.ones(c)
        sd[name + ".bias"] = 0.1 * torch.randn(c, generator=g) if stress else torch.zeros(c)

    def lin(name, out_f, in_f, bias=True):
        sd[name + ".weight"] = _tn(g, (out_f, in_f), 0.02)
        if bias:
            sd[name + ".bias"] = 0.02 * torch.randn(out_f, generator=g) if stress else torch.zeros(out_f)

    sd["patch_embed.proj.weight"] = _kaiming_uniform(g, (SWIN_EMBED, 3, 4, 4), 48)
    sd["patch_embed.proj.bias"] = _kaiming_uniform(g, (SWIN_EMBED,), 48)
    ln("patch_embed.norm", SWIN_EMBED)
    res = 56
    for s, (depth, heads) in enumerate(zip(SWIN_DEPTHS, SWIN_HEADS)):
        C = SWIN_EMBED * 2**s
        ws = min(SWIN_WINDOW, res)
        for b in range(depth):
            p = f"layers.{s}.blocks.{b}."
            ln(p + "norm1", C)
            sd[p + "attn.relative_position_bias_table"] = _tn(g, ((2 * ws - 1) ** 2, heads), 0.5 if stress else 0.02)
            sd[p + "attn.relative_position_index"] = swin_relative_position_index(ws)
            lin(p + "attn.qkv", 3 * C, C)
            if stress:
                sd[p + "attn.qkv.weight"][: 2 * C] *= 3.0
            lin(p + "attn.proj", C, C)
            ln(p + "norm2", C)
            lin(p + "mlp.fc1", 4 * C, C)
            lin(p + "mlp.fc2", C, 4 * C)
        if s < 3:
            lin(f"layers.{s}.downsample.reduction", 2 * C, 4 * C, bias=False)
            ln(f"layers.{s}.downsample.norm", 4 * C)
            res //= 2
    ln("norm", 768)
    lin("head", 3, 768)
    return {prefix + k: v for k, v in sd.items()}


def make_mdn_state_dict(seed: int, num_gaussians: int, dim: int = 768, stress: bool = False) -> dict:
    """GaussianMixtureDensityNetwork.__init__ (MixtureDensityNetwork.py:117-149): xavier-normal weights;
    pi/sigma biases keep the nn.Linear default, mu bias = 0.001 (HelperFunctions.py:19-23)."""
    g = torch.Generator().manual_seed(seed)
    K = num_gaussians
    sd = {
        "pi.weight": _xavier_normal(g, (K, dim)),
        "pi.bias": _kaiming_uniform(g, (K,), dim),
        "sigma.weight": _xavier_normal(g, (dim * K, dim)),
        "sigma.bias": _kaiming_uniform(g, (dim * K,), dim),
        "mu.weight": _xavier_normal(g, (dim * K, dim)),
        "mu.bias": torch.full((dim * K,), 0.001),
    }
    if stress:  # trained-like spread: wider sigma range (both ELU branches) and mixture means
        sd["sigma.weight"] *= 6.0
        sd["mu.weight"] *= 6.0
        sd["mu.bias"] = 0.3 * torch.randn(dim * K, generator=g)
        sd["sigma.bias"] = 0.3 * torch.randn(dim * K, generator=g)
    return sd


def make_nf_state_dict(seed: int, channels: int = 768, grid: int = 14, hidden_ratio: float = 0.16,
                       flow_steps: int = 20, stress: bool = False, subnet_gain: float = 4.0) -> dict:
    """NormalizingFlow.__init__ (NormalizingFlow.py:28-116) with FrEIA AllInOneBlock defaults.  `stress`: random global
    affine + the coupling subnets' output layer scaled by `subnet_gain` (the default init leaves the flow near identity;
    4 drives the image scores towards saturation, ~2 keeps them mid-range where they separate images best)."""
    g = torch.Generator().manual_seed(seed)
    rng = np.random.RandomState(seed)
    c2 = channels // 2
    c1 = channels - c2
    hidden = int(c1 * hidden_ratio)
    sd = {"layer_norm.weight": torch.ones(channels, grid, grid), "layer_norm.bias": torch.zeros(channels, grid, grid)}
    gs0 = float(2.0 * np.log(np.exp(0.5 * 10.0 * 1.0) - 1))
    for i in range(flow_steps):
        p = f"fast_flow_decoder.module_list.{i}."
        k = 1 if i % 2 == 1 else 3
        sd[p + "global_scale"] = torch.full((1, channels, 1, 1), gs0)
        sd[p + "global_offset"] = torch.zeros(1, channels, 1, 1)
        if stress:
            sd[p + "global_scale"] = sd[p + "global_scale"] + 0.5 * torch.randn(1, channels, 1, 1, generator=g)
            sd[p + "global_offset"] = 0.05 * torch.randn(1, channels, 1, 1, generator=g)
        w = np.zeros((channels, channels), dtype=np.float32)
        for r, c in enumerate(rng.permutation(channels)):
            w[r, c] = 1.0
        sd[p + "w_perm"] = torch.from_numpy(w).view(channels, channels, 1, 1)
        sd[p + "w_perm_inv"] = torch.from_numpy(w.T.copy()).view(channels, channels, 1, 1)
        sd[p + "subnet.0.weight"] = _kaiming_uniform(g, (hidden, c1, k, k), c1 * k * k)
        sd[p + "subnet.0.bias"] = _kaiming_uniform(g, (hidden,), c1 * k * k)
        sd[p + "subnet.2.weight"] = _kaiming_uniform(g, (2 * c2, hidden, k, k), hidden * k * k)
        sd[p + "subnet.2.bias"] = _kaiming_uniform(g, (2 * c2,), hidden * k * k)
        if stress:
            sd[p + "subnet.2.weight"] *= subnet_gain
    return sd


def make_small_decoder_state_dict(seed: int, z_space: int = 768, fmap: int = 7, prefix: str = "decoder.") -> dict:
    """DecoderVanillaCNN(z_space=768, first_feature_map_size=7) (CnnDecoder.py:16-117): init_weights = xavier-normal
    weights, bias 0.001; BatchNorm affine 1/0 with non-trivial (seeded) running statistics so eval-mode BN is
    exercised.  The Sequential re-registers the conv modules, so both key spellings are emitted."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    sd["decoder_lin.0.weight"] = _xavier_normal(g, (2 * z_space, z_space))
    sd["decoder_lin.0.bias"] = torch.full((2 * z_space,), 0.001)
    sd["decoder_lin.2.weight"] = _xavier_normal(g, (768 * fmap * fmap, 2 * z_space))
    sd["decoder_lin.2.bias"] = torch.full((768 * fmap * fmap,), 0.001)
    chans = [768, 384, 192, 96, 48, 3]
    for i in range(5):
        cin, cout = chans[i], chans[i + 1]
        # ConvTranspose2d weight is [in, out, k, k]; xavier fans as torch computes them for that shape
        std = math.sqrt(2.0 / ((cin + cout) * 9))
        w = torch.randn(cin, cout, 3, 3, generator=g) * std
        b = torch.full((cout,), 0.001)
        for name in (f"recon_conv{i + 1}", f"decoder_cnn.{3 * i}"):
            sd[name + ".weight"], sd[name + ".bias"] = w, b
        bn = f"decoder_cnn.{3 * i + 1}."
        sd[bn + "weight"] = 1 + 0.1 * torch.randn(cout, generator=g)
        sd[bn + "bias"] = 0.05 * torch.randn(cout, generator=g)
        sd[bn + "running_mean"] = 0.01 * torch.randn(cout, generator=g)
        sd[bn + "running_var"] = 0.5 + torch.rand(cout, generator=g)
        sd[bn + "num_batches_tracked"] = torch.tensor(1)
    return {prefix + k: v for k, v in sd.items()}


RESNET_DECODER_LAYERS = (("layer4", 512, 3, 2, 1024), ("layer3", 256, 4, 2, 512), ("layer2", 128, 6, 2, 256),
                         ("layer1", 64, 3, 1, 64))  # name, planes, blocks, stride of the last block, its output channels


def make_resnet_decoder_state_dict(seed: int, embedding: int = 768, prefix: str = "decoder.") -> dict:
    """DecoderResNetVariableEmbeddingSize(embedding_size=768) (CnnDecoder.py:158-196, ReverseResNet.py:106-209): the key
    layout of the reference's state_dict.  The reference's default initialisation leaves the transposed convolutions at
    kaiming-uniform(a=sqrt 5) and the BatchNorms at identity, so activations die out through 16 blocks and the image is
    tanh(~0); these weights are He-scaled with non-trivial BatchNorm affine/running statistics instead, so that every
    layer (and eval-mode BatchNorm folding) matters in the parity tests."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def bn(name, c, gamma=1.0):
        sd[name + ".weight"] = gamma * (1 + 0.1 * torch.randn(c, generator=g))
        sd[name + ".bias"] = 0.05 * torch.randn(c, generator=g)
        sd[name + ".running_mean"] = 0.05 * torch.randn(c, generator=g)
        sd[name + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        sd[name + ".num_batches_tracked"] = torch.tensor(1)

    def convt(name, cin, cout, k, taps):
        # ConvTranspose2d weight is [in, out, k, k]; `taps` = kernel elements that reach one output pixel on average
        sd[name + ".weight"] = torch.randn(cin, cout, k, k, generator=g) * math.sqrt(2.0 / (cin * taps))

    sd["fc1.0.weight"] = torch.randn(2 * embedding, embedding, generator=g) * math.sqrt(2.0 / embedding)
    sd["fc1.0.bias"] = 0.02 * torch.randn(2 * embedding, generator=g)
    sd["fc2.0.weight"] = torch.randn(2048, 2 * embedding, generator=g) * math.sqrt(2.0 / (2 * embedding))
    sd["fc2.0.bias"] = 0.02 * torch.randn(2048, generator=g)
    for layer, planes, blocks, stride, last_dim in RESNET_DECODER_LAYERS:
        cin = planes * 4
        for i in range(blocks):
            b = f"{layer}.{i}"
            last = i == blocks - 1
            cout = last_dim if last else cin
            convt(b + ".conv3", cin, planes, 1, 1)
            bn(b + ".bn3", planes)
            convt(b + ".conv2", planes, planes, 3, 9 / 4 if (last and stride == 2) else 9)
            bn(b + ".bn2", planes)
            convt(b + ".conv1", planes, cout, 1, 1)
            bn(b + ".bn1", cout, gamma=0.5)
            if last:
                convt(b + ".upsample.0", cin, cout, 1, 1)
                bn(b + ".upsample.1", cout, gamma=0.7)
    convt("de_conv1", 64, 3, 7, 49 / 4 * 4)  # nearest-upsampled input: neighbouring taps see the same pixel
    bn("bn1", 3, gamma=0.6)
    return {prefix + k: v for k, v in sd.items()}


def synthetic_images(seed: int, batch: int, size: int = 224) -> torch.Tensor:
    """fp32 NCHW in [0,1] — the loader's ToTensor contract (GeneralDataset.py:38-59)."""
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.rand(batch, 3, size, size, generator=g)


def synthetic_esvit_checkpoint(seed: int = 61):
    """A `student` state dict as an EsViT checkpoint trained with window 7 would hold it (tables 169 x nH, index
    49 x 49 in the first three stages): the case interpolate_position_encoding (TransformerEncoder.py:276-350) exists for."""
    sd = {k[len("esvit."):]: v.clone() for k, v in make_esvit_state_dict(seed=seed, stress=True).items()}
    g = torch.Generator().manual_seed(seed + 1)
    from .encoders import _relative_position_index as swin_relative_position_index

    for k in list(sd):
        if k.endswith("relative_position_bias_table") and sd[k].shape[0] == 729:
            sd[k] = torch.randn(169, sd[k].shape[1], generator=g) * 0.02
        if k.endswith("relative_position_index") and sd[k].shape[0] == 196:
            sd[k] = swin_relative_position_index(7)
    return {k: v for k, v in sd.items() if not k.startswith("head")}
