"""CPU oracle for the vit-ad scoring path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional fp32 restatement (torch on CPU) of what the reference computes between
``images.to(device)`` and the ``.cpu().numpy()`` calls of its validators.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import it;
the product package (vit-ad_b200/) never does.

Pinning: the reference ships no tests or golden vectors, so this oracle is pinned against
outputs of the reference's own Python classes run in the build container (``oracle/make_golden.py`` →
``tests/golden/*.npz``; checked by ``tests/test_oracle_cpu.py``).  Two third-party pieces the reference
imports are not vendored and not installable here — timm 0.6.13 (DeiT) and FrEIA 0.2 (AllInOneBlock);
their arithmetic is restated from the published sources in ``oracle/shims`` and here, and DeiT is
cross-checked against ``transformers.DeiTModel``.  For those two boundaries parity is pinned to the
restatement, not to the original packages.

All ``path:line`` citations are relative to the reference root.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LOG_SQRT_2PI = 0.5 * math.log(2 * math.pi)


# ----------------------------------------------------------------------------------------------
# DeiT-B distilled 16/224 encoder  (src/classes/transformer/TransformerEncoder.py:145-173; timm 0.6.13)
# ----------------------------------------------------------------------------------------------
def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu_erf(x: Tensor) -> Tensor:
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def deit_block(sd: dict, pre: str, x: Tensor, heads: int = 12) -> Tensor:
    """timm Block: x + attn(norm1(x)); x + mlp(norm2(x)).  LayerNorm eps 1e-6."""
    B, N, C = x.shape
    hd = C // heads
    h = layer_norm(x, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-6)
    qkv = h @ sd[pre + "attn.qkv.weight"].t() + sd[pre + "attn.qkv.bias"]
    qkv = qkv.reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = torch.softmax((q @ k.transpose(-2, -1)) * hd**-0.5, dim=-1)
    o = (attn @ v).transpose(1, 2).reshape(B, N, C)
    x = x + o @ sd[pre + "attn.proj.weight"].t() + sd[pre + "attn.proj.bias"]
    h = layer_norm(x, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-6)
    h = gelu_erf(h @ sd[pre + "mlp.fc1.weight"].t() + sd[pre + "mlp.fc1.bias"])
    return x + h @ sd[pre + "mlp.fc2.weight"].t() + sd[pre + "mlp.fc2.bias"]


def deit_forward(sd: dict, images: Tensor, block_index: int = 0, prefix: str = "deit.", depth: int = 12):
    """EncoderDeit.forward (TransformerEncoder.py:145-173) → (patch_embedding [B,196,768], cls [B,768]).

    block_index == 0: timm forward_features (all blocks, one final norm).  block_index != 0: blocks
    0..block_index with the final norm applied after EVERY block (:161-163).
    """
    p = prefix
    w = sd[p + "patch_embed.proj.weight"]  # [768,3,16,16]
    B = images.shape[0]
    ps = w.shape[-1]
    g = images.shape[-1] // ps
    # non-overlapping conv == per-patch GEMM; column order (c, i, j) as in the conv weight
    patches = images.reshape(B, 3, g, ps, g, ps).permute(0, 2, 4, 1, 3, 5).reshape(B, g * g, 3 * ps * ps)
    x = patches @ w.reshape(w.shape[0], -1).t() + sd[p + "patch_embed.proj.bias"]
    n_prefix = 2 if (p + "dist_token") in sd else 1  # DeiT distilled: cls + dist; plain ViT (EncoderVit): cls
    if n_prefix == 2:
        x = torch.cat((sd[p + "cls_token"].expand(B, -1, -1), sd[p + "dist_token"].expand(B, -1, -1), x), dim=1)
    else:
        x = torch.cat((sd[p + "cls_token"].expand(B, -1, -1), x), dim=1)
    x = x + sd[p + "pos_embed"]
    if block_index != 0:
        for i in range(block_index + 1):
            x = deit_block(sd, f"{p}blocks.{i}.", x)
            x = layer_norm(x, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    else:
        for i in range(depth):
            x = deit_block(sd, f"{p}blocks.{i}.", x)
        x = layer_norm(x, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    return x[:, n_prefix:, :], x[:, 0, :]


def vit_forward(sd: dict, images: Tensor):
    """EncoderVit.forward (TransformerEncoder.py:196-208): timm vit_base_patch16_224 forward_features, cls token
    dropped → (patch_embedding [B,196,768], cls [B,768]).  block_index is ignored by the reference class."""
    return deit_forward(sd, images, block_index=0, prefix="vit.")


# ----------------------------------------------------------------------------------------------
# EsViT Swin-T (window 14) encoder  (src/classes/transformer/SwinTransformerModule.py; EncoderEsVit,
# src/classes/transformer/TransformerEncoder.py:211-273).  Evaluated in .eval() mode: the reference leaves the
# encoder in train mode during MDN/NF validation, where DropPath(0.1) makes its features random
#  — a deliberate, documented deviation.
# ----------------------------------------------------------------------------------------------
SWIN_DEPTHS = (2, 2, 6, 2)
SWIN_HEADS = (3, 6, 12, 24)
SWIN_EMBED = 96
SWIN_WINDOW = 14


def swin_relative_position_index(ws: int) -> Tensor:
    """WindowAttention.__init__ (SwinTransformerModule.py:117-131)."""
    coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def swin_region_ids(H: int, ws: int, shift: int) -> Tensor:
    """Region label of every position of the shifted frame (create_attn_mask, :316-347): [H, H] in 0..8."""
    ids = torch.zeros(H, H, dtype=torch.long)
    bounds = (slice(0, -ws), slice(-ws, -shift), slice(-shift, None))
    cnt = 0
    for hs in bounds:
        for wsl in bounds:
            ids[hs, wsl] = cnt
            cnt += 1
    return ids


def swin_windows(x: Tensor, ws: int) -> Tensor:
    """window_partition (:50-63): [B,H,W,C] → [B*nW, ws*ws, C]."""
    B, H, W, C = x.shape
    return x.view(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)


def swin_block(sd: dict, pre: str, x: Tensor, H: int, heads: int, ws: int, shift: int) -> Tensor:
    """SwinTransformerBlock.forward (:349-416) + WindowAttention.forward (:144-193), eval mode, no padding
    (every stage resolution of a 224 input is a multiple of its window)."""
    B, L, C = x.shape
    hd = C // heads
    T = ws * ws
    h = layer_norm(x, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-5).view(B, H, H, C)
    if shift > 0:
        h = torch.roll(h, shifts=(-shift, -shift), dims=(1, 2))
    win = swin_windows(h, ws)  # [B*nW, T, C]
    qkv = (win @ sd[pre + "attn.qkv.weight"].t() + sd[pre + "attn.qkv.bias"]).reshape(-1, T, 3, heads, hd)
    qkv = qkv.permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * hd**-0.5, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    bias = sd[pre + "attn.relative_position_bias_table"][sd[pre + "attn.relative_position_index"].view(-1)]
    attn = attn + bias.view(T, T, heads).permute(2, 0, 1).unsqueeze(0)
    if shift > 0:
        ids = swin_windows(swin_region_ids(H, ws, shift).view(1, H, H, 1).float(), ws).view(-1, T)  # [nW, T]
        mask = (ids.unsqueeze(1) - ids.unsqueeze(2) != 0).float() * -100.0  # [nW, T, T]
        nW = mask.shape[0]
        attn = (attn.view(B, nW, heads, T, T) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, T, T)
    attn = torch.softmax(attn, dim=-1)
    o = (attn @ v).transpose(1, 2).reshape(-1, T, C)
    o = o @ sd[pre + "attn.proj.weight"].t() + sd[pre + "attn.proj.bias"]
    o = o.view(B, H // ws, H // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, H, C)  # window_reverse
    if shift > 0:
        o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
    x = x + o.reshape(B, L, C)
    h = layer_norm(x, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-5)
    h = gelu_erf(h @ sd[pre + "mlp.fc1.weight"].t() + sd[pre + "mlp.fc1.bias"])
    return x + h @ sd[pre + "mlp.fc2.weight"].t() + sd[pre + "mlp.fc2.bias"]


def swin_forward(sd: dict, images: Tensor, prefix: str = "esvit."):
    """EncoderEsVit.forward (TransformerEncoder.py:269-273) = SwinTransformer.forward_features (:821-837) →
    (patch_embedding = x_region [B,49,768], latent_space = avg-pooled [B,768])."""
    p = prefix
    B = images.shape[0]
    w = sd[p + "patch_embed.proj.weight"]  # [96,3,4,4]
    ps = w.shape[-1]
    g = images.shape[-1] // ps
    patches = images.reshape(B, 3, g, ps, g, ps).permute(0, 2, 4, 1, 3, 5).reshape(B, g * g, 3 * ps * ps)
    x = patches @ w.reshape(w.shape[0], -1).t() + sd[p + "patch_embed.proj.bias"]
    x = layer_norm(x, sd[p + "patch_embed.norm.weight"], sd[p + "patch_embed.norm.bias"], 1e-5)
    H = g
    for s, (depth, heads) in enumerate(zip(SWIN_DEPTHS, SWIN_HEADS)):
        ws = min(SWIN_WINDOW, H)
        for b in range(depth):
            shift = 0 if (b % 2 == 0 or H <= SWIN_WINDOW) else SWIN_WINDOW // 2
            x = swin_block(sd, f"{p}layers.{s}.blocks.{b}.", x, H, heads, ws, shift)
        if s < len(SWIN_DEPTHS) - 1:  # PatchMerging.forward (:478-505)
            C = x.shape[-1]
            xv = x.view(B, H, H, C)
            xm = torch.cat([xv[:, 0::2, 0::2], xv[:, 1::2, 0::2], xv[:, 0::2, 1::2], xv[:, 1::2, 1::2]], -1)
            xm = xm.view(B, -1, 4 * C)
            d = f"{p}layers.{s}.downsample."
            xm = layer_norm(xm, sd[d + "norm.weight"], sd[d + "norm.bias"], 1e-5)
            x = xm @ sd[d + "reduction.weight"].t()
            H //= 2
    x_region = layer_norm(x, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    return x_region, x_region.mean(dim=1)


# ----------------------------------------------------------------------------------------------
# MDN / "GMM" head  (src/classes/MixtureDensityNetwork.py)
# ----------------------------------------------------------------------------------------------
def gumbel_noise(shape, generator: torch.Generator) -> Tensor:
    """Same recipe as torch.nn.functional.gumbel_softmax: g = -log(Exp(1))."""
    return -torch.empty(shape).exponential_(generator=generator).log()


def mdn_log_pi(x: Tensor, sd: dict, gumbel: Tensor) -> Tensor:
    """log(softmax(pi(x) + g) + 1e-15)   (MixtureDensityNetwork.py:62-64,164; tau = 1)."""
    pi = x @ sd["pi.weight"].t() + sd["pi.bias"]
    return torch.log(torch.softmax(pi + gumbel, dim=-1) + 1e-15)


def mdn_patch_loglik(x: Tensor, sd: dict, gumbel: Tensor, token_chunk: int = 256) -> Tensor:
    """L[b,p] = mean_d logsumexp_k( log_pi[b,p,k] + log N(x[b,p,d]; mu[b,p,d,k], sigma[b,p,d,k]) ).

    forward (:151-171): sigma = ELU(W_s x + b_s) + 1 + 1e-15, mu = W_m x + b_m, viewed [B,P,D,K] (k fastest);
    log_gaussian_density (:35-46); log_likelihood (:49-72); mean over features (:86-88).
    Token-chunked so the [B,P,D,K] tensors are never materialised whole.
    """
    B, P, D = x.shape
    K = sd["pi.weight"].shape[0]
    log_pi = mdn_log_pi(x, sd, gumbel).reshape(B * P, K)
    xf = x.reshape(B * P, D)
    out = torch.empty(B * P, dtype=x.dtype)
    ws, bs, wm, bm = sd["sigma.weight"], sd["sigma.bias"], sd["mu.weight"], sd["mu.bias"]
    for s in range(0, B * P, token_chunk):
        xc = xf[s : s + token_chunk]
        sigma = (F.elu(xc @ ws.t() + bs) + 1 + 1e-15).view(-1, D, K)
        mu = (xc @ wm.t() + bm).view(-1, D, K)
        dens = -torch.log(sigma) - LOG_SQRT_2PI - 0.5 * torch.pow((xc.unsqueeze(-1) - mu) / sigma, 2)
        ll = torch.logsumexp(log_pi[s : s + token_chunk].unsqueeze(1) + dens, dim=-1)  # [n, D]
        out[s : s + token_chunk] = ll.mean(dim=1)
    return out.view(B, P)


def mdn_probability_map(L: Tensor) -> Tensor:
    """get_probability_map tail (:90-95): subtract the BATCH-global max, exp."""
    return torch.exp(L - L.max())


def bilinear_upsample(x: Tensor, out_size: int, align_corners: bool) -> Tensor:
    """Explicit bilinear interpolation of [N,1,h,w] → [N,1,S,S] with PyTorch's index conventions
    (used at ValidatorMDN.py:150-158 with align_corners=True and NormalizingFlow.py:138-143 with False)."""
    N, C, h, w = x.shape

    def src_index(o: Tensor, in_size: int) -> tuple[Tensor, Tensor, Tensor]:
        if align_corners:
            scale = (in_size - 1) / (out_size - 1) if out_size > 1 else 0.0
            s = o * scale
        else:
            scale = in_size / out_size
            s = torch.clamp((o + 0.5) * scale - 0.5, min=0.0)
        i0 = torch.clamp(s.floor().long(), max=in_size - 1)
        i1 = torch.clamp(i0 + 1, max=in_size - 1)
        lam = (s - i0.to(s.dtype)).to(x.dtype)
        return i0, i1, lam

    o = torch.arange(out_size, dtype=torch.float32)
    y0, y1, ly = src_index(o, h)
    x0, x1, lx = src_index(o, w)
    top = x[:, :, y0][:, :, :, x0] * (1 - lx) + x[:, :, y0][:, :, :, x1] * lx
    bot = x[:, :, y1][:, :, :, x0] * (1 - lx) + x[:, :, y1][:, :, :, x1] * lx
    return top * (1 - ly).view(1, 1, -1, 1) + bot * ly.view(1, 1, -1, 1)


def mdn_scores(prob: Tensor, img_size: int, patch_size: int):
    """ValidatorMdn.valid_loop_transformer tail (src/pipeline/ValidatorMDN.py:133-172):
    image score = 1 - amin_p(prob); pixel map = 1 - bilinear(prob as g×g, img_size, align_corners=True)."""
    B = prob.shape[0]
    g = int(img_size / patch_size)
    image_scores = 1 - prob.amin(dim=1)
    pixel = 1 - bilinear_upsample(prob.reshape(B, 1, g, g), img_size, align_corners=True)
    return image_scores, pixel


# ----------------------------------------------------------------------------------------------
# Normalizing-flow head  (src/classes/NormalizingFlow.py; FrEIA 0.2 AllInOneBlock)
# ----------------------------------------------------------------------------------------------
def nf_forward(sd: dict, x: Tensor, flow_steps: int, img_size: int, clamp: float = 2.0):
    """NormalizingFlow.forward (:118-145) → (loss [], anomaly_score_map [B,1,S,S], z, logdet).

    Per step i (kernel 3 if i even else 1, :96-100):  x1,x2 = split(C - C//2, C//2); a = 0.1*subnet(x1);
    s = clamp*tanh(a[:, :C//2]); y2 = x2*exp(s) + a[:, C//2:]; y = cat(x1,y2)*scale + offset with
    scale = 0.1*softplus_{beta=.5}(global_scale); out[:, i] = y[:, perm[i]] (w_perm is a 0/1 matrix);
    logdet += sum(s) + H*W*sum(log scale).  `layer_norm` exists in the state dict but is not applied (:124-125).
    """
    B, C, H, W = x.shape
    c2 = C // 2
    c1 = C - c2
    logdet = torch.zeros(B, dtype=x.dtype)
    for i in range(flow_steps):
        p = f"fast_flow_decoder.module_list.{i}."
        ksz = sd[p + "subnet.0.weight"].shape[-1]
        x1, x2 = x[:, :c1], x[:, c1:]
        h = F.relu(F.conv2d(x1, sd[p + "subnet.0.weight"], sd[p + "subnet.0.bias"], padding=ksz // 2))
        a = 0.1 * F.conv2d(h, sd[p + "subnet.2.weight"], sd[p + "subnet.2.bias"], padding=ksz // 2)
        s = clamp * torch.tanh(a[:, :c2])
        y = torch.cat((x1, x2 * torch.exp(s) + a[:, c2:]), dim=1)
        scale = 0.1 * F.softplus(sd[p + "global_scale"], beta=0.5)
        y = y * scale + sd[p + "global_offset"]
        perm = sd[p + "w_perm"].reshape(C, C).argmax(dim=1)  # row i has its single 1 at column perm[i]
        x = y[:, perm]
        logdet = logdet + s.sum(dim=(1, 2, 3)) + H * W * torch.log(scale).sum()
    z = x
    loss = torch.mean(0.5 * torch.sum(z**2, dim=(1, 2, 3)) - logdet)
    prob = torch.exp(-0.5 * torch.mean(z**2, dim=1, keepdim=True))
    amap = bilinear_upsample(1 - prob, img_size, align_corners=False)
    return loss, amap, z, logdet


def nf_scores(amap: Tensor) -> Tensor:
    """ValidatorNF.valid_loop_transformer_nf (src/pipeline/ValidatorNF.py:137-142): amax over the map."""
    return amap.amax(dim=(1, 2, 3))


def tokens_to_nchw(tokens: Tensor) -> Tensor:
    """ValidatorNF.py:130-134: [B,P,C] → [B,C,g,g]."""
    B, P, C = tokens.shape
    g = int(math.sqrt(P))
    return tokens.transpose(2, 1).reshape(B, C, g, g)


# ----------------------------------------------------------------------------------------------
# Reconstruction head scoring tail  (src/classes/CnnAutoEncoder.py:49,68-74; ValidatorRecon.py:109-116)
# ----------------------------------------------------------------------------------------------
def small_decoder_forward(sd: dict, z: Tensor, prefix: str = "decoder.", fmap: int = 7, eps: float = 1e-5) -> Tensor:
    """DecoderVanillaCNN.forward in eval mode (src/classes/CnnDecoder.py:16-117): Linear-ReLU-Linear-ReLU,
    unflatten to [768,f,f], 5 x (ConvTranspose2d k3 s2 p1 op1, BatchNorm(running stats), ReLU), Tanh last."""
    p = prefix
    h = F.relu(z @ sd[p + "decoder_lin.0.weight"].t() + sd[p + "decoder_lin.0.bias"])
    h = F.relu(h @ sd[p + "decoder_lin.2.weight"].t() + sd[p + "decoder_lin.2.bias"])
    h = h.reshape(z.shape[0], 768, fmap, fmap)
    for i in range(5):
        c, bn = f"{p}decoder_cnn.{3 * i}.", f"{p}decoder_cnn.{3 * i + 1}."
        h = F.conv_transpose2d(h, sd[c + "weight"], sd[c + "bias"], stride=2, padding=1, output_padding=1)
        h = (h - sd[bn + "running_mean"].view(1, -1, 1, 1)) / torch.sqrt(sd[bn + "running_var"].view(1, -1, 1, 1) + eps)
        h = h * sd[bn + "weight"].view(1, -1, 1, 1) + sd[bn + "bias"].view(1, -1, 1, 1)
        h = torch.tanh(h) if i == 4 else F.relu(h)
    return h


def _bn_eval(h: Tensor, sd: dict, name: str, eps: float = 1e-5) -> Tensor:
    """nn.BatchNorm2d in eval mode (running statistics)."""
    h = (h - sd[name + "running_mean"].view(1, -1, 1, 1)) / torch.sqrt(sd[name + "running_var"].view(1, -1, 1, 1) + eps)
    return h * sd[name + "weight"].view(1, -1, 1, 1) + sd[name + "bias"].view(1, -1, 1, 1)


RESNET_DECODER_LAYERS = (("layer4", 3, 2), ("layer3", 4, 2), ("layer2", 6, 2), ("layer1", 3, 1))  # name, blocks, last stride


def resnet_decoder_forward(sd: dict, z: Tensor, prefix: str = "decoder.") -> Tensor:
    """DecoderResNetVariableEmbeddingSize.forward in eval mode (src/classes/CnnDecoder.py:158-196) over
    ReverseResNet._forward_cnns_only (src/classes/resnet/ReverseResNet.py:228-235) and Bottleneck.forward (:86-103):
    fc1/fc2 + ReLU, unflatten to [2048,1,1], nearest upsample to 7x7, layer4..layer1 (the last block of a layer carries
    the stride-2 conv2 and the `upsample` identity path, :186-209), nearest upsample to 112, de_conv1 (k7 s2 p3 op1),
    bn1, tanh."""
    p = prefix
    h = F.relu(z @ sd[p + "fc1.0.weight"].t() + sd[p + "fc1.0.bias"])
    h = F.relu(h @ sd[p + "fc2.0.weight"].t() + sd[p + "fc2.0.bias"])
    h = h.reshape(z.shape[0], -1, 1, 1)
    h = F.interpolate(h, size=7, mode="nearest")
    for layer, blocks, last_stride in RESNET_DECODER_LAYERS:
        for i in range(blocks):
            b = f"{p}{layer}.{i}."
            last = i == blocks - 1
            stride = last_stride if last else 1
            op = stride - 1  # output_padding 1 with stride 2 (:180), 0 for layer1's last block (:141-143)
            out = F.relu(_bn_eval(F.conv_transpose2d(h, sd[b + "conv3.weight"]), sd, b + "bn3."))
            out = F.conv_transpose2d(out, sd[b + "conv2.weight"], stride=stride, padding=1, output_padding=op)
            out = F.relu(_bn_eval(out, sd, b + "bn2."))
            out = _bn_eval(F.conv_transpose2d(out, sd[b + "conv1.weight"]), sd, b + "bn1.")
            identity = h
            if last:
                identity = F.conv_transpose2d(h, sd[b + "upsample.0.weight"], stride=stride, output_padding=op)
                identity = _bn_eval(identity, sd, b + "upsample.1.")
            h = F.relu(out + identity)
    h = F.interpolate(h, size=112, mode="nearest")
    h = F.conv_transpose2d(h, sd[p + "de_conv1.weight"], stride=2, padding=3, output_padding=1)
    return torch.tanh(_bn_eval(h, sd, p + "bn1."))


def recon_l2_scores(recon: Tensor, images: Tensor):
    """MSELoss(reduction='none') → mean over channels (keepdim) → amax per image."""
    amap = ((recon - images) ** 2).mean(dim=1, keepdim=True)
    return amap.amax(dim=(1, 2, 3)), amap
