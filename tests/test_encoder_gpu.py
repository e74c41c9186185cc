"""GPU parity of the encoder kernels and of EncoderDeit.forward against the CPU oracle and the
reference-generated golden fixtures.  Tolerances: the CUDA path rounds GEMM operands to fp16 (fp32
accumulate, fp32 residual stream); the numbers below are ~4x the error predicted by
tools/numerics_study.py for that design."""
import numpy as np
import pytest
import torch

from helpers import golden, gumbel

pytestmark = pytest.mark.gpu


def test_layernorm_matches_torch():
    from vitad import ops

    g = torch.Generator().manual_seed(0)
    x = (torch.randn(6336, 768, generator=g) * 3 + 0.5).cuda()
    w = (1 + 0.1 * torch.randn(768, generator=g)).cuda()
    b = (0.1 * torch.randn(768, generator=g)).cuda()
    ref = torch.nn.functional.layer_norm(x, (768,), w, b, 1e-6)
    o16 = torch.empty(6336, 768, device="cuda", dtype=torch.float16)
    o32 = torch.empty(6336, 768, device="cuda")
    ops.layernorm(x, w, b, 1e-6, out_f16=o16, out_f32=o32)
    torch.cuda.synchronize()
    assert (o32 - ref).abs().max().item() <= 2e-5
    assert (o16.float() - ref).abs().max().item() <= 4e-3


def test_layernorm_token_remap_and_augmentation():
    from vitad import ops

    B, T, P = 3, 198, 196
    x = torch.randn(B * T, 768, device="cuda")
    w = torch.ones(768, device="cuda")
    b = torch.zeros(768, device="cuda")
    o16 = torch.full((B * P, 784), 7.0, device="cuda", dtype=torch.float16)
    o32 = torch.empty(B * P, 768, device="cuda")
    ops.layernorm(x, w, b, 1e-6, out_f16=o16, out_f32=o32, in_tokens=T, out_tokens=P, skip=2, aug_ones=2)
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x, (768,), w, b, 1e-6).view(B, T, 768)[:, 2:].reshape(B * P, 768)
    assert (o32 - ref).abs().max().item() <= 2e-5
    assert (o16[:, :768].float() - ref).abs().max().item() <= 4e-3
    assert (o16[:, 768:770] == 1).all() and (o16[:, 770:] == 0).all()


@pytest.mark.parametrize("C,rows", [(96, 100352), (96, 37), (192, 25088), (192, 5)])
def test_layernorm_narrow_rows(C, rows):
    """Swin stages 0/1: the sub-warp kernel (8 / 16 lanes per row), fp16 + fp32 outputs, ragged row counts, and the
    in-place fp32 form the patch-embed norm uses (SwinTransformerModule.py:645-655)."""
    from vitad import ops

    g = torch.Generator().manual_seed(C + rows)
    x = (torch.randn(rows, C, generator=g) * 2 + 0.3).cuda()
    w = (1 + 0.1 * torch.randn(C, generator=g)).cuda()
    b = (0.1 * torch.randn(C, generator=g)).cuda()
    ref = torch.nn.functional.layer_norm(x, (C,), w, b, 1e-5)
    o16 = torch.empty(rows, C, device="cuda", dtype=torch.float16)
    o32 = torch.empty(rows, C, device="cuda")
    ops.layernorm(x, w, b, 1e-5, out_f16=o16, out_f32=o32)
    xin = x.clone()
    ops.layernorm(xin, w, b, 1e-5, out_f32=xin)  # in place
    torch.cuda.synchronize()
    assert (o32 - ref).abs().max().item() <= 2e-5
    assert (o16.float() - ref).abs().max().item() <= 4e-3
    assert torch.equal(xin, o32)


def test_patchify_matches_unfold():
    from vitad import ops

    img = torch.rand(3, 3, 224, 224, device="cuda")
    out = ops.patchify(img, 16)
    ref = torch.nn.functional.unfold(img, kernel_size=16, stride=16).transpose(1, 2).reshape(3 * 196, 768)
    torch.cuda.synchronize()
    assert (out.float() - ref).abs().max().item() <= 5e-4  # fp16 rounding of values in [0,1]


@pytest.mark.parametrize("B,T", [(2, 198), (1, 198), (3, 130), (2, 49)])
def test_attention_matches_torch(B, T):
    from vitad import ops

    H = 12
    g = torch.Generator().manual_seed(T)
    q = (torch.randn(B, H, T, 64, generator=g) * 0.5).half().cuda()
    k = (torch.randn(B, H, T, 64, generator=g) * 1.0).half().cuda()
    v = torch.randn(B, H, T, 64, generator=g).half().cuda()
    vt = torch.zeros(B, H, 64, 256, device="cuda", dtype=torch.float16)
    vt[..., :T] = v.transpose(-1, -2)
    out = ops.attention(q, k, vt, T)
    torch.cuda.synchronize()
    attn = torch.softmax(q.float() @ k.float().transpose(-1, -2), dim=-1)
    ref = (attn @ v.float()).transpose(1, 2).reshape(B * T, H * 64)
    err = (out.float() - ref).abs().max().item()
    assert err <= 4e-3, err
    # V in its natural layout (MN-major tensor-core operand): the same products in the same order -> the same bits
    out_nat = ops.attention(q, k, None, T, v=v.contiguous())
    torch.cuda.synchronize()
    assert torch.equal(out_nat, out), (out_nat.float() - out.float()).abs().max().item()


def _run_deit(sd, images, block_index=0):
    from vitad.encoders import EncoderDeit

    enc = EncoderDeit(224)
    enc.load_state_dict(sd, strict=True)
    enc = enc.cuda().eval()
    with torch.no_grad():
        o = enc(images.cuda(), block_index=block_index)
    torch.cuda.synchronize()
    return o.patch_embedding.cpu(), o.latent_space.cpu(), o.patch_embedding._vitad_xaug[0].cpu()


@pytest.mark.parametrize("tag,stress", [("default", False), ("stress", True)])
@pytest.mark.parametrize("block_index", [0, 7])
def test_deit_forward_matches_reference_golden(tag, stress, block_index):
    from oracle import weights as W

    g = golden("deit_b2")
    sd = W.make_deit_state_dict(seed=11, stress=stress)
    tok, cls, xaug = _run_deit(sd, W.synthetic_images(seed=3, batch=2), block_index)
    k = f"{tag}_b{block_index}_"
    # tokens are LayerNorm outputs (unit scale): absolute tolerance
    assert np.abs(tok[:, ::14].numpy() - g[k + "tokens_sub"]).max() <= 1.5e-2
    assert np.sqrt(np.mean((tok[:, ::14].numpy() - g[k + "tokens_sub"]) ** 2)) <= 2.5e-3
    assert np.abs(cls.numpy() - g[k + "cls"]).max() <= 1.5e-2
    assert np.abs(xaug[:, :768].float().view(2, 196, 768)[:, ::14].numpy() - g[k + "tokens_sub"]).max() <= 2e-2
    assert (xaug[:, 768:770] == 1).all() and (xaug[:, 770:] == 0).all()


@pytest.mark.parametrize("tag,stress", [("default", False), ("stress", True)])
def test_vit_forward_matches_reference_golden(tag, stress):
    """EncoderVit (197 tokens, one prefix token) through the same CUDA forward; block_index is ignored as in the
    reference (TransformerEncoder.py:196-208); state_dict keys vit.* load strictly."""
    from oracle import weights as W
    from vitad.encoders import EncoderVit
    from vitad.model_helper import get_model

    g = golden("vit_b2")
    enc = get_model("enc_vit", 224, requires_grad=True)
    assert isinstance(enc, EncoderVit) and enc.num_embedded_patches == 196 and enc.size_patch_embedding == 768
    enc.load_state_dict(W.make_vit_state_dict(seed=13, stress=stress), strict=True)
    enc = enc.cuda().eval()
    x = W.synthetic_images(seed=4, batch=2).cuda()
    with torch.no_grad():
        o = enc(x)
        o7 = enc(x, block_index=7)
    torch.cuda.synchronize()
    tok, cls = o.patch_embedding.cpu(), o.latent_space.cpu()
    assert tok.shape == (2, 196, 768) and cls.shape == (2, 768)
    assert np.abs(tok[:, ::14].numpy() - g[f"{tag}_tokens_sub"]).max() <= 1.5e-2
    assert np.sqrt(np.mean((tok[:, ::14].numpy() - g[f"{tag}_tokens_sub"]) ** 2)) <= 2.5e-3
    assert np.abs(cls.numpy() - g[f"{tag}_cls"]).max() <= 1.5e-2
    assert torch.equal(o7.patch_embedding.cpu(), tok)


def test_deit_forward_from_uint8_images_equals_float_path():
    """uint8 pixels in, /255 (ToTensor, GeneralDataset.py:46-53) folded into the patch gather — the
    result is bit-identical to feeding the fp32 tensor ToTensor would have produced."""
    from oracle import weights as W
    from vitad.encoders import EncoderDeit

    enc = EncoderDeit(224)
    enc.load_state_dict(W.make_deit_state_dict(seed=11, stress=True))
    enc = enc.cuda().eval()
    u8 = torch.randint(0, 256, (3, 3, 224, 224), dtype=torch.uint8, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        a = enc(u8.cuda())
        b = enc((u8.float() / 255.0).cuda())
    torch.cuda.synchronize()
    assert torch.equal(a.patch_embedding, b.patch_embedding) and torch.equal(a.latent_space, b.latent_space)


def test_deit_forward_matches_oracle_batch_sizes():
    """Ragged batch sizes (the reference loader has no drop_last): B=1 and B=5 against the CPU oracle."""
    from oracle import vitad_oracle as O
    from oracle import weights as W

    sd = W.make_deit_state_dict(seed=12, stress=True)
    for B in (1, 5):
        imgs = W.synthetic_images(seed=20 + B, batch=B)
        tok, cls, _ = _run_deit(sd, imgs)
        with torch.no_grad():
            rt, rc = O.deit_forward(sd, imgs)
        assert (tok - rt).abs().max().item() <= 1.5e-2
        assert (tok - rt).pow(2).mean().sqrt().item() <= 2.5e-3
        assert (cls - rc).abs().max().item() <= 1.5e-2


def test_encoder_rejects_cpu_input():
    from vitad.encoders import EncoderDeit

    with pytest.raises(RuntimeError):
        EncoderDeit(224)(torch.rand(1, 3, 224, 224))


@pytest.mark.parametrize("B", [3, 33, 64])
def test_scoring_is_batch_invariant(B):
    """Every per-image quantity (tokens, cls, per-patch log-likelihood) must not depend on which other images share the
    batch: the tile schedule (full waves + cut tail, 4-CTA clusters, ragged last row block) changes with M, the
    per-element accumulation order does not — results are bit-identical between one batch of B and single images /
    sub-batches.  Guards the scheduling code at sizes other than the benchmark's 32."""
    from oracle import weights as W
    from vitad.encoders import EncoderDeit
    from vitad.mdn import GaussianMixtureDensityNetwork

    enc = EncoderDeit(224)
    enc.load_state_dict(W.make_deit_state_dict(seed=11, stress=True))
    head = GaussianMixtureDensityNetwork(768, 768, 100)
    head.load_state_dict(W.make_mdn_state_dict(seed=21, num_gaussians=100, stress=True))
    enc, head = enc.cuda().eval(), head.cuda().eval()
    imgs = W.synthetic_images(seed=60 + B, batch=B).cuda()
    gn = gumbel((B, 196, 100), 17).cuda()
    with torch.no_grad():
        full = enc(imgs)
        L_full = head.patch_log_likelihood(full.patch_embedding, gn)
        cut = B // 2 + 1
        parts = [(0, 1), (1, cut), (cut, B)]
        for lo, hi in parts:
            sub = enc(imgs[lo:hi].contiguous())
            assert torch.equal(sub.patch_embedding, full.patch_embedding[lo:hi]), (B, lo, hi)
            assert torch.equal(sub.latent_space, full.latent_space[lo:hi])
            L_sub = head.patch_log_likelihood(sub.patch_embedding, gn[lo:hi].contiguous())
            assert torch.equal(L_sub, L_full[lo:hi]), (B, lo, hi)
    assert torch.isfinite(L_full).all()
