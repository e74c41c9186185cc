"""GPU parity of the EsViT Swin-T encoder (windowed attention with relative-position bias and shift mask,
patch merging) against the reference-generated golden fixture and the CPU oracle, and of config 3
(EsViT + GMM head with 130 Gaussians)."""
import numpy as np
import pytest
import torch

from helpers import golden, gumbel

pytestmark = pytest.mark.gpu


def test_window_attention_matches_torch():
    """Head dim 32, 4 shifted windows of 196 tokens, dense bias + region mask, output in token order."""
    from vitad import ops
    from vitad.encoders import _window_maps

    B, H, heads, ws, shift = 2, 28, 6, 14, 7
    T, nW = ws * ws, (H // ws) ** 2
    g = torch.Generator().manual_seed(0)
    q = (torch.randn(B * nW, heads, T, 32, generator=g) * 0.5).half()
    k = torch.randn(B * nW, heads, T, 32, generator=g).half()
    v = torch.randn(B * nW, heads, T, 32, generator=g).half()
    bias = torch.randn(heads, T, T, generator=g)
    t2w, w2t, region = _window_maps(H, ws, shift)
    vt = torch.zeros(B * nW, heads, 32, 256, dtype=torch.float16)
    vt[..., :T] = v.transpose(-1, -2)
    out = ops.attention(q.cuda(), k.cuda(), vt.cuda(), T, windows=nW, bias=bias.transpose(1, 2).contiguous().cuda(),
                        region=region.cuda(), win2tok=w2t.cuda())  # bias is passed key-major [H, key, query]
    torch.cuda.synchronize()
    reg = region.view(nW, T).float()
    mask = (reg.unsqueeze(1) != reg.unsqueeze(2)).float() * -100.0  # [nW,T,T]
    s = q.float() @ k.float().transpose(-1, -2) + bias.unsqueeze(0) + mask.repeat(B, 1, 1).unsqueeze(1)
    o = (torch.softmax(s, -1) @ v.float()).transpose(1, 2).reshape(B, nW * T, heads * 32)  # window-major rows
    ref = torch.empty_like(o)
    ref[:, w2t.long()] = o  # window position -> token
    err = (out.cpu().float().view(B, nW * T, heads * 32) - ref).abs().max().item()
    assert err <= 4e-3, err


@pytest.mark.parametrize("tag,stress", [("default", False), ("stress", True)])
def test_esvit_forward_matches_reference_golden(tag, stress):
    from oracle import weights as W
    from vitad.encoders import EncoderEsVit

    g = golden("esvit_b2")
    enc = EncoderEsVit(224, requires_grad=True)
    enc.load_state_dict(W.make_esvit_state_dict(seed=51, stress=stress), strict=True)
    enc = enc.cuda().eval()
    with torch.no_grad():
        o = enc(W.synthetic_images(seed=9, batch=2).cuda())
    torch.cuda.synchronize()
    tok, lat = o.patch_embedding.cpu().numpy(), o.latent_space.cpu().numpy()
    assert tok.shape == (2, 49, 768) and lat.shape == (2, 768)
    assert np.abs(tok[:, ::3] - g[f"{tag}_tokens"]).max() <= 1.5e-2
    assert np.sqrt(np.mean((tok[:, ::3] - g[f"{tag}_tokens"]) ** 2)) <= 2.5e-3
    assert np.abs(lat - g[f"{tag}_latent"]).max() <= 5e-3


def test_esvit_gmm130_scores_match_oracle():
    """Config 3 shape: EsViT features (P = 49) + GMM head with 130 Gaussians, ragged batch of 3."""
    from oracle import vitad_oracle as O
    from oracle import weights as W
    from vitad import ops
    from vitad.encoders import EncoderEsVit
    from vitad.mdn import GaussianMixtureDensityNetwork

    enc_sd = W.make_esvit_state_dict(seed=51, stress=True)
    mdn_sd = W.make_mdn_state_dict(seed=23, num_gaussians=130, stress=True)
    imgs = W.synthetic_images(seed=12, batch=3)
    gn = gumbel((3, 49, 130), 900)
    enc = EncoderEsVit(224, requires_grad=True)
    enc.load_state_dict(enc_sd)
    head = GaussianMixtureDensityNetwork(768, 768, 130)
    head.load_state_dict(mdn_sd)
    enc, head = enc.cuda().eval(), head.cuda().eval()
    with torch.no_grad():
        f = enc(imgs.cuda())
        prob, scores = head.score(f.patch_embedding, gn.cuda())
        maps, _ = ops.bilinear_up(prob.view(-1, 7, 7), 224, align_corners=True, post_one_minus=True)
        tok, _ = O.swin_forward(enc_sd, imgs)
        rs, rm = O.mdn_scores(O.mdn_probability_map(O.mdn_patch_loglik(tok, mdn_sd, gn)), 224, 32)
    torch.cuda.synchronize()
    from helpers import assert_map_parity, assert_rel

    assert_rel(scores.cpu().numpy(), rs.numpy(), 1e-3, what="EsViT + GMM image scores")
    assert_map_parity(maps.cpu().numpy(), rm.numpy(), what="EsViT + GMM maps")
