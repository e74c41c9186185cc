"""CPU tests: the oracle (oracle/vitad_oracle.py) against the golden fixtures produced by the reference's
own classes (oracle/make_golden.py), an independent DeiT cross-check against transformers.DeiTModel, and
the explicit bilinear restatement against torch.nn.functional.interpolate."""
import numpy as np
import pytest
import torch

from helpers import golden, gumbel, rel_err
from oracle import vitad_oracle as O
from oracle import weights as W


@pytest.fixture(scope="module")
def deit_stress():
    return W.make_deit_state_dict(seed=11, stress=True)


@pytest.mark.parametrize("tag,stress", [("default", False), ("stress", True)])
def test_vit_oracle_matches_reference_golden(tag, stress):
    """EncoderVit (timm vit_base_patch16_224, one prefix token) — fixture from the reference class, oracle/make_golden.py."""
    g = golden("vit_b2")
    sd = W.make_vit_state_dict(seed=13, stress=stress)
    with torch.no_grad():
        tok, cls = O.vit_forward(sd, W.synthetic_images(seed=4, batch=2))
    np.testing.assert_allclose(tok[:, ::14].numpy(), g[f"{tag}_tokens_sub"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(tok.sum(-1).numpy(), g[f"{tag}_token_sum"], rtol=0, atol=5e-3)
    np.testing.assert_allclose(cls.numpy(), g[f"{tag}_cls"], rtol=0, atol=2e-4)


@pytest.mark.parametrize("tag,stress", [("default", False), ("stress", True)])
@pytest.mark.parametrize("block_index", [0, 7])
def test_deit_oracle_matches_reference_golden(tag, stress, block_index):
    g = golden("deit_b2")
    sd = W.make_deit_state_dict(seed=11, stress=stress)
    x = W.synthetic_images(seed=3, batch=2)
    with torch.no_grad():
        tok, cls = O.deit_forward(sd, x, block_index=block_index)
    k = f"{tag}_b{block_index}_"
    np.testing.assert_allclose(tok[:, ::14].numpy(), g[k + "tokens_sub"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(tok.sum(-1).numpy(), g[k + "token_sum"], rtol=0, atol=5e-3)
    np.testing.assert_allclose(tok.abs().sum(-1).numpy(), g[k + "token_abs"], rtol=1e-5, atol=0)
    np.testing.assert_allclose(cls.numpy(), g[k + "cls"], rtol=0, atol=2e-4)


def test_deit_restatement_matches_hf_deit(deit_stress):
    """Independent check of the un-vendored timm arithmetic: same weights in transformers.DeiTModel."""
    from transformers import DeiTConfig, DeiTModel

    cfg = DeiTConfig(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                     layer_norm_eps=1e-6, image_size=224, patch_size=16, hidden_act="gelu",
                     hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    hf = DeiTModel(cfg, add_pooling_layer=False).eval()
    sd = {k[len("deit."):]: v for k, v in deit_stress.items()}
    m = {"embeddings.cls_token": sd["cls_token"], "embeddings.distillation_token": sd["dist_token"],
         "embeddings.position_embeddings": sd["pos_embed"],
         "embeddings.patch_embeddings.projection.weight": sd["patch_embed.proj.weight"],
         "embeddings.patch_embeddings.projection.bias": sd["patch_embed.proj.bias"],
         "layernorm.weight": sd["norm.weight"], "layernorm.bias": sd["norm.bias"]}
    for i in range(12):
        t, h = f"blocks.{i}.", f"encoder.layer.{i}."
        qw, qb = sd[t + "attn.qkv.weight"], sd[t + "attn.qkv.bias"]
        for j, nm in enumerate(("query", "key", "value")):
            m[h + f"attention.attention.{nm}.weight"] = qw[768 * j:768 * (j + 1)]
            m[h + f"attention.attention.{nm}.bias"] = qb[768 * j:768 * (j + 1)]
        m[h + "attention.output.dense.weight"] = sd[t + "attn.proj.weight"]
        m[h + "attention.output.dense.bias"] = sd[t + "attn.proj.bias"]
        m[h + "layernorm_before.weight"], m[h + "layernorm_before.bias"] = sd[t + "norm1.weight"], sd[t + "norm1.bias"]
        m[h + "layernorm_after.weight"], m[h + "layernorm_after.bias"] = sd[t + "norm2.weight"], sd[t + "norm2.bias"]
        m[h + "intermediate.dense.weight"], m[h + "intermediate.dense.bias"] = sd[t + "mlp.fc1.weight"], sd[t + "mlp.fc1.bias"]
        m[h + "output.dense.weight"], m[h + "output.dense.bias"] = sd[t + "mlp.fc2.weight"], sd[t + "mlp.fc2.bias"]
    missing, unexpected = hf.load_state_dict(m, strict=False)
    assert not unexpected and all("mask_token" in k for k in missing), (missing, unexpected)
    x = W.synthetic_images(seed=3, batch=1)
    with torch.no_grad():
        ref = hf(pixel_values=x).last_hidden_state
        tok, cls = O.deit_forward(deit_stress, x)
    np.testing.assert_allclose(tok.numpy(), ref[:, 2:].numpy(), rtol=0, atol=3e-4)
    np.testing.assert_allclose(cls.numpy(), ref[:, 0].numpy(), rtol=0, atol=3e-4)


@pytest.mark.parametrize("tag,stress", [("default", False), ("stress", True)])
def test_gmm_validator_oracle_matches_reference_golden(tag, stress, deit_stress):
    g = golden("gmm_validator_k100")
    mdn = W.make_mdn_state_dict(seed=21, num_gaussians=100, stress=stress)
    imgs = W.synthetic_images(seed=5, batch=3)
    scores, maps = [], []
    with torch.no_grad():
        for call, (s, e) in enumerate(((0, 2), (2, 3))):
            tok, _ = O.deit_forward(deit_stress, imgs[s:e])
            L = O.mdn_patch_loglik(tok, mdn, gumbel((e - s, 196, 100), 700 + call))
            if call == 0:
                np.testing.assert_allclose(L.numpy(), g[f"{tag}_L_batch0"], rtol=0, atol=2e-4)
                np.testing.assert_allclose(O.mdn_probability_map(L).numpy(), g[f"{tag}_prob_batch0"], rtol=0, atol=2e-4)
            sc, mp = O.mdn_scores(O.mdn_probability_map(L), 224, 16)
            scores.append(sc)
            maps.append(mp)
    scores, maps = torch.cat(scores).numpy(), torch.cat(maps).numpy()
    np.testing.assert_allclose(scores, g[f"{tag}_image_scores"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(maps[:, :, ::8, ::8], g[f"{tag}_pixel_scores_sub"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(maps.sum(axis=(1, 2, 3)), g[f"{tag}_pixel_scores_sum"], rtol=2e-4)


@pytest.mark.parametrize("tag,stress", [("default", False), ("stress", True)])
def test_gmm_head_k130_oracle_matches_reference_golden(tag, stress):
    g = golden("gmm_head_k130_p49")
    sd = W.make_mdn_state_dict(seed=22, num_gaussians=130, stress=stress)
    x = torch.randn(3, 49, 768, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        L = O.mdn_patch_loglik(x, sd, gumbel((3, 49, 130), 800))
    np.testing.assert_allclose(L.numpy(), g[f"{tag}_L"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(O.mdn_probability_map(L).numpy(), g[f"{tag}_prob"], rtol=0, atol=2e-4)


@pytest.mark.parametrize("tag,stress", [("default", False), ("stress", True)])
def test_nf_validator_oracle_matches_reference_golden(tag, stress, deit_stress):
    g = golden("nf_validator")
    nf = W.make_nf_state_dict(seed=31, stress=stress)
    imgs = W.synthetic_images(seed=6, batch=2)
    with torch.no_grad():
        tok, _ = O.deit_forward(deit_stress, imgs)
        loss, amap, _, _ = O.nf_forward(nf, O.tokens_to_nchw(tok), flow_steps=20, img_size=224)
    np.testing.assert_allclose(O.nf_scores(amap).numpy(), g[f"{tag}_image_scores"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(amap.numpy()[:, :, ::8, ::8], g[f"{tag}_pixel_scores_sub"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(amap.sum(dim=(1, 2, 3)).numpy(), g[f"{tag}_pixel_scores_sum"], rtol=1e-4)
    np.testing.assert_allclose(loss.numpy(), g[f"{tag}_loss"], rtol=1e-4)


def test_recon_l2_oracle_matches_reference_golden():
    g = golden("recon_l2")
    gen = torch.Generator().manual_seed(4)
    images = torch.rand(3, 3, 224, 224, generator=gen)
    recon = torch.tanh(torch.randn(3, 3, 224, 224, generator=gen))
    score, amap = O.recon_l2_scores(recon, images)
    np.testing.assert_allclose(score.numpy(), g["image_scores"], rtol=1e-6)
    np.testing.assert_allclose(amap.numpy()[:, :, ::8, ::8], g["map_sub"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(amap.sum(dim=(1, 2, 3)).numpy(), g["map_sum"], rtol=1e-5)


@pytest.mark.parametrize("align", [True, False])
@pytest.mark.parametrize("g_in", [14, 7])
def test_bilinear_restatement_matches_torch(align, g_in):
    x = torch.rand(3, 1, g_in, g_in, generator=torch.Generator().manual_seed(1))
    ref = torch.nn.functional.interpolate(x, size=(224, 224), mode="bilinear", align_corners=align)
    np.testing.assert_allclose(O.bilinear_upsample(x, 224, align).numpy(), ref.numpy(), rtol=0, atol=1e-6)


@pytest.mark.parametrize("tag,stress", [("default", False), ("stress", True)])
def test_esvit_oracle_matches_reference_golden(tag, stress):
    """Swin-T W14 restatement vs the reference's vendored SwinTransformerModule (eval mode)."""
    g = golden("esvit_b2")
    sd = W.make_esvit_state_dict(seed=51, stress=stress)
    with torch.no_grad():
        tok, latent = O.swin_forward(sd, W.synthetic_images(seed=9, batch=2))
    np.testing.assert_allclose(tok.numpy()[:, ::3], g[f"{tag}_tokens"], rtol=0, atol=3e-4)
    np.testing.assert_allclose(tok.sum(-1).numpy(), g[f"{tag}_token_sum"], rtol=0, atol=5e-3)
    np.testing.assert_allclose(latent.numpy(), g[f"{tag}_latent"], rtol=0, atol=3e-4)


def test_designed_anomaly_sets_are_separated_under_the_oracle():
    """The fixed sets of the AUROC parity tests (helpers.DESIGNED_*; tools/design_anomaly_sets.py) keep their property on
    this machine's CPU: the oracle's image scores of the WHOLE set are pairwise >= 10x the allowed numerical noise (1e-3
    of the largest score) apart, with both labels interleaved.  (The GPU tests assert the same before comparing.)"""
    from helpers import (DESIGNED_GMM, DESIGNED_NF, DESIGNED_RECON, GMM_NOISE_SEED_BASE, NF_TEST_GAIN,
                         assert_designed_separation)
    from oracle import vitad_oracle as O
    from oracle import weights as W
    from vitad.synthetic import make_designed_set

    enc_sd = W.make_deit_state_dict(seed=11, stress=True)
    with torch.no_grad():
        images, labels, _ = make_designed_set(DESIGNED_GMM)
        g = torch.stack([O.gumbel_noise((196, 100), torch.Generator().manual_seed(GMM_NOISE_SEED_BASE + s)) for s in DESIGNED_GMM])
        tok, _ = O.deit_forward(enc_sd, images)
        s, _ = O.mdn_scores(O.mdn_probability_map(O.mdn_patch_loglik(tok, W.make_mdn_state_dict(21, 100, stress=True), g)), 224, 16)
        assert_designed_separation(s.numpy(), labels.numpy())

        images, labels, _ = make_designed_set(DESIGNED_NF)
        tok, _ = O.deit_forward(enc_sd, images, block_index=0)
        nf_sd = W.make_nf_state_dict(seed=31, stress=True, subnet_gain=NF_TEST_GAIN)
        _, amap, _, _ = O.nf_forward(nf_sd, O.tokens_to_nchw(tok), flow_steps=20, img_size=224)
        assert_designed_separation(O.nf_scores(amap).numpy(), labels.numpy())

        images, labels, _ = make_designed_set(DESIGNED_RECON)
        sd = {("encoder." + k): v for k, v in enc_sd.items()}
        sd.update(W.make_resnet_decoder_state_dict(seed=43))
        _, cls = O.deit_forward(sd, images, prefix="encoder.deit.")
        sc, _ = O.recon_l2_scores(O.resnet_decoder_forward(sd, cls), images)
        assert_designed_separation(sc.numpy(), labels.numpy())
