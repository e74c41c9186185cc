"""Reconstruction path (config 4): the oracle and the CUDA drop-in against ValidatorRecon.valid_loop_mse of the
reference (golden fixture recon_validator.npz: ae_deit_small = DeiT + small CNN decoder, B=2)."""
import numpy as np
import pytest
import torch

from helpers import golden
from oracle import vitad_oracle as O
from oracle import weights as W


def _weights():
    sd = {("encoder." + k): v for k, v in W.make_deit_state_dict(seed=11, stress=True).items()}
    sd.update(W.make_small_decoder_state_dict(seed=41))
    return sd


def test_recon_oracle_matches_reference_golden():
    g = golden("recon_validator")
    sd = _weights()
    imgs = W.synthetic_images(seed=8, batch=2)
    with torch.no_grad():
        _, cls = O.deit_forward(sd, imgs, prefix="encoder.deit.")
        recon = O.small_decoder_forward(sd, cls)
        score, amap = O.recon_l2_scores(recon, imgs)
    # five BatchNorms with running_var ~0.015 amplify fp32 rounding differences ~8x per layer
    np.testing.assert_allclose(recon.numpy()[:, :, ::8, ::8], g["recons_sub"], rtol=0, atol=3e-4)
    np.testing.assert_allclose(score.numpy(), g["image_scores"], rtol=1e-4)
    np.testing.assert_allclose(amap.numpy()[:, :, ::8, ::8], g["pixel_scores_sub"], rtol=1e-3, atol=1e-6)


def test_get_model_names_and_state_dict_keys():
    from vitad.model_helper import get_model, get_possible_models

    assert {"enc_deit", "ae_deit", "ae_deit_small"} <= set(get_possible_models())
    assert get_model("nope") is None
    model = get_model("ae_deit_small", 224)
    assert set(model.state_dict().keys()) == set(_weights().keys())
    assert model.architecture == "transformer" and type(model.decoder).__name__ == "DecoderVanillaCNN"
    big = get_model("ae_deit", 224)  # reverse-ResNet decoder (the reference's default)
    assert sorted(big.state_dict().keys()) == sorted(str(k) for k in golden("recon_validator_resnet")["state_dict_keys"])


@pytest.mark.gpu
def test_recon_validator_matches_reference_golden():
    from vitad.model_helper import get_model
    from vitad.validators import ValidatorRecon

    g = golden("recon_validator")
    model = get_model("ae_deit_small", 224)
    imgs = W.synthetic_images(seed=8, batch=2)
    batches = [(imgs, torch.zeros(2, 1, 224, 224), torch.tensor([0, 1]))]
    props = {"dataset": "synthetic", "dataclass": "x", "fp_thres": 0.3}
    val = ValidatorRecon(model, None, props, weights_object=_weights())
    res = val.valid_loop_mse(batches)
    assert set(res) >= {"image_scores", "pixel_scores", "image_labels", "pixel_labels", "origs", "recons"}
    assert res["pixel_scores"].shape == (2, 1, 224, 224) and res["recons"].shape == (2, 3, 224, 224)
    # the cls token comes from the fp16-operand encoder; the decoder amplifies it through 7 layers
    assert np.abs(res["recons"][:, :, ::8, ::8] - g["recons_sub"]).max() <= 2e-3
    from helpers import assert_map_parity, assert_rel

    assert_rel(res["image_scores"], g["image_scores"], 1e-3, what="recon (small decoder) image scores")
    assert_map_parity(res["pixel_scores"][:, :, ::8, ::8], g["pixel_scores_sub"], what="recon (small decoder) L2 maps")


@pytest.mark.gpu
def test_recon_validator_resnet_matches_reference_golden():
    """get_model('ae_deit') (DeiT + reverse-ResNet decoder) through ValidatorRecon.valid_loop_mse against the reference's
    own run (oracle/make_golden.py case_recon_validator_resnet).  53 GEMM layers sit between the cls token and the image;
    the decoder runs them in split-fp16 arithmetic (tests/test_resnet_decoder.py), so what is left is the encoder's
    fp16-operand error on the cls token, amplified through the decoder."""
    from vitad.model_helper import get_model
    from vitad.validators import ValidatorRecon

    g = golden("recon_validator_resnet")
    model = get_model("ae_deit", 224)
    sd = {("encoder." + k): v for k, v in W.make_deit_state_dict(seed=11, stress=True).items()}
    sd.update(W.make_resnet_decoder_state_dict(seed=43))
    imgs = W.synthetic_images(seed=8, batch=2)
    batches = [(imgs, torch.zeros(2, 1, 224, 224), torch.tensor([0, 1]))]
    props = {"dataset": "synthetic", "dataclass": "x", "fp_thres": 0.3}
    val = ValidatorRecon(model, None, props, weights_object=sd)
    res = val.valid_loop_mse(batches)
    assert res["pixel_scores"].shape == (2, 1, 224, 224) and res["recons"].shape == (2, 3, 224, 224)
    from helpers import assert_map_parity, assert_rel

    assert np.abs(res["recons"][:, :, ::8, ::8] - g["recons_sub"]).max() <= 2e-3
    assert_rel(res["image_scores"], g["image_scores"], 1e-3, what="recon (reverse-ResNet decoder) image scores")
    assert_map_parity(res["pixel_scores"][:, :, ::8, ::8], g["pixel_scores_sub"], what="recon (reverse-ResNet decoder) L2 maps")
    np.testing.assert_allclose(res["pixel_scores"].sum(axis=(1, 2, 3)), g["pixel_scores_sum"], rtol=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 5, 32])
def test_cnn_decoder_matches_fp32_modules(B):
    """the CUDA decoder (two Linear GEMMs, four ConvTranspose2d+BN+ReLU layers as one phase-GEMM each,
    fused last layer) against the reference module stack evaluated in fp32 on the CPU (CnnDecoder.py:16-117, eval mode:
    BatchNorm running statistics), on the fixture's decoder weights (running_var ~ 0.015: every layer amplifies)."""
    from vitad.autoencoders import DecoderVanillaCNN

    dec = DecoderVanillaCNN(z_space=768, first_feature_map_size=7)
    dec.load_state_dict({k[len("decoder."):]: v for k, v in W.make_small_decoder_state_dict(seed=41).items()})
    dec.eval()
    lat = torch.randn(B, 768, generator=torch.Generator().manual_seed(B)) * 0.7
    with torch.no_grad():
        ref = dec.decoder_cnn(dec.unflatten(dec.decoder_lin(lat)))  # parameter containers, fp32 on the CPU
        got = dec.cuda()(lat.cuda())
    torch.cuda.synchronize()
    assert got.shape == (B, 3, 224, 224)
    err = (got.cpu() - ref).abs().max().item()
    assert err <= 1e-3 * max(ref.abs().max().item(), 1e-3) + 2e-4, (err, ref.abs().max().item())


@pytest.mark.gpu
def test_recon_image_auroc_identical_to_4_decimals_on_synthetic_anomaly_set():
    """north_star's AUROC criterion for the reconstruction path (config 4, reverse-ResNet decoder): ValidatorRecon over the
    whole designed set (16 images, oracle scores >= 10x the allowed noise apart — asserted) against the oracle (DeiT cls
    token -> decoder -> per-pixel L2 -> amax)."""
    from sklearn.metrics import roc_auc_score

    from helpers import DESIGNED_RECON, assert_designed_separation, assert_map_parity, assert_rel
    from vitad.model_helper import get_model
    from vitad.synthetic import batches, make_designed_set
    from vitad.validators import ValidatorRecon

    images, labels, masks = make_designed_set(DESIGNED_RECON)
    sd = {("encoder." + k): v for k, v in W.make_deit_state_dict(seed=11, stress=True).items()}
    sd.update(W.make_resnet_decoder_state_dict(seed=43))
    with torch.no_grad():
        _, cls = O.deit_forward(sd, images, prefix="encoder.deit.")
        ref_scores, ref_maps = O.recon_l2_scores(O.resnet_decoder_forward(sd, cls), images)
    ref_scores = ref_scores.numpy()
    assert_designed_separation(ref_scores, labels.numpy(), factor=10.0)
    props = {"dataset": "synthetic", "dataclass": "designed", "fp_thres": 0.3}
    val = ValidatorRecon(get_model("ae_deit", 224), None, props, weights_object=sd)
    res = val.valid_loop_mse(batches(images, labels, masks, batch_size=8))
    assert_rel(res["image_scores"], ref_scores, 1e-3, what="recon image scores")
    assert_map_parity(res["pixel_scores"], ref_maps.numpy(), what="recon L2 maps")
    assert np.array_equal(np.argsort(ref_scores), np.argsort(res["image_scores"]))
    assert round(roc_auc_score(res["image_labels"], res["image_scores"]), 4) == round(roc_auc_score(labels.numpy(), ref_scores), 4)
