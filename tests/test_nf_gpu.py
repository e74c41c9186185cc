"""GPU parity of the normalizing-flow head against the reference-generated golden fixture
(ValidatorNF.valid_loop_transformer_nf, DeiT stress weights + 20-step flow, B=2) and against the CPU oracle on
plain random features (head only, several batch sizes)."""
import numpy as np
import pytest
import torch

from helpers import golden

pytestmark = pytest.mark.gpu


def _flow(sd):
    from vitad.nf import NormalizingFlow

    np.random.seed(0)
    nf = NormalizingFlow(num_channels=768, img_size=224, num_patches=196, hidden_ratio=0.16, flow_steps=20)
    nf.load_state_dict(sd, strict=True)
    return nf.cuda().eval()


@pytest.mark.parametrize("stress,B", [(False, 1), (True, 1), (False, 3), (True, 3), (True, 32)])
def test_nf_head_matches_oracle(stress, B):
    from oracle import vitad_oracle as O
    from oracle import weights as W

    sd = W.make_nf_state_dict(seed=31, stress=stress)
    nf = _flow(sd)
    tokens = torch.randn(B, 196, 768, generator=torch.Generator().manual_seed(B))
    with torch.no_grad():
        loss, amap, _, _ = O.nf_forward(sd, O.tokens_to_nchw(tokens), flow_steps=20, img_size=224)
        r = nf.forward_tokens(tokens.cuda())
        r2 = nf(O.tokens_to_nchw(tokens).cuda())  # the reference's NCHW entry point
    torch.cuda.synchronize()
    ref = amap.numpy()
    from helpers import assert_map_parity, assert_rel

    assert_map_parity(r.anomaly_score_map.cpu().numpy(), ref, what="NF map (tokens)")
    assert_map_parity(r2.anomaly_score_map.cpu().numpy(), ref, what="NF map (NCHW)")
    assert_rel(r.image_max.cpu().numpy(), O.nf_scores(amap).numpy(), 1e-3, what="NF image max")
    assert abs(r.loss.item() - loss.item()) <= 2e-3 * abs(loss.item())


@pytest.mark.parametrize("tag,stress", [("default", False), ("stress", True)])
def test_nf_validator_matches_reference_golden(tag, stress):
    from oracle import weights as W
    from vitad.encoders import EncoderDeit
    from vitad.validators import ValidatorNF

    g = golden("nf_validator")
    enc = EncoderDeit(224)
    enc.load_state_dict(W.make_deit_state_dict(seed=11, stress=True))
    np.random.seed(0)
    from vitad.nf import NormalizingFlow

    nf = NormalizingFlow(768, 224, 196, hidden_ratio=0.16, flow_steps=20)
    imgs = W.synthetic_images(seed=6, batch=2)
    batches = [(imgs, torch.zeros(2, 1, 224, 224), torch.tensor([0, 1]))]
    props = {"dataset": "synthetic", "dataclass": "x", "fp_thres": 0.3}
    val = ValidatorNF([nf], enc, None, props, weights_object=[W.make_nf_state_dict(seed=31, stress=stress)])
    res = val.valid_loop_transformer_nf(batches)
    ref_s, ref_m = g[f"{tag}_image_scores"], g[f"{tag}_pixel_scores_sub"]
    assert res["pixel_scores"].shape == (2, 1, 224, 224)
    from helpers import assert_map_parity, assert_rel

    assert_rel(res["image_scores"], ref_s, 1e-3, what="NF image scores")
    assert_map_parity(res["pixel_scores"][:, :, ::8, ::8], ref_m, what="NF maps")
    np.testing.assert_allclose(res["pixel_scores"].sum(axis=(1, 2, 3)), g[f"{tag}_pixel_scores_sum"], rtol=2e-3)


def test_nf_image_auroc_identical_to_4_decimals_on_synthetic_anomaly_set():
    """north_star's AUROC criterion for the NF head (config 2): ValidatorNF over the whole designed set against the oracle
    (DeiT last-block features -> 20-step flow -> amax of the bilinear map).  NF image scores = max over pixels of
    1 - exp(-mean_c z^2 / 2) live in a narrow band (random-init flows map every image to similar |z|): the designed set is
    the 10 images of a 160-image pool whose oracle scores are >= 10x the allowed noise apart (asserted), scored with flow
    weights whose subnet gain keeps the scores out of saturation."""
    from sklearn.metrics import roc_auc_score

    from helpers import DESIGNED_NF, NF_TEST_GAIN, assert_designed_separation, assert_map_parity, assert_rel
    from oracle import vitad_oracle as O
    from oracle import weights as W
    from vitad.encoders import EncoderDeit
    from vitad.synthetic import batches, make_designed_set
    from vitad.validators import BLOCK_INDEX_DEIT, ValidatorNF

    images, labels, masks = make_designed_set(DESIGNED_NF)
    enc_sd = W.make_deit_state_dict(seed=11, stress=True)
    nf_sd = W.make_nf_state_dict(seed=31, stress=True, subnet_gain=NF_TEST_GAIN)
    with torch.no_grad():
        tok, _ = O.deit_forward(enc_sd, images, block_index=BLOCK_INDEX_DEIT)
        _, amap, _, _ = O.nf_forward(nf_sd, O.tokens_to_nchw(tok), flow_steps=20, img_size=224)
        ref_scores = O.nf_scores(amap).numpy()
    assert_designed_separation(ref_scores, labels.numpy(), factor=10.0)
    enc = EncoderDeit(224)
    enc.load_state_dict(enc_sd)
    props = {"dataset": "synthetic", "dataclass": "designed", "fp_thres": 0.3}
    val = ValidatorNF([_flow(nf_sd)], enc, None, props)
    res = val.valid_loop_transformer_nf(batches(images, labels, masks, batch_size=4))  # NF scores are per-image: any batching
    assert_rel(res["image_scores"], ref_scores, 1e-3, what="NF image scores")
    assert_map_parity(res["pixel_scores"], amap.numpy(), what="NF anomaly maps")
    assert np.array_equal(np.argsort(ref_scores), np.argsort(res["image_scores"]))
    assert round(roc_auc_score(res["image_labels"], res["image_scores"]), 4) == round(roc_auc_score(labels.numpy(), ref_scores), 4)
