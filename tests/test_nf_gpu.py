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
    assert np.abs(r.anomaly_score_map.cpu().numpy() - ref).max() <= 1e-3 * np.abs(ref).max()
    assert np.abs(r2.anomaly_score_map.cpu().numpy() - ref).max() <= 1e-3 * np.abs(ref).max()
    assert np.abs(r.image_max.cpu().numpy() - O.nf_scores(amap).numpy()).max() <= 1e-3 * np.abs(ref).max()
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
    assert np.abs(res["image_scores"] - ref_s).max() <= 1e-3 * np.abs(ref_s).max(), (res["image_scores"], ref_s)
    assert np.abs(res["pixel_scores"][:, :, ::8, ::8] - ref_m).max() <= 1e-3 * np.abs(ref_m).max()
    np.testing.assert_allclose(res["pixel_scores"].sum(axis=(1, 2, 3)), g[f"{tag}_pixel_scores_sum"], rtol=2e-3)


def test_nf_image_auroc_identical_to_4_decimals_on_synthetic_anomaly_set():
    """north_star's AUROC criterion for the NF head (config 2): ValidatorNF over 24 synthetic MVTec-shaped images, about half of
    them with pasted anomalies, against the oracle (DeiT block 8 features -> 20-step flow -> amax of the bilinear map)."""
    from sklearn.metrics import roc_auc_score

    from oracle import vitad_oracle as O
    from oracle import weights as W
    from vitad.encoders import EncoderDeit
    from vitad.synthetic import batches, make_category
    from vitad.validators import BLOCK_INDEX_DEIT, ValidatorNF

    n = 24
    images, labels, masks = make_category("cable", n, seed=78)
    enc_sd = W.make_deit_state_dict(seed=11, stress=True)
    nf_sd = W.make_nf_state_dict(seed=31, stress=True)
    with torch.no_grad():
        tok, _ = O.deit_forward(enc_sd, images, block_index=BLOCK_INDEX_DEIT)
        _, amap, _, _ = O.nf_forward(nf_sd, O.tokens_to_nchw(tok), flow_steps=20, img_size=224)
        ref_scores = O.nf_scores(amap).numpy()
    enc = EncoderDeit(224)
    enc.load_state_dict(enc_sd)
    props = {"dataset": "synthetic", "dataclass": "cable", "fp_thres": 0.3}
    val = ValidatorNF([_flow(nf_sd)], enc, None, props)
    res = val.valid_loop_transformer_nf(batches(images, labels, masks, batch_size=8))  # NF scores are per-image: any batching
    noise = 1e-3 * np.abs(ref_scores).max()
    assert np.abs(res["image_scores"] - ref_scores).max() <= noise
    assert np.abs(res["pixel_scores"] - amap.numpy()).max() <= 1e-3 * np.abs(amap.numpy()).max()
    # images whose oracle scores are separated by >= 4x the allowed noise (a near-tie could swap without any kernel error)
    keep, last = [], -np.inf
    for i in np.argsort(ref_scores):
        if ref_scores[i] - last >= 4 * noise:
            keep.append(i)
            last = ref_scores[i]
    keep = np.asarray(sorted(keep))
    lab = labels.numpy()[keep]
    assert len(keep) >= 8 and 2 <= lab.sum() <= len(keep) - 2, (len(keep), lab.sum())
    assert round(roc_auc_score(res["image_labels"][keep], res["image_scores"][keep]), 4) == round(
        roc_auc_score(lab, ref_scores[keep]), 4)
