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


@pytest.mark.parametrize("stress", [False, True])
@pytest.mark.parametrize("B", [1, 3])
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
