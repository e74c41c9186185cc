"""Input path: transforms.Resize((224, 224)) of the reference's loader (GeneralDataset.py:38-59) =
Pillow's fixed-point BILINEAR resize.  Integer/byte work: every comparison is bit-exact."""
import numpy as np
import pytest
import torch

from helpers import golden
from oracle.resize_oracle import bilinear_coeffs, resize_bilinear_u8

SIZES = [(900, 900), (1024, 1024), (700, 840), (224, 224), (100, 150), (225, 223)]
GOLDEN_SIZES = [(900, 900), (700, 840), (100, 150)]  # Pillow outputs committed for these (down-, mixed and up-scaling)


def _image(h, w, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    smooth = 127 + 100 * np.sin(yy / 17.0)[..., None] * np.cos(xx[..., None] / 23.0 + np.arange(3))
    return np.clip(smooth + rng.integers(-40, 40, (h, w, 3)), 0, 255).astype(np.uint8)


def test_oracle_matches_pillow_golden():
    """tests/golden/resize_u8.npz holds Pillow 12.2 outputs (oracle/make_golden_resize.py, run in the build container)."""
    g = golden("resize_u8")
    for h, w in GOLDEN_SIZES:
        np.testing.assert_array_equal(resize_bilinear_u8(_image(h, w, SIZES.index((h, w))), 224), g[f"out_{h}x{w}"])


def test_oracle_matches_pillow_when_installed():
    Image = pytest.importorskip("PIL.Image")
    for i, (h, w) in enumerate([(640, 480), (1600, 1200), (37, 53)]):
        img = _image(h, w, 10 + i)
        np.testing.assert_array_equal(resize_bilinear_u8(img, 224), np.asarray(Image.fromarray(img).resize((224, 224), Image.BILINEAR)))


@pytest.mark.parametrize("n_in,n_out", [(900, 224), (1024, 224), (224, 224), (100, 224), (225, 224), (7, 3)])
def test_host_plan_equals_the_oracle_coefficients(n_in, n_out):
    """vitad_resize_plan is host code: runs without a GPU."""
    from vitad import ops

    xmin, cnt, kk = bilinear_coeffs(n_in, n_out)
    plan = ops.resize_plan(n_in, n_out).numpy()
    ks = kk.shape[1]
    assert plan.size == 2 * n_out + n_out * ks
    np.testing.assert_array_equal(plan[:n_out], xmin)
    np.testing.assert_array_equal(plan[n_out:2 * n_out], cnt)
    np.testing.assert_array_equal(plan[2 * n_out:].reshape(n_out, ks), kk)


@pytest.mark.gpu
@pytest.mark.parametrize("h,w", SIZES)
def test_device_resize_is_bit_identical(h, w):
    from vitad import ops

    g = golden("resize_u8")
    i = SIZES.index((h, w))
    imgs = np.stack([_image(h, w, i), _image(h, w, 100 + i)])
    out = ops.resize_u8(torch.from_numpy(imgs).cuda(), 224).cpu().numpy()
    assert out.shape == (2, 3, 224, 224)
    if (h, w) in GOLDEN_SIZES:
        np.testing.assert_array_equal(out[0].transpose(1, 2, 0), g[f"out_{h}x{w}"])
    np.testing.assert_array_equal(out[0].transpose(1, 2, 0), resize_bilinear_u8(imgs[0], 224))
    np.testing.assert_array_equal(out[1].transpose(1, 2, 0), resize_bilinear_u8(imgs[1], 224))


@pytest.mark.gpu
def test_resized_uint8_feeds_the_encoder_like_the_loader_tensor():
    """Resize + ToTensor on the CPU then the fp32 encoder path == device resize then the uint8 encoder path, bit for bit."""
    from oracle import weights as W
    from vitad.encoders import EncoderDeit
    from vitad import ops

    imgs = np.stack([_image(700, 840, 3), _image(700, 840, 4)])
    cpu = np.stack([resize_bilinear_u8(im, 224) for im in imgs]).transpose(0, 3, 1, 2).astype(np.float32) / 255.0
    enc = EncoderDeit(224)
    enc.load_state_dict(W.make_deit_state_dict(seed=11, stress=True))
    enc = enc.cuda().eval()
    with torch.no_grad():
        a = enc(torch.from_numpy(cpu).cuda()).patch_embedding
        b = enc(ops.resize_u8(torch.from_numpy(imgs).cuda(), 224)).patch_embedding
    assert torch.equal(a, b)


@pytest.mark.gpu
def test_validator_accepts_decoded_uint8_images():
    """ValidatorNF.score_batch on raw-size uint8 HWC images == on the loader's Resize + ToTensor tensors, bit for bit."""
    from oracle import weights as W
    from vitad.encoders import EncoderDeit
    from vitad.nf import NormalizingFlow
    from vitad.validators import ValidatorNF

    imgs = np.stack([_image(900, 900, 7), _image(900, 900, 8), _image(900, 900, 9)])
    cpu = np.stack([resize_bilinear_u8(im, 224) for im in imgs]).transpose(0, 3, 1, 2).astype(np.float32) / 255.0
    enc = EncoderDeit(224)
    enc.load_state_dict(W.make_deit_state_dict(seed=11, stress=True))
    np.random.seed(0)
    nf = NormalizingFlow(768, 224, 196, 0.16, 20)
    props = {"dataset": "synthetic", "dataclass": "x", "fp_thres": 0.3}
    val = ValidatorNF([nf], enc, None, props, weights_object=[W.make_nf_state_dict(31)])
    labels = (torch.zeros(3, 1, 224, 224), torch.tensor([0, 1, 0]))
    a = val.valid_loop_transformer_nf([(torch.from_numpy(cpu), *labels)], keep_origs=False)
    b = val.valid_loop_transformer_nf([(torch.from_numpy(imgs), *labels)], keep_origs=False)
    np.testing.assert_array_equal(a["image_scores"], b["image_scores"])
    np.testing.assert_array_equal(a["pixel_scores"], b["pixel_scores"])


def test_resize_plan_rejects_bad_arguments():
    from vitad import _lib

    assert _lib.lib.vitad_resize_ksize(0, 224) == 0 and _lib.lib.vitad_resize_ksize(224, -1) == 0
    assert _lib.lib.vitad_resize_plan(0, 224, None) < 0
    assert _lib.lib.vitad_resize_plan(900, 224, None) < 0  # null plan


@pytest.mark.gpu
def test_resize_rejects_bad_arguments_and_handles_one_pixel_rows():
    from vitad import _lib, ops

    x = torch.zeros(1, 8, 8, 3, dtype=torch.uint8, device="cuda")
    out = torch.empty(1, 3, 4, 4, dtype=torch.uint8, device="cuda")
    plan = ops.resize_plan(8, 4).cuda()
    s = torch.cuda.current_stream().cuda_stream
    f = _lib.lib.vitad_resize_bilinear_u8
    assert f(x.data_ptr(), 1, 8, 8, 4, plan.data_ptr(), plan.data_ptr(), None, out.data_ptr(), s) < 0  # null tmp
    assert f(x.data_ptr(), 0, 8, 8, 4, plan.data_ptr(), plan.data_ptr(), x.data_ptr(), out.data_ptr(), s) < 0  # empty batch
    with pytest.raises(AssertionError):
        ops.resize_u8(torch.zeros(1, 3, 8, 8, dtype=torch.float32, device="cuda"), 4)
    # degenerate geometry: a 1 x 5 image up-scaled to 6 x 6 (every output row reads the single source row)
    img = np.arange(15, dtype=np.uint8).reshape(1, 5, 3) * 15
    got = ops.resize_u8(torch.from_numpy(img[None]).cuda(), 6).cpu().numpy()[0].transpose(1, 2, 0)
    np.testing.assert_array_equal(got, resize_bilinear_u8(img, 6))
