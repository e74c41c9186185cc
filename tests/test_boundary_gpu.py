"""The reference's evaluation flow through the module overlay on a GPU: `validate_mdn` (validation_loop.py:35-84) and
`validate_nf` (:160-207) restated call for call with `src.*` imports — factory, head constructor, Validator with
`weights_base_path` / `weights_name` pointing at a saved .pth, `calc_all_metrics()` — against a tiny MVTec-layout
dataset on disk.  The reference tree does not exist on the GPU box, so its loader (src/data_loader/GeneralDataset.py:
38-114: PIL open → Resize → ToTensor, masks from ground_truth/*_mask.png) is restated in `DiskLoader` below."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class DiskLoader:
    """GeneralDataLoader(batch_size, base_path, data_path='test', img_size, validation_mode=True).get_dataloader()."""

    def __init__(self, batch_size, base_path, data_path, img_size, validation_mode=True):
        self.batch_size, self.img_size = batch_size, img_size
        root = os.path.join(base_path, data_path)
        self.files = sorted(os.path.join(root, d, f) for d in sorted(os.listdir(root)) for f in sorted(os.listdir(os.path.join(root, d))))

    def _item(self, path):
        from PIL import Image

        def to_tensor(img, size):  # transforms.Resize((s, s)) + ToTensor (GeneralDataset.py:38-59)
            img = img.resize((size, size), Image.BILINEAR)
            a = np.asarray(img, dtype=np.float32) / 255.0
            return torch.from_numpy(a.reshape(size, size, -1)).permute(2, 0, 1).contiguous()

        image = to_tensor(Image.open(path).convert("RGB"), self.img_size)
        if os.path.dirname(path).endswith("good"):
            return image, torch.zeros(1, self.img_size, self.img_size), 0
        mask = Image.open(path.replace("/test/", "/ground_truth/").replace(".png", "_mask.png"))
        pt = to_tensor(mask, self.img_size)
        pt[pt != 0] = 1  # GeneralDataset.py:108-110
        return image, pt, 1

    def get_dataloader(self, amount_data=0, centering=False, only_labels=False):
        items = [self._item(p) for p in self.files]
        out = []
        for s in range(0, len(items), self.batch_size):
            chunk = items[s:s + self.batch_size]
            out.append((torch.stack([c[0] for c in chunk]), torch.stack([c[1] for c in chunk]),
                        torch.tensor([c[2] for c in chunk])))
        return out


def _write_dataset(root, n_good=7, n_bad=6, size=256):
    from PIL import Image

    from vitad.synthetic import make_category

    images, labels, masks = make_category("bottle", n_good + n_bad + 8, seed=91, size=size)
    good = [i for i in range(len(labels)) if labels[i] == 0][:n_good]
    bad = [i for i in range(len(labels)) if labels[i] == 1][:n_bad]
    assert len(good) == n_good and len(bad) == n_bad
    for sub in ("test/good", "test/broken_large", "ground_truth/broken_large"):
        os.makedirs(os.path.join(root, "bottle", sub))
    for j, i in enumerate(good):
        Image.fromarray((images[i].permute(1, 2, 0).numpy() * 255).astype(np.uint8)).save(
            os.path.join(root, "bottle", "test/good", f"{j:03d}.png"))
    for j, i in enumerate(bad):
        Image.fromarray((images[i].permute(1, 2, 0).numpy() * 255).astype(np.uint8)).save(
            os.path.join(root, "bottle", "test/broken_large", f"{j:03d}.png"))
        Image.fromarray((masks[i, 0].numpy() * 255).astype(np.uint8)).save(
            os.path.join(root, "bottle", "ground_truth/broken_large", f"{j:03d}_mask.png"))


def test_validate_mdn_flow_through_the_overlay(tmp_path):
    from src.classes.MixtureDensityNetwork import GaussianMixtureDensityNetwork
    from src.pipeline.ValidatorMDN import ValidatorMdn
    from src.util.ModelHelper import get_model

    from oracle import weights as W
    from vitad.metrics import calc_all_metrics

    data_root, weights_root = str(tmp_path / "data"), str(tmp_path / "weights")
    _write_dataset(data_root)
    os.makedirs(weights_root)
    weight = "100_gaussians_bottle.pth"  # validation_loop.py:38-39: "<K>_..._<dataclass>.pth"
    torch.save(W.make_mdn_state_dict(seed=21, num_gaussians=100, stress=True), os.path.join(weights_root, weight))

    # ---- validate_mdn (validation_loop.py:35-84), one weight
    num_gaussians = int(weight.split("_")[0])
    dataclass = weight.split("_")[-1][:-4]
    feature_extractor = get_model(name="enc_deit", img_size=224, requires_grad=False)
    feature_extractor.load_state_dict(W.make_deit_state_dict(seed=11, stress=True))  # no network for timm's download
    dataloader = DiskLoader(batch_size=4, base_path=f"{data_root}/{dataclass}", data_path="test", img_size=224,
                            validation_mode=True)
    gmm_1 = GaussianMixtureDensityNetwork(cluster_centers=None, input_dim=feature_extractor.size_patch_embedding,
                                          output_dim=feature_extractor.size_patch_embedding, num_gaussians=num_gaussians)
    validator = ValidatorMdn(
        gmm_model=[gmm_1], feature_extractor=feature_extractor, dataloader=dataloader, weights_base_path=weights_root,
        weights_name=[weight],
        props={"architecture": f"{type(feature_extractor).__name__}_{type(gmm_1).__name__}",
               "encoder_type": feature_extractor.architecture, "encoder": type(feature_extractor).__name__,
               "num_gaussians": num_gaussians, "dataclass": dataclass, "dataset": "mvtec", "experiment": "t",
               "fp_thres": 0.3})
    torch.manual_seed(5)  # the validator's noise seed follows torch's generator
    metrics = validator.calc_all_metrics()
    for k in ("image_auroc_score", "image_prauc_score", "pixel_auroc_score", "pro_score_0.3fp"):
        assert 0.0 <= metrics[k] <= 1.0, (k, metrics)

    # the result dictionary of ValidatorMDN.py:177-183 and the sklearn metrics of ValidationHelper.py:131-211 on it
    res = validator.valid_loop_transformer(dataloader.get_dataloader())
    assert set(res) >= {"image_scores", "pixel_scores", "image_labels", "pixel_labels", "origs"}
    assert res["image_scores"].shape == (13,) and res["pixel_scores"].shape == (13, 1, 224, 224)
    assert res["origs"].shape == (13, 3, 224, 224) and res["image_labels"].sum() == 6
    ref = calc_all_metrics(res, fp_thres=0.3, dataset_name="mvtec_bottle")
    on_dev = validator.calc_all_metrics()  # same seed -> same noise -> same scores
    for k, v in ref.items():
        if isinstance(v, float):
            assert round(on_dev[k], 4) == round(v, 4), (k, on_dev[k], v)
    # loaded from the .pth, not the constructor's init
    sd = torch.load(os.path.join(weights_root, weight))
    assert torch.equal(gmm_1.mu.weight.detach().cpu(), sd["mu.weight"])


def test_validate_nf_flow_through_the_overlay(tmp_path):
    from src.classes.NormalizingFlow import NormalizingFlow
    from src.pipeline.ValidatorNF import ValidatorNF
    from src.util.ModelHelper import get_model

    from oracle import weights as W

    data_root, weights_root = str(tmp_path / "data"), str(tmp_path / "weights")
    _write_dataset(data_root)
    os.makedirs(weights_root)
    weight = "nf_bottle.pth"
    torch.save(W.make_nf_state_dict(seed=31, stress=True), os.path.join(weights_root, weight))
    # ---- validate_nf (validation_loop.py:160-207)
    feature_extractor = get_model(name="enc_deit", img_size=224, requires_grad=False)
    feature_extractor.load_state_dict(W.make_deit_state_dict(seed=11, stress=True))
    dataloader = DiskLoader(batch_size=32, base_path=f"{data_root}/bottle", data_path="test", img_size=224)
    np.random.seed(0)
    nf = NormalizingFlow(num_channels=feature_extractor.size_patch_embedding, img_size=224,
                         num_patches=feature_extractor.num_embedded_patches, hidden_ratio=0.16, flow_steps=20)
    validator = ValidatorNF(nf_model=[nf], feature_extractor=feature_extractor, dataloader=dataloader,
                            weights_base_path=weights_root, weights_name=[weight],
                            props={"dataclass": "bottle", "dataset": "mvtec", "fp_thres": 0.3})
    metrics = validator.calc_all_metrics()
    for k in ("image_auroc_score", "pixel_auroc_score"):
        assert 0.0 <= metrics[k] <= 1.0, (k, metrics)


def test_cnn_encoder_is_rejected_by_the_validators():
    from vitad.mdn import GaussianMixtureDensityNetwork
    from vitad.validators import ValidatorMdn

    v = ValidatorMdn([GaussianMixtureDensityNetwork(768, 768, 100)], torch.nn.Conv2d(3, 8, 3), None,
                     {"dataset": "d", "dataclass": "c", "num_gaussians": 100})
    with pytest.raises(NotImplementedError, match="CNN feature extractor"):
        v.valid_loop_transformer([])
    with pytest.raises(NotImplementedError, match="valid_loop_resnet"):
        v.valid_loop_resnet([])
