"""Generate tests/golden/*.npz by running the REFERENCE's own Python classes (imported unchanged from
/root/reference, with the stand-ins in oracle/shims for the un-installable timm/FrEIA/torchmetrics/
matplotlib/IPython) on seeded synthetic inputs and weights.  TEST INFRASTRUCTURE.

Run in the build container only (the reference tree does not exist on the GPU box):
    python oracle/make_golden.py [--only NAME]
The fixtures are committed; tests regenerate inputs/weights from the same seeds (oracle/weights.py).

Determinism: modules in .eval(); torch.no_grad(); the Gumbel noise of
MixtureDensityNetwork.py:62 is injected by patching torch.nn.functional.gumbel_softmax to
softmax(logits + g) with g from oracle.vitad_oracle.gumbel_noise — no reference file is edited.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("VITAD_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, REF)

from oracle import weights as W  # noqa: E402
from oracle.vitad_oracle import gumbel_noise  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


class GumbelInjector:
    """Context manager: k-th gumbel_softmax call inside uses noise from generator seed (seed + k)."""

    def __init__(self, seed: int):
        self.seed, self.calls = seed, 0

    def __enter__(self):
        self._orig = torch.nn.functional.gumbel_softmax

        def patched(logits, tau=1, hard=False, eps=1e-10, dim=-1):
            g = gumbel_noise(logits.shape, torch.Generator().manual_seed(self.seed + self.calls))
            self.calls += 1
            return torch.softmax((logits + g) / tau, dim=dim)

        torch.nn.functional.gumbel_softmax = patched
        return self

    def __exit__(self, *exc):
        torch.nn.functional.gumbel_softmax = self._orig


def build_ref_deit(sd):
    from src.classes.transformer.TransformerEncoder import EncoderDeit

    enc = EncoderDeit(img_size=224, requires_grad=True)  # => pretrained=False (TransformerEncoder.py:134-136)
    enc.load_state_dict(sd, strict=True)
    for p in enc.parameters():
        p.requires_grad = False
    return enc.eval()


def sub_tokens(t):  # [B,196,768] -> every 14th token
    return t[:, ::14, :].contiguous().numpy()


def save(name, **arrays):
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


# ------------------------------------------------------------------------------------------ cases
def case_deit():
    """EncoderDeit.forward, block_index 0 and 7, default and stress weights, B=2."""
    out = {}
    for tag, stress in (("default", False), ("stress", True)):
        sd = W.make_deit_state_dict(seed=11, stress=stress)
        enc = build_ref_deit(sd)
        x = W.synthetic_images(seed=3, batch=2)
        with torch.no_grad():
            for bi in (0, 7):
                o = enc(x, block_index=bi)
                out[f"{tag}_b{bi}_tokens_sub"] = sub_tokens(o.patch_embedding)
                out[f"{tag}_b{bi}_token_sum"] = o.patch_embedding.sum(-1).numpy()
                out[f"{tag}_b{bi}_token_abs"] = o.patch_embedding.abs().sum(-1).numpy()
                out[f"{tag}_b{bi}_cls"] = o.latent_space.numpy()
    save("deit_b2", **out)


def case_vit():
    """EncoderVit.forward (timm vit_base_patch16_224, one prefix token), default and stress weights, B=2."""
    from src.classes.transformer.TransformerEncoder import EncoderVit

    out = {}
    for tag, stress in (("default", False), ("stress", True)):
        enc = EncoderVit(img_size=224, requires_grad=True)
        enc.load_state_dict(W.make_vit_state_dict(seed=13, stress=stress), strict=True)
        enc.eval()
        x = W.synthetic_images(seed=4, batch=2)
        with torch.no_grad():
            o = enc(x)
        out[f"{tag}_tokens_sub"] = sub_tokens(o.patch_embedding)
        out[f"{tag}_token_sum"] = o.patch_embedding.sum(-1).numpy()
        out[f"{tag}_cls"] = o.latent_space.numpy()
    save("vit_b2", **out)


def case_esvit_interpolate():
    """interpolate_position_encoding of the reference on a window-7 checkpoint loaded into the window-14 model."""
    from src.classes.transformer.SwinTransformerModule import SwinTransformer
    from src.classes.transformer.TransformerEncoder import interpolate_position_encoding

    model = SwinTransformer(patch_size=4, img_size=224, num_classes=3, window_size=14, use_dense_prediction=True)
    delattr(model, "head")
    out = interpolate_position_encoding(weights=W.synthetic_esvit_checkpoint(), model=model)
    keep = {k.replace(".", "__"): v.float().numpy() for k, v in out.items()
            if ("relative_position" in k and (k.startswith("layers.0.blocks.1") or k.startswith("layers.2.blocks.0")
                                              or k.startswith("layers.3.blocks.1")))}
    save("esvit_interpolate", **keep)


class ListLoader:
    """Stands in for GeneralDataLoader: get_dataloader() returns an iterable of batches."""

    def __init__(self, batches):
        self.batches = batches

    def get_dataloader(self, centering=False):
        return self.batches


def synthetic_labels(batch, seed):
    g = torch.Generator().manual_seed(seed)
    image_labels = (torch.rand(batch, generator=g) > 0.5).long()
    pixel_labels = torch.zeros(batch, 1, 224, 224)
    return pixel_labels, image_labels


def case_gmm_validator():
    """ValidatorMdn.valid_loop_transformer end to end: DeiT (stress weights) + MDN K=100, two batches (2 + 1
    images: the short tail batch the reference DataLoader produces without drop_last)."""
    from src.classes.MixtureDensityNetwork import GaussianMixtureDensityNetwork, get_probability_map, log_likelihood
    from src.pipeline.ValidatorMDN import ValidatorMdn

    enc = build_ref_deit(W.make_deit_state_dict(seed=11, stress=True))
    out = {}
    for tag, stress in (("default", False), ("stress", True)):
        mdn_sd = W.make_mdn_state_dict(seed=21, num_gaussians=100, stress=stress)
        mdn = GaussianMixtureDensityNetwork(768, 768, 100)
        imgs = W.synthetic_images(seed=5, batch=3)
        batches = []
        for s, e in ((0, 2), (2, 3)):
            pl, il = synthetic_labels(e - s, seed=s)
            batches.append((imgs[s:e], pl, il))
        props = {"dataset": "synthetic", "dataclass": "x", "num_gaussians": 100, "fp_thres": 0.3}
        val = ValidatorMdn([mdn], enc, ListLoader(batches), props, weights_object=[mdn_sd])
        with torch.no_grad(), GumbelInjector(seed=700):
            res = val.valid_loop_transformer(batches)
        out[f"{tag}_image_scores"] = res["image_scores"]
        out[f"{tag}_pixel_scores_sub"] = res["pixel_scores"][:, :, ::8, ::8]
        out[f"{tag}_pixel_scores_sum"] = res["pixel_scores"].sum(axis=(1, 2, 3))
        # per-patch mean log-likelihood before the batch-max normalisation (first batch only)
        with torch.no_grad(), GumbelInjector(seed=700):
            f = enc(batches[0][0])
            r = mdn(f.patch_embedding)
            out[f"{tag}_L_batch0"] = log_likelihood(f.patch_embedding, r.pi, r.sigma, r.mu).mean(2).numpy()
        with torch.no_grad(), GumbelInjector(seed=700):
            out[f"{tag}_prob_batch0"] = get_probability_map(f.patch_embedding, r.pi, r.sigma, r.mu).numpy()
    save("gmm_validator_k100", **out)


def case_gmm_head_k130():
    """Head only, EsViT-shaped input: B=3, P=49, K=130 on random unit-variance features."""
    from src.classes.MixtureDensityNetwork import GaussianMixtureDensityNetwork, get_probability_map, log_likelihood

    out = {}
    for tag, stress in (("default", False), ("stress", True)):
        sd = W.make_mdn_state_dict(seed=22, num_gaussians=130, stress=stress)
        mdn = GaussianMixtureDensityNetwork(768, 768, 130)
        mdn.load_state_dict(sd, strict=True)
        mdn.eval()
        x = torch.randn(3, 49, 768, generator=torch.Generator().manual_seed(9))
        with torch.no_grad(), GumbelInjector(seed=800):
            r = mdn(x)
            out[f"{tag}_L"] = log_likelihood(x, r.pi, r.sigma, r.mu).mean(2).numpy()
        with torch.no_grad(), GumbelInjector(seed=800):
            out[f"{tag}_prob"] = get_probability_map(x, r.pi, r.sigma, r.mu).numpy()
    save("gmm_head_k130_p49", **out)


def case_gmm_head_k150():
    """Head only at the reference's default mixture count (startTraining_mdn.py:37, README `-n 150`) and at the largest
    count its result tables report (170, csv_results_gmm): B=2, P=49, random unit-variance features."""
    from src.classes.MixtureDensityNetwork import GaussianMixtureDensityNetwork, get_probability_map, log_likelihood

    out = {}
    for K in (150, 170):
        sd = W.make_mdn_state_dict(seed=20 + K, num_gaussians=K, stress=True)
        mdn = GaussianMixtureDensityNetwork(768, 768, K)
        mdn.load_state_dict(sd, strict=True)
        mdn.eval()
        x = torch.randn(2, 49, 768, generator=torch.Generator().manual_seed(K))
        with torch.no_grad(), GumbelInjector(seed=900 + K):
            r = mdn(x)
            out[f"k{K}_L"] = log_likelihood(x, r.pi, r.sigma, r.mu).mean(2).numpy()
        with torch.no_grad(), GumbelInjector(seed=900 + K):
            out[f"k{K}_prob"] = get_probability_map(x, r.pi, r.sigma, r.mu).numpy()
    save("gmm_head_k150_k170_p49", **out)


def case_nf_validator():
    """ValidatorNF.valid_loop_transformer_nf: DeiT (stress) + NormalizingFlow(768,224,196,0.16,20), B=2."""
    from src.classes.NormalizingFlow import NormalizingFlow
    from src.pipeline.ValidatorNF import ValidatorNF

    enc = build_ref_deit(W.make_deit_state_dict(seed=11, stress=True))
    out = {}
    for tag, stress in (("default", False), ("stress", True)):
        np.random.seed(0)
        nf = NormalizingFlow(num_channels=768, img_size=224, num_patches=196, hidden_ratio=0.16, flow_steps=20)
        nf_sd = W.make_nf_state_dict(seed=31, stress=stress)
        imgs = W.synthetic_images(seed=6, batch=2)
        pl, il = synthetic_labels(2, seed=0)
        batches = [(imgs, pl, il)]
        props = {"dataset": "synthetic", "dataclass": "x", "fp_thres": 0.3}
        val = ValidatorNF([nf], enc, ListLoader(batches), props, weights_object=[nf_sd])
        with torch.no_grad():
            res = val.valid_loop_transformer_nf(batches)
            emb = enc(imgs, block_index=0).patch_embedding
            r = nf(emb.transpose(2, 1).reshape(-1, 768, 14, 14))
        out[f"{tag}_image_scores"] = res["image_scores"]
        out[f"{tag}_pixel_scores_sub"] = res["pixel_scores"][:, :, ::8, ::8]
        out[f"{tag}_pixel_scores_sum"] = res["pixel_scores"].sum(axis=(1, 2, 3))
        out[f"{tag}_loss"] = r.loss.numpy()
    save("nf_validator", **out)


def case_recon_l2():
    """VanillaAutoEncoder.MSELoss(reduction='none') + ValidatorRecon tail on random recon/images, B=3."""
    from torch import nn

    g = torch.Generator().manual_seed(4)
    images = torch.rand(3, 3, 224, 224, generator=g)
    recon = torch.tanh(torch.randn(3, 3, 224, 224, generator=g))
    mse = nn.MSELoss(reduction="none")(recon, images)  # CnnAutoEncoder.py:49,68-74
    amap = torch.mean(input=mse, dim=1, keepdim=True)  # ValidatorRecon.py:111
    score = torch.amax(amap, (1, 2, 3))  # ValidatorRecon.py:116
    save("recon_l2", map_sub=amap[:, :, ::8, ::8].numpy(), map_sum=amap.sum(dim=(1, 2, 3)).numpy(),
         image_scores=score.numpy())


def case_recon_validator():
    """ValidatorRecon.valid_loop_mse with get_model('ae_deit_small') (DeiT stress weights + small CNN decoder), B=2."""
    from src.pipeline.ValidatorRecon import ValidatorRecon
    from src.util.ModelHelper import get_model

    model = get_model("ae_deit_small", 224, requires_grad=True)
    sd = {("encoder." + k): v for k, v in W.make_deit_state_dict(seed=11, stress=True).items()}
    sd.update(W.make_small_decoder_state_dict(seed=41))
    imgs = W.synthetic_images(seed=8, batch=2)
    pl, il = synthetic_labels(2, seed=0)
    batches = [(imgs, pl, il)]
    props = {"dataset": "synthetic", "dataclass": "x", "fp_thres": 0.3}
    val = ValidatorRecon(model, ListLoader(batches), props, weights_object=sd)
    with torch.no_grad():
        res = val.valid_loop_mse(batches)
    save("recon_validator", image_scores=res["image_scores"], pixel_scores_sub=res["pixel_scores"][:, :, ::8, ::8],
         pixel_scores_sum=res["pixel_scores"].sum(axis=(1, 2, 3)), recons_sub=res["recons"][:, :, ::8, ::8])


def case_resnet_decoder():
    """DecoderResNetVariableEmbeddingSize(768) (CnnDecoder.py:158-196, imported unchanged) in eval mode on seeded latents."""
    from src.classes.CnnDecoder import DecoderResNetVariableEmbeddingSize

    dec = DecoderResNetVariableEmbeddingSize(embedding_size=768)
    sd = W.make_resnet_decoder_state_dict(seed=43)
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items()})
    dec.eval()
    z = torch.randn(2, 768, generator=torch.Generator().manual_seed(1)) * 0.7
    with torch.no_grad():
        recon = dec(z)
    save("resnet_decoder", recon_sub=recon[:, :, ::4, ::4].numpy(), recon_sum=recon.sum(dim=(2, 3)).numpy(),
         state_dict_keys=np.array(sorted(dec.state_dict().keys())))


def case_recon_validator_resnet():
    """ValidatorRecon.valid_loop_mse with get_model('ae_deit') (DeiT stress weights + reverse-ResNet decoder), B=2."""
    from src.pipeline.ValidatorRecon import ValidatorRecon
    from src.util.ModelHelper import get_model

    model = get_model("ae_deit", 224, requires_grad=True)
    sd = {("encoder." + k): v for k, v in W.make_deit_state_dict(seed=11, stress=True).items()}
    sd.update(W.make_resnet_decoder_state_dict(seed=43))
    imgs = W.synthetic_images(seed=8, batch=2)
    pl, il = synthetic_labels(2, seed=0)
    batches = [(imgs, pl, il)]
    props = {"dataset": "synthetic", "dataclass": "x", "fp_thres": 0.3}
    val = ValidatorRecon(model, ListLoader(batches), props, weights_object=sd)
    with torch.no_grad():
        res = val.valid_loop_mse(batches)
    save("recon_validator_resnet", image_scores=res["image_scores"], pixel_scores_sub=res["pixel_scores"][:, :, ::8, ::8],
         pixel_scores_sum=res["pixel_scores"].sum(axis=(1, 2, 3)), recons_sub=res["recons"][:, :, ::8, ::8],
         state_dict_keys=np.array(sorted(model.state_dict().keys())))


CASES = {
    "deit": case_deit,
    "vit": case_vit,
    "esvit_interpolate": case_esvit_interpolate,
    "gmm_validator": case_gmm_validator,
    "gmm_head_k130": case_gmm_head_k130,
    "gmm_head_k150": case_gmm_head_k150,
    "nf_validator": case_nf_validator,
    "recon_l2": case_recon_l2,
    "recon_validator": case_recon_validator,
    "resnet_decoder": case_resnet_decoder,
    "recon_validator_resnet": case_recon_validator_resnet,
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    for name, fn in CASES.items():
        if a.only in (None, name):
            print("== case", name)
            fn()
