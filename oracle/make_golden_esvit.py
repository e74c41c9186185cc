"""Golden fixture for the EsViT Swin-T encoder: runs the reference's vendored SwinTransformer through
EncoderEsVit (imported unchanged, timm.models.layers stand-in) in .eval() mode.  TEST INFRASTRUCTURE.
Separate from make_golden.py only to keep that script's cases independent."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, os.environ.get("VITAD_REFERENCE", "/root/reference"))

from oracle import weights as W  # noqa: E402
from oracle.make_golden import save  # noqa: E402


def main():
    from src.classes.transformer.TransformerEncoder import EncoderEsVit

    out = {}
    for tag, stress in (("default", False), ("stress", True)):
        enc = EncoderEsVit(img_size=224, requires_grad=True)  # skips the checkpoint load (TransformerEncoder.py:242)
        enc.load_state_dict(W.make_esvit_state_dict(seed=51, stress=stress), strict=True)
        enc.eval()  # deliberate: the reference leaves DropPath(0.1) active in MDN/NF validation
        x = W.synthetic_images(seed=9, batch=2)
        with torch.no_grad():
            o = enc(x)
        out[f"{tag}_tokens"] = o.patch_embedding.numpy()[:, ::3]  # every 3rd of the 49 region tokens
        out[f"{tag}_token_sum"] = o.patch_embedding.sum(-1).numpy()
        out[f"{tag}_latent"] = o.latent_space.numpy()
    save("esvit_b2", **out)


if __name__ == "__main__":
    main()
