"""Model factory with the reference's names (src/util/ModelHelper.py:8-65).  Only the encoders / auto-encoders
of the scoring path are registered; the other reference names (CNN, EfficientNet, NesT, EfficientFormer
baselines) are outside this implementation and are reported as unknown, as the reference does for typos."""
from __future__ import annotations

from . import encoders

MODEL_DICT = {
    "enc_deit": encoders.EncoderDeit,
    "enc_vit": encoders.EncoderVit,
}


def _register_optional():
    try:
        from . import autoencoders

        MODEL_DICT["ae_deit"] = autoencoders.AutoEncoderDeit
        MODEL_DICT["ae_deit_small"] = autoencoders.AutoEncoderDeit
    except ImportError:
        pass
    if hasattr(encoders, "EncoderEsVit"):
        MODEL_DICT["enc_esvit"] = encoders.EncoderEsVit


_register_optional()


def get_model(name: str, img_size: int = 224, requires_grad: bool = False):
    """ModelHelper.py:33-65: `ae*` models get red_mse='none' (and decoder='cnn' for `*_small`)."""
    try:
        cls = MODEL_DICT[name]
        if "ae" in name:
            if "small" in name:
                return cls(img_size=img_size, requires_grad=requires_grad, red_mse="none", decoder="cnn")
            return cls(img_size=img_size, requires_grad=requires_grad, red_mse="none")
        return cls(img_size=img_size, requires_grad=requires_grad)
    except KeyError:
        print(f"Defined model ${name} not known. Please specify one of the following model names: \n {get_possible_models()}")
        return None


def get_possible_models():
    return list(MODEL_DICT.keys())
