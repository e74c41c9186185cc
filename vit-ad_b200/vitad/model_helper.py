"""Model factory with the reference's names (src/util/ModelHelper.py:8-65).  Only the encoders / auto-encoders
of the scoring path are registered; the other reference names (CNN, EfficientNet, NesT, EfficientFormer
baselines) are outside this implementation and are reported as unknown, as the reference does for typos."""
from __future__ import annotations

import importlib

from . import encoders

# https://pytorch.org/hub/pytorch_vision_resnet/ (ModelHelper.py:4-6; imported by validation_loop.py:13)
RES_NET_MEAN = [0.485, 0.456, 0.406]
RES_NET_STD = [0.229, 0.224, 0.225]

# Reference names served by the reference's own PyTorch classes when its tree is reachable through the module overlay
# (vit-ad_b200/src with VITAD_REFERENCE_ROOT): the CNN baselines are outside this implementation.
_REFERENCE_FALLTHROUGH = {
    "enc_cnn": ("src.classes.CnnEncoder", "EncoderVanillaCNN"),
    "enc_eff_net": ("src.classes.CnnEncoder", "EfficientNetEncoder"),
    "enc_res_net": ("src.classes.CnnEncoder", "ResNetEncoder"),
    "ae_cnn": ("src.classes.CnnAutoEncoder", "VanillaAutoEncoder"),
    "ae_res_net": ("src.classes.CnnAutoEncoder", "AutoEncoderResNet"),
    "ae_res_net_small": ("src.classes.CnnAutoEncoder", "AutoEncoderResNetSmallDecoder"),
}

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
    if name in _REFERENCE_FALLTHROUGH and name not in MODEL_DICT:
        mod, attr = _REFERENCE_FALLTHROUGH[name]
        try:
            cls = getattr(importlib.import_module(mod), attr)
        except ImportError as e:
            raise NotImplementedError(
                f"model '{name}' is one of the reference's CNN baselines, which the B200 scoring path does not implement; it is "
                f"only available by fall-through to the reference tree (set VITAD_REFERENCE_ROOT, INTEGRATION.md): {e}") from e
        # ModelHelper.py:42-47
        return cls(img_size=img_size, red_mse="none") if "ae" in name else cls(img_size=img_size)
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
