"""src/util/ModelHelper.py of the reference (model factory, :8-65) for the names on the scoring path."""
from vitad.model_helper import MODEL_DICT, RES_NET_MEAN, RES_NET_STD, get_model, get_possible_models  # noqa: F401

__all__ = ["MODEL_DICT", "RES_NET_MEAN", "RES_NET_STD", "get_model", "get_possible_models"]
