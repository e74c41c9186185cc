from . import layers  # noqa: F401
from .vision_transformer import (  # noqa: F401
    VisionTransformerDistilled,
    deit_base_distilled_patch16_224,
    vit_base_patch16_224,
)


def _unavailable(name):
    def ctor(*a, **k):
        raise NotImplementedError(f"timm stand-in: {name} is outside the scoring path (SURVEY.md §2 #11)")

    return ctor


jx_nest_tiny = _unavailable("jx_nest_tiny")
efficientformer_l3 = _unavailable("efficientformer_l3")
