"""Drop-in overlay of the reference's `src` package: put `vit-ad_b200/` in front of the reference root on
sys.path and `from src.classes.MixtureDensityNetwork import …`, `from src.util.ModelHelper
import get_model`, `from src.pipeline.ValidatorMDN import ValidatorMdn` resolve to the CUDA implementations."""
from ._overlay import extend

__path__ = extend(list(__path__))
