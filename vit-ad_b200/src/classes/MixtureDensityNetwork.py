"""src/classes/MixtureDensityNetwork.py of the reference → CUDA implementations (vitad.mdn)."""
from vitad.mdn import (  # noqa: F401
    GaussianMixtureDensityNetwork,
    MdnReturn,
    get_probability_map,
    log_gaussian_density,
    log_likelihood,
    mdn_loss,
)

__all__ = ["GaussianMixtureDensityNetwork", "MdnReturn", "get_probability_map", "log_gaussian_density", "log_likelihood",
           "mdn_loss"]
