"""src/pipeline/ValidatorMDN.py of the reference → vitad.validators.ValidatorMdn."""
from vitad.validators import ValidatorMdn  # noqa: F401

__all__ = ["ValidatorMdn"]
