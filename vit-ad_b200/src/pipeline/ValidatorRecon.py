"""src/pipeline/ValidatorRecon.py of the reference → vitad.validators.ValidatorRecon."""
from vitad.validators import ValidatorRecon  # noqa: F401

__all__ = ["ValidatorRecon"]
