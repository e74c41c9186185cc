"""src/pipeline/ValidatorNF.py of the reference → vitad.validators.ValidatorNF."""
from vitad.validators import BLOCK_INDEX_DEIT, ValidatorNF  # noqa: F401

__all__ = ["ValidatorNF", "BLOCK_INDEX_DEIT"]
