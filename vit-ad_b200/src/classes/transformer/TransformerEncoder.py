"""src/classes/transformer/TransformerEncoder.py of the reference → CUDA implementations (vitad.encoders)."""
from vitad.encoders import (  # noqa: F401
    EncoderDeit,
    EncoderEsVit,
    EncoderVit,
    TransformerEncoder,
    TransformerEncoderOutput,
)

__all__ = ["EncoderDeit", "EncoderEsVit", "EncoderVit", "TransformerEncoder", "TransformerEncoderOutput"]
