"""src/classes/transformer/TransformerEncoder.py of the reference → CUDA implementations (vitad.encoders)."""
from vitad.encoders import EncoderDeit, TransformerEncoder, TransformerEncoderOutput  # noqa: F401

__all__ = ["EncoderDeit", "TransformerEncoder", "TransformerEncoderOutput"]
