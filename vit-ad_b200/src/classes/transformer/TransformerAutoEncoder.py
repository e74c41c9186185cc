"""src/classes/transformer/TransformerAutoEncoder.py of the reference → vitad.autoencoders."""
from vitad.autoencoders import AutoEncoderDeit, AutoEncoderOutput  # noqa: F401

__all__ = ["AutoEncoderDeit", "AutoEncoderOutput"]
