"""Drop-in for the hot-path classes of `DosePrediction/Models/Networks/c3d.py` (SingleConv :11, UpConv :25, Encoder :41,
Decoder :75, BaseUNet :118)."""
from .networks import BaseUNet, Decoder, Encoder, SingleConv, UpConv  # noqa: F401

__all__ = ["SingleConv", "UpConv", "Encoder", "Decoder", "BaseUNet"]
