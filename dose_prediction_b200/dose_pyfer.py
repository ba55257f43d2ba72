"""Drop-in for `DosePrediction/Models/Networks/dose_pyfer.py` (same public names: ViTEncoder :22, PyMSCDecoder :150,
MainSubsetModel :245, Model :325, create_pretrained_unet :363): replace

    from DosePrediction.Models.Networks.dose_pyfer import *     ->     from dose_prediction_b200.dose_pyfer import *
"""
from .networks import MainSubsetModel, Model, PyMSCDecoder, ViTEncoder, create_pretrained_unet  # noqa: F401

__all__ = ["ViTEncoder", "PyMSCDecoder", "MainSubsetModel", "Model", "create_pretrained_unet"]
