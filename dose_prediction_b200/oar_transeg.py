"""Drop-in for `OARSegmentation/Models/Networks/oar_transeg.py` (`Model` :14) and, as `TRANSEG`, for
`OARSegmentation/OldModels/Networks/oar_transeg.py` (the class train_light_transeg.py:20,110 and LinkedNet
train_light_linked_model.py:89 instantiate)."""
from .networks import OARTranseg as Model  # noqa: F401
from .networks import TRANSEG  # noqa: F401

__all__ = ["Model", "TRANSEG"]
