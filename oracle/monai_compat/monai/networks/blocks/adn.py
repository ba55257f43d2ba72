"""monai.networks.blocks.ADN (0.7.0): optional Activation / Dropout / Normalisation chain."""
import torch.nn as nn

from ..layers.factories import Act, Norm


def _split(spec):
    if isinstance(spec, (tuple, list)):
        return spec[0], dict(spec[1])
    return spec, {}


class ADN(nn.Sequential):
    def __init__(self, ordering="NDA", in_channels=None, act="RELU", norm=None, norm_dim=None,
                 dropout=None, dropout_dim=1):
        super().__init__()
        ops = {}
        if norm is not None:
            name, kw = _split(norm)
            ops["N"] = Norm[name, norm_dim](in_channels, **kw)
        if act is not None:
            name, kw = _split(act)
            ops["A"] = Act[name](**kw)
        if dropout is not None:
            p = dropout if isinstance(dropout, (int, float)) else _split(dropout)[1].get("p", 0.5)
            ops["D"] = (nn.Dropout, nn.Dropout2d, nn.Dropout3d)[dropout_dim - 1](p)
        for ch in ordering.upper():
            if ch in ops:
                self.add_module(ch, ops[ch])
