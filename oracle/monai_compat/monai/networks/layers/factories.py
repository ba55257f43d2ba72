"""monai.networks.layers.factories (0.7.0) — name→layer lookups, 3-D subset."""
import torch.nn as nn


class _Factory:
    def __init__(self, table):
        self._t = {k.upper(): v for k, v in table.items()}
        for k in self._t:
            setattr(self, k, k)

    def __getitem__(self, key):
        dim = None
        if isinstance(key, tuple):
            key, dim = key
        fn = self._t[str(key).upper()]
        return fn(dim) if dim is not None and callable(fn) and getattr(fn, "_dimmed", False) else fn


def _dimmed(fn):
    fn._dimmed = True
    return fn


Act = _Factory({
    "relu": nn.ReLU, "leakyrelu": nn.LeakyReLU, "prelu": nn.PReLU, "gelu": nn.GELU,
    "mish": nn.Mish, "sigmoid": nn.Sigmoid, "tanh": nn.Tanh, "elu": nn.ELU,
})
Norm = _Factory({
    "instance": _dimmed(lambda d: (nn.InstanceNorm1d, nn.InstanceNorm2d, nn.InstanceNorm3d)[d - 1]),
    "batch": _dimmed(lambda d: (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)[d - 1]),
    "group": nn.GroupNorm, "layer": nn.LayerNorm,
})
Conv = _Factory({
    "conv": _dimmed(lambda d: (nn.Conv1d, nn.Conv2d, nn.Conv3d)[d - 1]),
    "convtrans": _dimmed(lambda d: (nn.ConvTranspose1d, nn.ConvTranspose2d, nn.ConvTranspose3d)[d - 1]),
})
Conv.CONVTRANS = "CONVTRANS"
