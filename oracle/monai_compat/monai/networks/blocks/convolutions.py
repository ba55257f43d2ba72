"""monai.networks.blocks.convolutions.Convolution (0.7.0): Sequential(conv[, adn])."""
import numpy as np
import torch.nn as nn

from ..layers.convutils import same_padding
from ..layers.factories import Conv
from .adn import ADN


class Convolution(nn.Sequential):
    def __init__(self, dimensions, in_channels, out_channels, strides=1, kernel_size=3,
                 adn_ordering="NDA", act="PRELU", norm="INSTANCE", dropout=None, dropout_dim=1,
                 dilation=1, groups=1, bias=True, conv_only=False, is_transposed=False,
                 padding=None, output_padding=None):
        super().__init__()
        self.dimensions = dimensions
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.is_transposed = is_transposed
        if padding is None:
            padding = same_padding(kernel_size, dilation)
        ctor = Conv[Conv.CONVTRANS if is_transposed else Conv.CONV, dimensions]
        if is_transposed:
            if output_padding is None:
                output_padding = np.atleast_1d(strides) - 1
                output_padding = tuple(int(v) for v in output_padding)
            conv = ctor(in_channels, out_channels, kernel_size=kernel_size, stride=strides,
                        padding=padding, output_padding=output_padding, groups=groups, bias=bias,
                        dilation=dilation)
        else:
            conv = ctor(in_channels, out_channels, kernel_size=kernel_size, stride=strides,
                        padding=padding, dilation=dilation, groups=groups, bias=bias)
        self.add_module("conv", conv)
        if not conv_only:
            self.add_module("adn", ADN(ordering=adn_ordering, in_channels=out_channels, act=act,
                                       norm=norm, norm_dim=dimensions, dropout=dropout,
                                       dropout_dim=dropout_dim))
