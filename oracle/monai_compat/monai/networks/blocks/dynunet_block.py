"""monai.networks.blocks.dynunet_block (0.7.0) — the pieces UNETR-style nets use."""
import numpy as np
import torch.nn as nn

from ..layers.factories import Act, Norm
from .convolutions import Convolution


def get_padding(kernel_size, stride):
    k = np.atleast_1d(kernel_size)
    s = np.atleast_1d(stride)
    p = (k - s + 1) / 2
    if np.min(p) < 0:
        raise AssertionError("padding value should not be negative, please change the kernel size and/or stride.")
    p = tuple(int(v) for v in p)
    return p if len(p) > 1 else p[0]


def get_output_padding(kernel_size, stride, padding):
    k = np.atleast_1d(kernel_size)
    s = np.atleast_1d(stride)
    p = np.atleast_1d(padding)
    o = 2 * p + s - k
    if np.min(o) < 0:
        raise AssertionError("out_padding value should not be negative, please change the kernel size and/or stride.")
    o = tuple(int(v) for v in o)
    return o if len(o) > 1 else o[0]


def get_conv_layer(spatial_dims, in_channels, out_channels, kernel_size=3, stride=1, act=Act.PRELU,
                   norm=Norm.INSTANCE, dropout=None, bias=False, conv_only=True, is_transposed=False):
    padding = get_padding(kernel_size, stride)
    output_padding = None
    if is_transposed:
        output_padding = get_output_padding(kernel_size, stride, padding)
    return Convolution(spatial_dims, in_channels, out_channels, strides=stride, kernel_size=kernel_size,
                       act=act, norm=norm, dropout=dropout, bias=bias, conv_only=conv_only,
                       is_transposed=is_transposed, padding=padding, output_padding=output_padding)


def get_norm_layer(name, spatial_dims, channels):
    kw = {}
    if isinstance(name, (tuple, list)):
        name, kw = name[0], dict(name[1])
    return Norm[name, spatial_dims](channels, **kw)


class UnetResBlock(nn.Module):
    """conv1→norm1→lrelu→conv2→norm2 (+ conv3→norm3 on the residual when shape changes) →lrelu.
    0.7.0 constructs conv3/norm3 unconditionally (they sit unused in state_dict when in==out)."""

    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, stride, norm_name, dropout=None):
        super().__init__()
        self.conv1 = get_conv_layer(spatial_dims, in_channels, out_channels, kernel_size=kernel_size,
                                    stride=stride, dropout=dropout, conv_only=True)
        self.conv2 = get_conv_layer(spatial_dims, out_channels, out_channels, kernel_size=kernel_size,
                                    stride=1, dropout=dropout, conv_only=True)
        self.conv3 = get_conv_layer(spatial_dims, in_channels, out_channels, kernel_size=1,
                                    stride=stride, dropout=dropout, conv_only=True)
        self.lrelu = nn.LeakyReLU(inplace=True, negative_slope=0.01)
        self.norm1 = get_norm_layer(norm_name, spatial_dims, out_channels)
        self.norm2 = get_norm_layer(norm_name, spatial_dims, out_channels)
        self.norm3 = get_norm_layer(norm_name, spatial_dims, out_channels)
        self.downsample = in_channels != out_channels
        if not np.all(np.atleast_1d(stride) == 1):
            self.downsample = True

    def forward(self, inp):
        residual = inp
        out = self.lrelu(self.norm1(self.conv1(inp)))
        out = self.norm2(self.conv2(out))
        if self.downsample:
            residual = self.norm3(self.conv3(residual))
        out += residual
        return self.lrelu(out)


class UnetBasicBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, stride, norm_name, dropout=None):
        super().__init__()
        self.conv1 = get_conv_layer(spatial_dims, in_channels, out_channels, kernel_size=kernel_size,
                                    stride=stride, dropout=dropout, conv_only=True)
        self.conv2 = get_conv_layer(spatial_dims, out_channels, out_channels, kernel_size=kernel_size,
                                    stride=1, dropout=dropout, conv_only=True)
        self.lrelu = nn.LeakyReLU(inplace=True, negative_slope=0.01)
        self.norm1 = get_norm_layer(norm_name, spatial_dims, out_channels)
        self.norm2 = get_norm_layer(norm_name, spatial_dims, out_channels)

    def forward(self, inp):
        out = self.lrelu(self.norm1(self.conv1(inp)))
        return self.lrelu(self.norm2(self.conv2(out)))


class UnetOutBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, dropout=None):
        super().__init__()
        self.conv = get_conv_layer(spatial_dims, in_channels, out_channels, kernel_size=1, stride=1,
                                   dropout=dropout, bias=True, conv_only=True)

    def forward(self, inp):
        return self.conv(inp)
