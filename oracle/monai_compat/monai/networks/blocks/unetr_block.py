"""monai.networks.blocks.unetr_block (0.7.0)."""
import torch
import torch.nn as nn

from .dynunet_block import UnetBasicBlock, UnetResBlock, get_conv_layer


class UnetrUpBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, upsample_kernel_size,
                 norm_name, res_block=False):
        super().__init__()
        self.transp_conv = get_conv_layer(spatial_dims, in_channels, out_channels,
                                          kernel_size=upsample_kernel_size, stride=upsample_kernel_size,
                                          conv_only=True, is_transposed=True)
        blk = UnetResBlock if res_block else UnetBasicBlock
        self.conv_block = blk(spatial_dims, out_channels + out_channels, out_channels,
                              kernel_size=kernel_size, stride=1, norm_name=norm_name)

    def forward(self, inp, skip):
        out = self.transp_conv(inp)
        out = torch.cat((out, skip), dim=1)
        return self.conv_block(out)


class UnetrPrUpBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, num_layer, kernel_size, stride,
                 upsample_kernel_size, norm_name, conv_block=False, res_block=False):
        super().__init__()
        us = upsample_kernel_size
        self.transp_conv_init = get_conv_layer(spatial_dims, in_channels, out_channels, kernel_size=us,
                                               stride=us, conv_only=True, is_transposed=True)
        if conv_block:
            blk = UnetResBlock if res_block else UnetBasicBlock
            self.blocks = nn.ModuleList([
                nn.Sequential(
                    get_conv_layer(spatial_dims, out_channels, out_channels, kernel_size=us, stride=us,
                                   conv_only=True, is_transposed=True),
                    blk(spatial_dims=spatial_dims, in_channels=out_channels, out_channels=out_channels,
                        kernel_size=kernel_size, stride=stride, norm_name=norm_name),
                ) for _ in range(num_layer)])
        else:
            self.blocks = nn.ModuleList([
                get_conv_layer(spatial_dims, out_channels, out_channels, kernel_size=us, stride=us,
                               conv_only=True, is_transposed=True) for _ in range(num_layer)])

    def forward(self, x):
        x = self.transp_conv_init(x)
        for blk in self.blocks:
            x = blk(x)
        return x


class UnetrBasicBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, stride, norm_name,
                 res_block=False):
        super().__init__()
        blk = UnetResBlock if res_block else UnetBasicBlock
        self.layer = blk(spatial_dims=spatial_dims, in_channels=in_channels, out_channels=out_channels,
                         kernel_size=kernel_size, stride=stride, norm_name=norm_name)

    def forward(self, inp):
        return self.layer(inp)
