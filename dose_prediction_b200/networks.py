"""Drop-in nn.Module mirrors of the reference networks, executed by hand-written sm_100a kernels.

Same constructor signatures, forward signatures / return structure and state_dict layout (keys, shapes,
order — SURVEY Appendix A, tests/golden/manifest_*.json) as

  DosePrediction/Models/Networks/dose_pyfer.py : ViTEncoder :22, PyMSCDecoder :150, MainSubsetModel :245,
                                                 Model :325, create_pretrained_unet :363
  DosePrediction/Models/Networks/c3d.py        : SingleConv :11, UpConv :25, Encoder :41, Decoder :75, BaseUNet :118
  OARSegmentation/Models/Networks/oar_transeg.py : Model :14
  OARSegmentation/Models/Nets/base_blocks.py   : MultiUnetBasicBlock :12, ModifiedUnetrUpBlock :91, ModifiedUnetOutBlock :144
  OARSegmentation/Models/Nets/blocks_MDUNet.py : conv_block_3 :64, conv_block_7 :98, conv_3_1 :132
  monai==0.7.0 (un-vendored dependency)        : ViT, PatchEmbeddingBlock, TransformerBlock, SABlock, MLPBlock,
                                                 UnetResBlock, UnetrBasicBlock, UnetrPrUpBlock, get_conv_layer

The torch layers below are PARAMETER CONTAINERS only (they give the reference's state_dict and default
initialisation); their own forward() is never called.  Each top-level forward() builds (once per input
shape) an engine.Plan — a static schedule of C-ABI kernel launches — and replays it.  Inference
(eval-mode BatchNorm) only; there is no CPU / eager fallback.
"""
import os
from typing import Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from . import engine
from .engine import Act, Plan, Raw, Tokens, blocks16, ceil_div


def _tuple3(v):
    if isinstance(v, (list, tuple)):
        if len(v) != 3:
            raise ValueError(f"Sequence must have length 3, got {len(v)}.")
        return tuple(int(x) for x in v)
    return (int(v),) * 3


# =========================================================================== parameter containers
class _Conv(nn.Sequential):
    """monai Convolution(conv_only=True): Sequential with a single child named 'conv'."""

    def __init__(self, conv, adn=False):
        super().__init__()
        self.add_module("conv", conv)
        if adn:
            self.add_module("adn", nn.Sequential())


def _conv_layer(cin, cout, k, stride=1, bias=False, transposed=False, adn=False):
    """monai dynunet_block.get_conv_layer for 3-D: padding (k-s+1)//2, output_padding 2p+s-k."""
    pad = (k - stride + 1) // 2
    if transposed:
        return _Conv(nn.ConvTranspose3d(cin, cout, k, stride=stride, padding=pad, output_padding=2 * pad + stride - k,
                                        bias=bias), adn)
    return _Conv(nn.Conv3d(cin, cout, k, stride=stride, padding=pad, bias=bias), adn)


class SingleConv(nn.Module):
    def __init__(self, in_ch, out_ch, kernel_size, stride, padding):
        super().__init__()
        self.single_conv = nn.Sequential(
            nn.Conv3d(in_ch, out_ch, kernel_size=kernel_size, padding=padding, stride=stride, bias=True),
            nn.InstanceNorm3d(out_ch, affine=True), nn.ReLU(inplace=True))


class UpConv(nn.Module):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv3d(in_ch, out_ch, kernel_size=3, padding=1, stride=1, bias=True),
                                  nn.InstanceNorm3d(out_ch, affine=True), nn.ReLU(inplace=True))


class Encoder(nn.Module):
    def __init__(self, in_ch, list_ch):
        super().__init__()
        chans = [in_ch] + list(list_ch[1:6])
        for s in range(1, 6):
            setattr(self, f"encoder_{s}", nn.Sequential(
                SingleConv(chans[s - 1], chans[s], kernel_size=3, stride=1 if s == 1 else 2, padding=1),
                SingleConv(chans[s], chans[s], kernel_size=3, stride=1, padding=1)))


class Decoder(nn.Module):
    def __init__(self, list_ch):
        super().__init__()
        for s in (4, 3, 2, 1):
            setattr(self, f"upconv_{s}", UpConv(list_ch[s + 1], list_ch[s]))
            convs = [SingleConv(2 * list_ch[s], list_ch[s], kernel_size=3, stride=1, padding=1)]
            if s != 1:
                convs.append(SingleConv(list_ch[s], list_ch[s], kernel_size=3, stride=1, padding=1))
            setattr(self, f"decoder_conv_{s}", nn.Sequential(*convs))


class BaseUNet(nn.Module):
    """c3d.py:118 — the cascade's first stage (net_A)."""

    def __init__(self, in_ch, list_ch):
        super().__init__()
        self.in_ch, self.list_ch = in_ch, list(list_ch)
        self.encoder = Encoder(in_ch, list_ch)
        self.decoder = Decoder(list_ch)
        for m in self.modules():           # c3d.py:127-142
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0.0)
            elif isinstance(m, nn.InstanceNorm3d):
                nn.init.constant_(m.weight, 1.0)
                nn.init.constant_(m.bias, 0.0)

    def forward(self, x):
        return _run_single(self, x, _plan_base_unet)


class _MLP(nn.Module):
    def __init__(self, hidden, mlp_dim, p):
        super().__init__()
        self.linear1, self.linear2 = nn.Linear(hidden, mlp_dim), nn.Linear(mlp_dim, hidden)
        self.fn, self.drop1, self.drop2 = nn.GELU(), nn.Dropout(p), nn.Dropout(p)


class _SA(nn.Module):
    def __init__(self, hidden, heads, p):
        super().__init__()
        self.num_heads = heads
        self.out_proj = nn.Linear(hidden, hidden)
        self.qkv = nn.Linear(hidden, hidden * 3, bias=False)
        self.drop_output, self.drop_weights = nn.Dropout(p), nn.Dropout(p)
        self.head_dim = hidden // heads
        self.scale = self.head_dim ** -0.5


class _TransformerBlock(nn.Module):
    def __init__(self, hidden, mlp_dim, heads, p):
        super().__init__()
        self.mlp = _MLP(hidden, mlp_dim, p)
        self.norm1 = nn.LayerNorm(hidden)
        self.attn = _SA(hidden, heads, p)
        self.norm2 = nn.LayerNorm(hidden)


class _PatchEmbedding(nn.Module):
    def __init__(self, in_channels, img_size, patch_size, hidden, pos_embed, p):
        super().__init__()
        if pos_embed not in ("perceptron", "conv"):
            raise ValueError(f"unsupported pos_embed {pos_embed!r}")
        self.pos_embed = pos_embed
        for m, q in zip(img_size, patch_size):
            if m < q:
                raise ValueError("patch_size should be smaller than img_size.")
            if m % q != 0:
                raise ValueError("patch_size should be divisible by img_size for perceptron.")
        self.n_patches = int(np.prod([i // q for i, q in zip(img_size, patch_size)]))
        self.patch_dim = int(in_channels * np.prod(patch_size))
        if pos_embed == "conv":      # monai: Conv3d(kernel = stride = patch) then flatten(2).transpose(-1,-2)
            self.patch_embeddings = nn.Conv3d(in_channels, hidden, kernel_size=patch_size, stride=patch_size)
        else:
            self.patch_embeddings = nn.Sequential(nn.Identity(), nn.Linear(self.patch_dim, hidden))
        self.position_embeddings = nn.Parameter(torch.zeros(1, self.n_patches, hidden))
        self.cls_token = nn.Parameter(torch.zeros(1, 1, hidden))
        self.dropout = nn.Dropout(p)
        nn.init.trunc_normal_(self.position_embeddings, mean=0.0, std=0.02, a=-2.0, b=2.0)
        if pos_embed == "perceptron":
            lin = self.patch_embeddings[1]
            nn.init.trunc_normal_(lin.weight, mean=0.0, std=0.02, a=-2.0, b=2.0)
            nn.init.constant_(lin.bias, 0)


class ViT(nn.Module):
    def __init__(self, in_channels, img_size, patch_size, hidden_size=768, mlp_dim=3072, num_layers=12, num_heads=12,
                 pos_embed="conv", classification=False, dropout_rate=0.0, spatial_dims=3):
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise ValueError("hidden_size should be divisible by num_heads.")
        if classification or spatial_dims != 3:
            raise NotImplementedError("only the 3-D, non-classification ViT used by the reference nets is built")
        self.hidden_size, self.mlp_dim, self.num_heads, self.num_layers = hidden_size, mlp_dim, num_heads, num_layers
        self.patch_embedding = _PatchEmbedding(in_channels, img_size, patch_size, hidden_size, pos_embed, dropout_rate)
        self.blocks = nn.ModuleList(
            [_TransformerBlock(hidden_size, mlp_dim, num_heads, dropout_rate) for _ in range(num_layers)])
        self.norm = nn.LayerNorm(hidden_size)


class UnetResBlock(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1):
        super().__init__()
        self.conv1 = _conv_layer(in_channels, out_channels, kernel_size, stride)
        self.conv2 = _conv_layer(out_channels, out_channels, kernel_size, 1)
        self.conv3 = _conv_layer(in_channels, out_channels, 1, stride)      # always constructed in 0.7.0
        self.lrelu = nn.LeakyReLU(inplace=True, negative_slope=0.01)
        self.norm1, self.norm2, self.norm3 = (nn.InstanceNorm3d(out_channels) for _ in range(3))
        self.downsample = in_channels != out_channels or stride != 1


class UnetBasicBlock(nn.Module):
    """monai dynunet_block.UnetBasicBlock: conv -> IN -> lrelu -> conv -> IN -> lrelu (k3, no bias)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1):
        super().__init__()
        self.conv1 = _conv_layer(in_channels, out_channels, kernel_size, stride)
        self.conv2 = _conv_layer(out_channels, out_channels, kernel_size, 1)
        self.lrelu = nn.LeakyReLU(inplace=True, negative_slope=0.01)
        self.norm1, self.norm2 = nn.InstanceNorm3d(out_channels), nn.InstanceNorm3d(out_channels)


class UnetrUpBlock(nn.Module):
    """monai unetr_block.UnetrUpBlock (res_block=False), the decoder of PyMSCDecoder(mode_multi=False)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.transp_conv = _conv_layer(in_channels, out_channels, 2, 2, transposed=True)
        self.conv_block = UnetBasicBlock(out_channels + out_channels, out_channels)

    def forward(self, inp, skip):
        """monai UnetrUpBlock.forward: transp_conv(inp) -> cat(out, skip) -> conv_block."""
        return _up_block_forward(self, inp, skip)


class UnetrBasicBlock(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.layer = UnetResBlock(in_channels, out_channels)


class UnetrPrUpBlock(nn.Module):
    def __init__(self, in_channels, out_channels, num_layer):
        super().__init__()
        self.transp_conv_init = _conv_layer(in_channels, out_channels, 2, 2, transposed=True)
        self.blocks = nn.ModuleList([
            nn.Sequential(_conv_layer(out_channels, out_channels, 2, 2, transposed=True),
                          UnetResBlock(out_channels, out_channels)) for _ in range(num_layer)])


def _act_layer(act):
    return nn.ReLU(inplace=True) if act == "relu" else nn.Mish(inplace=True)


class conv_block_3(nn.Module):
    def __init__(self, ch_in, ch_out, act="relu"):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv3d(ch_in, ch_out, kernel_size=3, stride=1, padding=1, bias=True), nn.InstanceNorm3d(ch_out), _act_layer(act),
            nn.Conv3d(ch_out, ch_out, kernel_size=3, stride=1, padding=1, bias=True), nn.InstanceNorm3d(ch_out), _act_layer(act))


class conv_block_7(nn.Module):
    def __init__(self, ch_in, ch_out):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv3d(ch_in, ch_out, kernel_size=7, stride=1, padding=3, bias=True), nn.BatchNorm3d(ch_out), nn.ReLU(inplace=True),
            nn.Conv3d(ch_out, ch_out, kernel_size=7, stride=1, padding=3, bias=True), nn.BatchNorm3d(ch_out), nn.ReLU(inplace=True))


class conv_3_1(nn.Module):
    def __init__(self, ch_in, ch_out, act):
        super().__init__()
        self.act = act
        self.conv_3 = nn.Sequential(conv_block_3(ch_in, ch_out), nn.InstanceNorm3d(ch_out), _act_layer(act))
        self.conv_7 = nn.Sequential(conv_block_7(ch_in, ch_out), nn.InstanceNorm3d(ch_out), _act_layer(act))
        self.conv = nn.Sequential(nn.Conv3d(ch_out * 2, ch_out, kernel_size=1, stride=1, padding=0, bias=True),
                                  nn.InstanceNorm3d(ch_out), _act_layer(act))


class _dilated_block(nn.Module):
    """blocks_MDUNet.py:64-78 / :160-191: conv(dil) -> IN -> act -> conv(dil) -> IN -> act."""

    def __init__(self, ch_in, ch_out, dil, act="relu"):
        super().__init__()
        self.dil = dil
        self.conv = nn.Sequential(
            nn.Conv3d(ch_in, ch_out, kernel_size=3, stride=1, padding=dil, dilation=dil, bias=True), nn.InstanceNorm3d(ch_out), _act_layer(act),
            nn.Conv3d(ch_out, ch_out, kernel_size=3, stride=1, padding=dil, dilation=dil, bias=True), nn.InstanceNorm3d(ch_out), _act_layer(act))


class DualDilatedBlock(nn.Module):
    """blocks_MDUNet.py:194-215 (multiS_conv=False): dilation 1/2/3 branches -> cat -> 1^3 conv -> IN -> act."""

    def __init__(self, ch_in, ch_out, act="relu"):
        super().__init__()
        self.act = act
        self.conv_3 = _dilated_block(ch_in, ch_out, 1, act)
        self.conv_5 = _dilated_block(ch_in, ch_out, 2, act)
        self.conv_7 = _dilated_block(ch_in, ch_out, 3, act)
        self.conv = nn.Sequential(nn.Conv3d(ch_out * 3, ch_out, kernel_size=1, stride=1, padding=0, bias=True),
                                  nn.InstanceNorm3d(ch_out), _act_layer(act))


class _conv_block_bn(nn.Module):
    """OldModels/Nets/blocks_MDUNet.py:64-78,98-112: conv -> BN -> ReLU -> conv -> BN -> ReLU (k = 3 or 7)."""

    def __init__(self, ch_in, ch_out, k):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv3d(ch_in, ch_out, kernel_size=k, stride=1, padding=k // 2, bias=True), nn.BatchNorm3d(ch_out), nn.ReLU(inplace=True),
            nn.Conv3d(ch_out, ch_out, kernel_size=k, stride=1, padding=k // 2, bias=True), nn.BatchNorm3d(ch_out), nn.ReLU(inplace=True))


class conv_3_1_old(nn.Module):
    """OldModels/Nets/blocks_MDUNet.py:132-147 — the layout the reference's trained TRANSEG checkpoints use."""

    def __init__(self, ch_in, ch_out):
        super().__init__()
        self.conv_3 = _conv_block_bn(ch_in, ch_out, 3)
        self.conv_7 = _conv_block_bn(ch_in, ch_out, 7)
        self.conv = nn.Conv3d(ch_out * 2, ch_out, kernel_size=1, stride=1, padding=0, bias=True)


class MultiUnetBasicBlock(nn.Module):
    def __init__(self, in_channels, out_channels, multiS_conv=True, act="relu", old=False):
        super().__init__()
        if old:
            self.cov_ = conv_3_1_old(in_channels, out_channels)
        elif multiS_conv:
            self.cov_ = conv_3_1(ch_in=in_channels, ch_out=out_channels, act=act)
        else:
            self.cov_ = DualDilatedBlock(ch_in=in_channels, ch_out=out_channels, act=act)


class ModifiedUnetrUpBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, upsample_kernel_size, act="relu", norm="instance",
                 multiS_conv=True, old=False):
        super().__init__()
        if spatial_dims != 3 or upsample_kernel_size != 2:
            raise NotImplementedError("only the 3-D, 2x up-sampling block used by the reference nets is built")
        self.act = act
        self.transp_conv = _conv_layer(in_channels, out_channels, 2, 2, transposed=True)
        self.conv_block = MultiUnetBasicBlock(out_channels + out_channels, out_channels, act=act, multiS_conv=multiS_conv, old=old)

    def forward(self, inp, skip):
        """base_blocks.py:136-141: transp_conv(inp) -> cat(out, skip) -> conv_block.  inp [B,C_in,d,h,w], skip
        [B,C_out,2d,2h,2w] fp32 CUDA -> [B,C_out,2d,2h,2w]."""
        return _up_block_forward(self, inp, skip)


class ModifiedUnetOutBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, dropout=None):
        super().__init__()
        self.conv = _conv_layer(in_channels, out_channels, 1, 1, bias=True, adn=True)


# =========================================================================== plan builders (the forward)
class Precision:
    """Operand precision recipe (DESIGN.md 'precision'; frozen with oracle/precision_probe.py).
    conv3: tensor-core mode of 3^3 convs; conv7: of 7^3 convs; lo: store activations as hi+lo pairs."""

    def __init__(self, conv3, conv7, lo):
        self.conv3, self.conv7, self.lo = conv3, conv7, lo


PREC_NET_A = Precision("p3", "p3", True)      # net_A dominates the dose error: ~22-bit operands
PREC_NET_B = Precision("p1", "p1", False)     # fp16 operands everywhere
PREC_SEG = Precision("p3", "p1", True)        # seg: fp16 on 7^3 convs and the ViT, ~22-bit elsewhere ...
PREC_SEG_DEEP = Precision("p1", "p1", False)  # ... except the two coarsest levels (<= 1/4 resolution): plain fp16
                                              # (oracle/precision_probe.py: argmax agreement unchanged, 99.95 / 99.91 %)


# conv_3_1's 3^3 branch normalised on load by the 1^3 conv (statistics-only pass + two-stage IN in dp_pointwise_tc) instead of
# materialising relu(IN(raw)).  Correct and 4 bytes / element lighter, but measured NEUTRAL at batch 8 x 128^3 (the read-only
# statistics pass runs at 3.4 TB/s: 0.41 -> 0.32 ms, while the Mish-on-load 1^3 conv gets 0.03-0.07 ms slower): off by default.
FUSE_BRANCH_NORM = os.environ.get("DP_FUSE_BRANCH_NORM", "0") != "0"

_S2D_TAPS = {0: ((1, 1),), 1: ((0, 0), (1, 2))}     # input parity -> ((tap of the 3^3 stride-1 conv, original tap), ...)


def _s2d_weight(w):
    """3^3 stride-2 pad-1 conv over x  ==  sparse 3^3 stride-1 conv over space_to_depth(x) (8*C channels):
    out[o] = sum_k W[k] x[2o+k-1]; input 2o+k-1 has parity (k+1)&1 and half-resolution index o + (k-1)>>1."""
    Co, C = w.shape[0], w.shape[1]
    wp = torch.zeros((Co, 8, C, 3, 3, 3), dtype=torch.float32, device=w.device)
    masks = []
    for cls in range(8):
        pd, ph, pw = (cls >> 2) & 1, (cls >> 1) & 1, cls & 1
        m = 0
        for td, kd in _S2D_TAPS[pd]:
            for th, kh in _S2D_TAPS[ph]:
                for tw, kw in _S2D_TAPS[pw]:
                    wp[:, cls, :, td, th, tw] = w[:, :, kd, kh, kw]
                    m |= 1 << ((td * 3 + th) * 3 + tw)
        masks.append(m)
    return wp.view(Co, 8 * C, 3, 3, 3), masks


class _S2DMask:
    """per-K-chunk tap masks for the sparse space-to-depth conv (chunks are 16 channels, C >= 16 per parity)."""

    def __init__(self, C, class_masks, flops):
        self.C, self.class_masks, self.flops = C, class_masks, flops

    def __call__(self, nch):
        per_term = 8 * self.C // 16
        return [self.class_masks[((i % per_term) * 16) // self.C] for i in range(nch)]


def _in_conv_norm(P, parts, conv, norm, k, mode, out, act="relu", stride=1, N=None, dims=None, s2d_in=None, s2d_out=None):
    """conv(+bias) -> InstanceNorm(affine?) -> act, for c3d SingleConv/UpConv.  Stride-2 convs read the
    space-to-depth copy `s2d_in` of their input and run as tap-masked stride-1 convs on the tensor cores."""
    Co = conv.weight.shape[0]
    scale, shift = P.affine(Co, bias=conv.bias)
    odims = dims
    raw = P.get_raw(N, Co, odims)
    if stride == 1:
        P.conv_tc(parts, conv.weight, k, 1, mode, scale, shift, False, out_raw=raw)
    elif s2d_in is not None:
        w = conv.weight.detach().to(P.device, torch.float32)
        wp, class_masks = _s2d_weight(w)
        vox_out = odims[0] * odims[1] * odims[2]
        fn = _S2DMask(w.shape[1], class_masks, 2.0 * N * vox_out * 27 * w.shape[1] * Co)
        P.conv_tc([s2d_in], wp, 3, 1, mode, scale, shift, False, out_raw=raw, tap_mask_fn=fn)
    else:
        assert len(parts) == 1
        P.conv_direct(parts[0], conv.weight, k, stride, 1, scale, shift, False, out_raw=raw)
    g = P.dev(norm.weight) if getattr(norm, "weight", None) is not None else None
    b = P.dev(norm.bias) if getattr(norm, "bias", None) is not None else None
    P.norm_act(raw, out, gamma=g, beta=b, act=act, s2d=s2d_out)
    P.release(raw)


def _emit_base_unet(P, net, x_act, out_act, prec=PREC_NET_A):
    """c3d.py:144-149: Encoder (:65-72) then Decoder (:99-115); the final decoder tensor lands in out_act."""
    N, dims0 = x_act.N, x_act.dims
    ch = net.list_ch
    lo = prec.lo
    dims = [tuple(max(1, (d + (1 << s) - 1) >> s) for d in dims0) for s in range(5)]   # stage s+1 spatial dims
    cat = {}
    for s in (1, 2, 3, 4):
        cat[s] = P.new_concat(N, [ch[s], ch[s]], dims[s - 1], lo=lo)       # [upconv_s output | encoder_s output]
    h = x_act
    s2d = None                  # space-to-depth copy of the previous stage's output (input of the stride-2 conv)
    for s in range(1, 6):
        enc = getattr(net.encoder, f"encoder_{s}")
        a = P.new_act(N, ch[s], dims[s - 1], lo=lo)
        _in_conv_norm(P, [h], enc[0].single_conv[0], enc[0].single_conv[1], 3, prec.conv3, a,
                      stride=1 if s == 1 else 2, N=N, dims=dims[s - 1], s2d_in=s2d)
        dst = cat[s][1] if s <= 4 else P.new_act(N, ch[s], dims[s - 1], lo=lo)
        s2d = None
        if s <= 4 and lo and ch[s] % 16 == 0 and all(d % 2 == 0 for d in dims[s - 1]):
            s2d = P.new_act(N, 8 * ch[s], dims[s], lo=True)
        _in_conv_norm(P, [a], enc[1].single_conv[0], enc[1].single_conv[1], 3, prec.conv3, dst, N=N, dims=dims[s - 1],
                      s2d_out=s2d)
        h = dst
    for s in (4, 3, 2, 1):
        up = P.new_act(N, ch[s + 1], dims[s - 1], lo=lo)
        P.upsample2x(h, up)                                               # c3d.py:36
        upc = getattr(net.decoder, f"upconv_{s}").conv
        _in_conv_norm(P, [up], upc[0], upc[1], 3, prec.conv3, cat[s][0], N=N, dims=dims[s - 1])
        dec = getattr(net.decoder, f"decoder_conv_{s}")
        last = (s == 1)
        d0 = out_act if last else P.new_act(N, ch[s], dims[s - 1], lo=lo)
        _in_conv_norm(P, cat[s], dec[0].single_conv[0], dec[0].single_conv[1], 3, prec.conv3, d0, N=N, dims=dims[s - 1])
        h = d0
        if not last:
            d1 = P.new_act(N, ch[s], dims[s - 1], lo=lo)
            _in_conv_norm(P, [h], dec[1].single_conv[0], dec[1].single_conv[1], 3, prec.conv3, d1, N=N, dims=dims[s - 1])
            h = d1
    return out_act


def _emit_vit(P, vit, parts, N, S, taps, x_planar=None):
    """monai ViT.forward (perceptron patch embedding) -> (LN(x_L) tokens, {layer index: hidden-state tokens}).
    x_planar: the ONE-channel input as planar fp32 (seg net): the patch matrix is then built straight from it with
    K = 4096 (dp_patchify_planar) instead of K = 8 * 4096 through the c8 copy."""
    hidden, heads, L = vit.hidden_size, vit.num_heads, vit.num_layers
    hd = hidden // heads
    grid = tuple(s // 16 for s in S)
    T = grid[0] * grid[1] * grid[2]
    M = N * T
    first = parts[0]
    planar = (x_planar is not None and len(parts) == 1 and parts[0].C == 1 and not P.training
              and os.environ.get("DP_PATCHIFY_PLANAR", "1") != "0")
    ncb = sum(ceil_div(a.C, 8) if i == len(parts) - 1 else blocks16(a.C) for i, a in enumerate(parts))
    # logical channel -> slot inside the contiguous block range starting at parts[0].cb_off
    slots, base = [], 0
    for a in parts:
        assert a.cb_off == first.cb_off + base // 8 and a.buf is first.buf, "ViT input parts must be adjacent"
        slots += [base + c for c in range(a.C)]
        base += blocks16(a.C) * 8
    K = ncb * 4096 * 8
    Cin = len(slots)
    if vit.patch_embedding.pos_embed == "conv":      # Conv3d weight [hidden, C, 16,16,16] == Linear over (c, p1, p2, p3)
        lin = vit.patch_embedding.patch_embeddings
        w = lin.weight.detach().to(P.device, torch.float32).permute(0, 2, 3, 4, 1).contiguous()
    else:
        lin = vit.patch_embedding.patch_embeddings[1]
        w = lin.weight.detach().to(P.device, torch.float32).view(hidden, 16, 16, 16, Cin)
    if planar:
        K = 4096
        wpe = w.reshape(hidden, K).contiguous().half()           # (p1, p2, p3, c = 1): already the planar flatten order
        del w
        P.keep.append(wpe)
        A = P.zeros((M, K), torch.float16)
        P.add("dp_patchify_planar", x_planar.data_ptr(), N, S[0], S[1], S[2], A.data_ptr())
    else:
        wfull = torch.zeros((hidden, 16, 16, 16, ncb * 8), device=P.device)
        wfull[..., torch.tensor(slots, device=P.device)] = w
        wpe = wfull.view(hidden, 16, 16, 16, ncb, 8).permute(0, 4, 1, 2, 3, 5).reshape(hidden, K).contiguous().half()
        del w, wfull
        P.keep.append(wpe)
        # S = (32k, 128, 128): the GEMM gathers its A tiles from the c8 activation by TMA (dp_gemm_patch_embed); other
        # shapes materialise the patch matrix first (dp_patchify)
        gather = (S[1] == 128 and S[2] == 128 and S[0] % 32 == 0 and not P.training
                  and os.environ.get("DP_PATCH_GATHER", "1") != "0")
        A = None
        if not gather:
            A = P.zeros((M, K), torch.float16)
            P.patchify(first, ncb, A)
    x = P.zeros((M, hidden), torch.float32)
    pos = P.dev(vit.patch_embedding.position_embeddings.reshape(T, hidden))
    tiles = ceil_div(M, 128) * ceil_div(hidden, 128)
    split_k = max(1, min(K // 64, (2 * 148) // tiles))
    total_kb = K // 64
    split_k = ceil_div(total_kb, ceil_div(total_kb, split_k))          # no empty splits
    if not planar and gather:
        bias_d = P.dev(lin.bias)
        P.count_flops("dp_gemm_patch_embed", 2.0 * M * hidden * K)
        if split_k == 1:
            P.add("dp_gemm_patch_embed", first.buf.data_ptr(), first.cb_total, first.cb_off, ncb, N, S[0], S[1], S[2],
                  wpe.data_ptr(), hidden, 1, bias_d.data_ptr(), pos.data_ptr(), T, x.data_ptr(), P.err.data_ptr())
        else:
            ws = P.zeros((split_k, M, hidden), torch.float32)
            P.add("dp_gemm_patch_embed", first.buf.data_ptr(), first.cb_total, first.cb_off, ncb, N, S[0], S[1], S[2],
                  wpe.data_ptr(), hidden, split_k, None, None, 0, ws.data_ptr(), P.err.data_ptr())
            P.add("dp_splitk_reduce", ws.data_ptr(), split_k, M, hidden, bias_d.data_ptr(), pos.data_ptr(), T, x.data_ptr())
    else:
        P.gemm_splitk(A, wpe, M, hidden, K, split_k, x, bias=P.dev(lin.bias), rowvec=pos, row_period=T)
    ln = P.zeros((M, hidden), torch.float16)
    q = P.zeros((N * heads, T, hd), torch.float16)
    k = P.zeros((N * heads, T, hd), torch.float16)
    Tp = ceil_div(T, 8) * 8          # TMA row pitch must be a multiple of 16 B: pad the key axis with zeros
    vt = P.zeros((N * heads, hd, Tp), torch.float16)
    fused = hd in (64, 128) and os.environ.get("DP_FUSED_ATTENTION", "1") != "0"
    if not fused:
        scores = P.zeros((N * heads, T, T), torch.float32)
        probs = P.zeros((N * heads, T, Tp), torch.float16)
    o = P.zeros((M, hidden), torch.float16)
    hmid = P.zeros((M, vit.mlp_dim), torch.float16)
    hs = {}
    for i, blk in enumerate(vit.blocks):
        P.layernorm(x, P.dev(blk.norm1.weight), P.dev(blk.norm1.bias), M, hidden, out_f16=ln)
        P.gemm(ln, P.dev(blk.attn.qkv.weight, torch.float16), M, 3 * hidden, hidden, qkv=(heads, hd, T, q, k, vt, hd ** -0.5))
        if fused:
            P.attention(q, k, vt, N, heads, T, hd, o)
        else:
            P.gemm(q, k, T, T, hd, batch=N * heads, a_batch_rows=T, b_batch_rows=T, c_batch_stride=T * T, ldc=T, out_f32=scores)
            P.softmax(scores, N * heads * T, T, probs)
            P.gemm(probs, vt, T, hd, Tp, batch=N * heads, a_batch_rows=T, b_batch_rows=hd, c_batch_stride=T * hidden,
                   c_batch_period=heads, c_batch_stride2=hd, ldc=hidden, out_f16=o)
        P.gemm(o, P.dev(blk.attn.out_proj.weight, torch.float16), M, hidden, hidden, bias=P.dev(blk.attn.out_proj.bias),
               resid=x, out_f32=x)
        P.layernorm(x, P.dev(blk.norm2.weight), P.dev(blk.norm2.bias), M, hidden, out_f16=ln)
        P.gemm(ln, P.dev(blk.mlp.linear1.weight, torch.float16), M, vit.mlp_dim, hidden, bias=P.dev(blk.mlp.linear1.bias),
               act="gelu", out_f16=hmid)
        tap = None
        if i in taps:
            tap = P.zeros((N, T, hidden), torch.float16)
            hs[i] = Tokens(tap, grid)
        P.gemm(hmid, P.dev(blk.mlp.linear2.weight, torch.float16), M, hidden, vit.mlp_dim, bias=P.dev(blk.mlp.linear2.bias),
               resid=x, out_f32=x, out_f16=tap)
    z = P.zeros((N, T, hidden), torch.float16)
    P.layernorm(x, P.dev(vit.norm.weight), P.dev(vit.norm.bias), M, hidden, out_f16=z)
    return Tokens(z, grid), hs


def _emit_res_block(P, blk, parts, out, prec, x_planar=None):
    """monai UnetResBlock.forward (k3 s1, InstanceNorm no affine, LeakyReLU 0.01).
    x_planar: the block's input as a planar fp32 tensor when it has ONE channel (seg encoder1): conv1 then runs as an exact
    fp32 direct conv and the residual branch norm3(conv3(x)) is evaluated in closed form (engine.conv3_c1 / norm_act_resx)."""
    N, dims = parts[0].N, parts[0].dims
    Co = blk.conv1.conv.weight.shape[0]
    one, zero = P.affine(Co)
    raw1 = P.get_raw(N, Co, dims)
    from .engine import CONV_C1
    c1 = (x_planar is not None and CONV_C1 and Co == 16 and len(parts) == 1 and parts[0].C == 1 and blk.downsample
          and not P.training)
    if c1:
        xstats = P.conv3_c1(x_planar, blk.conv1.conv.weight, None, raw1)
    else:
        P.conv_tc(parts, blk.conv1.conv.weight, 3, 1, prec.conv3, one, zero, False, out_raw=raw1)
    a1 = P.new_act(N, Co, dims, lo=prec.lo)
    P.norm_act(raw1, a1, act="lrelu")
    P.release(raw1)
    raw2 = P.get_raw(N, Co, dims)
    P.conv_tc([a1], blk.conv2.conv.weight, 3, 1, prec.conv3, one, zero, False, out_raw=raw2)
    if c1:
        P.norm_act_resx(raw2, out, x_planar, blk.conv3.conv.weight, xstats, "lrelu")
        P.release(raw2)
        return
    if blk.downsample:
        raw3 = P.get_raw(N, Co, dims)
        P.pointwise([(a, None, None) for a in parts], blk.conv3.conv.weight, None, out_raw=raw3)
        P.norm_act(raw2, out, res=raw3, act_after_res="lrelu")
        P.release(raw3)
    else:
        assert len(parts) == 1
        P.norm_act(raw2, out, res=parts[0], act_after_res="lrelu")
    P.release(raw2)


def _emit_pr_up(P, blk, tokens, out, prec):
    """monai UnetrPrUpBlock.forward: deconv, then num_layer x (deconv -> UnetResBlock)."""
    N = tokens.t.shape[0]
    Co = blk.transp_conv_init.conv.weight.shape[1]
    dims = tuple(2 * g for g in tokens.grid)
    n_layers = len(blk.blocks)
    h = out if n_layers == 0 else P.new_act(N, Co, dims, lo=prec.lo)
    P.deconv2x(tokens, blk.transp_conv_init.conv.weight, h)
    for i, seq in enumerate(blk.blocks):
        dims = tuple(2 * d for d in dims)
        u = P.new_act(N, Co, dims, lo=prec.lo)
        P.deconv2x(h, seq[0].conv.weight, u)
        dst = out if i == n_layers - 1 else P.new_act(N, Co, dims, lo=prec.lo)
        _emit_res_block(P, seq[1], [u], dst, prec)
        h = dst
    return out


def _emit_conv_3_1(P, blk, parts, out, prec):
    """blocks_MDUNet.py:150-157 with conv_block_3 (:64-78) and conv_block_7 (:98-112, eval BatchNorm folded)."""
    N, dims = parts[0].N, parts[0].dims
    act = blk.act
    c3, c7 = blk.conv_3[0].conv, blk.conv_7[0].conv
    C = c3[0].weight.shape[0]
    # --- 3^3 branch: conv -> IN -> ReLU -> conv -> IN -> ReLU -> IN -> act
    raw = P.get_raw(N, C, dims)
    P.conv_tc(parts, c3[0].weight, 3, 1, prec.conv3, *P.affine(C, bias=c3[0].bias), False, out_raw=raw)
    a = P.new_act(N, C, dims, lo=prec.lo)
    P.norm_act(raw, a, act="relu")
    P.release(raw)
    raw3 = P.get_raw(N, C, dims)
    P.conv_tc([a], c3[3].weight, 3, 1, prec.conv3, *P.affine(C, bias=c3[3].bias), False, out_raw=raw3)
    st3 = P.new_stats(N, C)
    fused3 = FUSE_BRANCH_NORM and P.pointwise_tc_ok([C, C], blk.conv[0].weight.shape[0])
    if fused3:
        # relu(IN(raw3)) is never materialised: one statistics-only pass (reads raw3), then the 1^3 conv applies IN + ReLU and
        # the block's IN + act on load, straight from the fp32 conv output (8 instead of 12 bytes per element moved)
        P.norm_act(raw3, None, act="relu", stats_out=st3)
        src3 = (raw3, st3, act, raw3.stats, "relu")
    else:
        y3 = P.new_act(N, C, dims, lo=prec.lo)
        P.norm_act(raw3, y3, act="relu", stats_out=st3)
        P.release(raw3)
        src3 = (y3, st3, act)
    # --- 7^3 branch: conv -> BN -> ReLU -> conv -> BN -> ReLU -> IN -> act
    a7 = P.new_act(N, C, dims, lo=(prec.conv7 != "p1"))
    P.conv_tc(parts, c7[0].weight, 7, 1, prec.conv7, *P.affine(C, bias=c7[0].bias, bn=c7[1]), True, out_act=a7)
    y7 = P.new_act(N, C, dims, lo=prec.lo)
    st7 = P.new_stats(N, C)
    P.conv_tc([a7], c7[3].weight, 7, 1, prec.conv7, *P.affine(C, bias=c7[3].bias, bn=c7[4]), True, out_act=y7, stats=st7)
    # --- cat -> 1^3 conv -> IN -> act
    raw = P.get_raw(N, C, dims)
    P.pointwise([src3, (y7, st7, act)], blk.conv[0].weight, blk.conv[0].bias, out_raw=raw)
    if fused3:
        P.release(raw3)
    P.norm_act(raw, out, act=act)
    P.release(raw)


def _emit_dual_dilated(P, blk, parts, out, prec):
    """blocks_MDUNet.py:206-215: three dilated branches; their last IN+act is applied on load by the 1^3 conv."""
    N, dims = parts[0].N, parts[0].dims
    act = blk.act
    C = blk.conv_3.conv[0].weight.shape[0]
    srcs, raws = [], []
    for br in (blk.conv_3, blk.conv_5, blk.conv_7):
        c = br.conv
        raw = P.get_raw(N, C, dims)
        P.conv_tc(parts, c[0].weight, 3, br.dil, prec.conv3, *P.affine(C, bias=c[0].bias), False, out_raw=raw)
        a = P.new_act(N, C, dims, lo=prec.lo)
        P.norm_act(raw, a, act=act)
        P.release(raw)
        raw = P.get_raw(N, C, dims)
        P.conv_tc([a], c[3].weight, 3, br.dil, prec.conv3, *P.affine(C, bias=c[3].bias), False, out_raw=raw)
        srcs.append((raw, raw.stats, act))
        raws.append(raw)
    o = P.get_raw(N, C, dims)
    P.pointwise(srcs, blk.conv[0].weight, blk.conv[0].bias, out_raw=o)
    for r in raws:
        P.release(r)
    P.norm_act(o, out, act=act)
    P.release(o)


def _emit_conv_3_1_old(P, blk, parts, out, prec):
    """OldModels conv_3_1: every norm is an eval-mode BatchNorm, folded with ReLU into the conv epilogues;
    the closing 1^3 conv has no norm / activation."""
    N, dims = parts[0].N, parts[0].dims
    C = blk.conv.weight.shape[0]
    ys = []
    for br, k, mode in ((blk.conv_3, 3, prec.conv3), (blk.conv_7, 7, prec.conv7)):
        c = br.conv
        lo = mode != "p1"
        a = P.new_act(N, C, dims, lo=lo)
        P.conv_tc(parts, c[0].weight, k, 1, mode, *P.affine(C, bias=c[0].bias, bn=c[1]), True, out_act=a)
        y = P.new_act(N, C, dims, lo=prec.lo)
        P.conv_tc([a], c[3].weight, k, 1, mode, *P.affine(C, bias=c[3].bias, bn=c[4]), True, out_act=y)
        ys.append((y, None, None))
    P.pointwise(ys, blk.conv.weight, blk.conv.bias, out_act=out)


def _emit_basic_block(P, blk, parts, out, prec):
    """monai UnetBasicBlock.forward."""
    N, dims = parts[0].N, parts[0].dims
    Co = blk.conv1.conv.weight.shape[0]
    one, zero = P.affine(Co)
    raw = P.get_raw(N, Co, dims)
    P.conv_tc(parts, blk.conv1.conv.weight, 3, 1, prec.conv3, one, zero, False, out_raw=raw)
    a = P.new_act(N, Co, dims, lo=prec.lo)
    P.norm_act(raw, a, act="lrelu")
    P.release(raw)
    raw = P.get_raw(N, Co, dims)
    P.conv_tc([a], blk.conv2.conv.weight, 3, 1, prec.conv3, one, zero, False, out_raw=raw)
    P.norm_act(raw, out, act="lrelu")
    P.release(raw)


def _emit_up_block(P, blk, inp, skip_slot_pair, out, prec):
    """base_blocks.py:136-141: deconv -> cat(out, skip) -> MultiUnetBasicBlock (or monai UnetrUpBlock.forward)."""
    P.deconv2x(inp, blk.transp_conv.conv.weight, skip_slot_pair[0])
    if isinstance(blk, UnetrUpBlock):
        return _emit_basic_block(P, blk.conv_block, skip_slot_pair, out, prec)
    cov = blk.conv_block.cov_
    if isinstance(cov, conv_3_1_old):
        _emit_conv_3_1_old(P, cov, skip_slot_pair, out, prec)
    elif isinstance(cov, DualDilatedBlock):
        _emit_dual_dilated(P, cov, skip_slot_pair, out, prec)
    else:
        _emit_conv_3_1(P, cov, skip_slot_pair, out, prec)


def _emit_unetr(P, vit, enc_blocks, dec_blocks, parts, taps, prec, prec_deep=None, x_planar=None):
    """Shared UNETR-shaped body of MainSubsetModel.forward (dose_pyfer.py:311-319) and oar_transeg Model.forward
    (oar_transeg.py:171-185).  enc_blocks = (res-block, prup2, prup3, prup4); dec_blocks from coarse to fine.
    prec_deep: precision recipe of the two coarsest levels (1/4 and 1/8 resolution), default = prec."""
    N, dims = parts[0].N, parts[0].dims
    fs = enc_blocks[0].layer.conv1.conv.weight.shape[0]
    precs = [prec, prec, prec_deep or prec, prec_deep or prec]          # per level, fine -> coarse
    sizes = [dims, tuple(d // 2 for d in dims), tuple(d // 4 for d in dims), tuple(d // 8 for d in dims)]
    cats = [P.new_concat(N, [fs << i, fs << i], sizes[i], lo=precs[i].lo) for i in range(4)]     # [deconv out | skip]
    z = _emit_unetr_encoder(P, vit, enc_blocks, parts, taps, precs, [c[1] for c in cats], x_planar=x_planar)
    return _emit_unetr_decoder(P, dec_blocks, z, cats, precs)


def _emit_unetr_encoder(P, vit, enc_blocks, parts, taps, precs, skip_slots, x_planar=None):
    """ViTEncoder.forward (dose_pyfer.py:124-144) / the encoder half of oar_transeg Model.forward: ViT, then the four conv
    skips written into `skip_slots` (usually the second halves of the decoder's concat buffers); returns the z12 tokens."""
    N, dims = parts[0].N, parts[0].dims
    z, hs = _emit_vit(P, vit, parts, N, dims, taps, x_planar=x_planar)
    _emit_res_block(P, enc_blocks[0].layer, parts, skip_slots[0], precs[0], x_planar=x_planar)
    _emit_pr_up(P, enc_blocks[1], hs[taps[0]], skip_slots[1], precs[1])
    _emit_pr_up(P, enc_blocks[2], hs[taps[1]], skip_slots[2], precs[2])
    _emit_pr_up(P, enc_blocks[3], hs[taps[2]], skip_slots[3], precs[3])
    return z


def _emit_unetr_decoder(P, dec_blocks, z, cats, precs):
    """PyMSCDecoder.forward (dose_pyfer.py:232-239) / decoder5..2 of oar_transeg: coarse to fine; cats[lvl] = the
    [deconv out | skip] concat buffer of level lvl.  Returns [dec1 (full res), dec2, dec3, dec4]."""
    decs = []
    inp = z
    for lvl, blk in zip((3, 2, 1, 0), dec_blocks):
        N, C, dims = cats[lvl][1].N, cats[lvl][1].C, cats[lvl][1].dims
        out = P.new_act(N, C, dims, lo=precs[lvl].lo)
        _emit_up_block(P, blk, inp, cats[lvl], out, precs[lvl])
        decs.append(out)
        inp = out
    return decs[::-1]


class _PlanCache:
    """per-module cache of engine.Plan objects keyed by input shape; rebuilt when parameters change."""

    def __init__(self):
        self.plans = {}
        self.version = None

    def get(self, module, key, builder):
        ver = tuple(t._version for t in list(module.parameters()) + list(module.buffers()))
        ptrs = tuple(t.data_ptr() for t in module.parameters())
        ver = (ver, getattr(module, "_dp_epoch", 0))      # bumped by invalidate_plans(): raw-pointer updates (trainers)
        if self.version != (ver, ptrs):
            self.plans.clear()
            self.version = (ver, ptrs)
        if key not in self.plans:
            self.plans[key] = builder()
            if engine.COMPACT:
                self.plans[key].compact()
        return self.plans[key]


def invalidate_plans(model):
    """Mark every cached inference plan under `model` stale.  The trainers update parameters and BatchNorm running
    statistics through raw device pointers (dp_adamw / dp_batch_combine), which moves no tensor version counter."""
    for m in model.modules():
        object.__setattr__(m, "_dp_epoch", getattr(m, "_dp_epoch", 0) + 1)


def _check_input(module, x, channels):
    if module.training:
        raise RuntimeError("dose_prediction_b200: module.forward is the inference path (call .eval()); the training step "
                           "(train-mode forward + loss + backward + AdamW) runs through training.DoseTrainer.step()")
    if not (x.is_cuda and x.dim() == 5):
        raise RuntimeError("expected a CUDA tensor [B,C,D,H,W]; dose_prediction_b200 has no CPU fallback")
    if x.shape[1] != channels:
        raise ValueError(f"expected {channels} input channels, got {x.shape[1]}")
    return x.detach().to(torch.float32).contiguous()


def _check_tensor(module, t, channels=None):
    if module.training:
        raise RuntimeError("dose_prediction_b200: this sub-module's forward is inference only (call .eval()); train-mode "
                           "forward is built for the top-level networks (Model, OARTranseg / TRANSEG)")
    if not (t.is_cuda and t.dim() == 5):
        raise RuntimeError("expected a CUDA tensor [B,C,D,H,W]; dose_prediction_b200 has no CPU fallback")
    if channels is not None and t.shape[1] != channels:
        raise ValueError(f"expected {channels} input channels, got {t.shape[1]}")
    return t.detach().to(torch.float32).contiguous()


def _run_block(module, inputs, builder):
    """sub-block forward: one cached plan per input shapes; builder(P, shapes) must set P.inputs (static fp32 NCDHW
    tensors to fill) and P.result (callable)."""
    if not hasattr(module, "_cache"):
        object.__setattr__(module, "_cache", _PlanCache())
    key = (tuple(tuple(t.shape) for t in inputs), inputs[0].device.index)

    def build():
        P = Plan(inputs[0].device)
        builder(P, [tuple(t.shape) for t in inputs])
        return P
    plan = module._cache.get(module, key, build)
    for dst, src in zip(plan.inputs, inputs):
        dst.copy_(src)
    plan.replay()
    return plan.result()


def _run_single(module, x, plan_fn):
    x = _check_input(module, x, module.in_ch)
    if not hasattr(module, "_cache"):
        object.__setattr__(module, "_cache", _PlanCache())
    plan = module._cache.get(module, (tuple(x.shape), x.device.index), lambda: plan_fn(module, x.shape, x.device))
    plan.x_in.copy_(x)
    plan.replay()
    return plan.result()


def _plan_base_unet(net, shape, device):
    P = Plan(device)
    N, dims = shape[0], tuple(shape[2:])
    P.x_in = P.zeros(tuple(shape), torch.float32)
    x_act = P.new_act(N, net.in_ch, dims, lo=True)
    P.pack_input(P.x_in, x_act)
    out = P.new_act(N, net.list_ch[1], dims, lo=True)
    _emit_base_unet(P, net, x_act, out)
    y = P.zeros((N, net.list_ch[1]) + dims, torch.float32)
    P.unpack(out, y)
    P.result = lambda: y.clone()
    return P


def _up_block_forward(blk, inp, skip):
    Ci, Co = blk.transp_conv.conv.weight.shape[0], blk.transp_conv.conv.weight.shape[1]
    inp, skip = _check_tensor(blk, inp, Ci), _check_tensor(blk, skip, Co)
    if tuple(skip.shape[2:]) != tuple(2 * d for d in inp.shape[2:]) or skip.shape[0] != inp.shape[0]:
        raise ValueError(f"skip {tuple(skip.shape)} does not match the 2x up-sampled input {tuple(inp.shape)}")

    def build(P, shapes):
        N, dims_in, dims = shapes[0][0], shapes[0][2:], shapes[1][2:]
        P.inputs = [P.zeros(shapes[0], torch.float32), P.zeros(shapes[1], torch.float32)]
        a = P.new_act(N, Ci, dims_in, lo=True)
        P.pack_input(P.inputs[0], a)
        cat = P.new_concat(N, [Co, Co], dims, lo=True)
        P.pack_input(P.inputs[1], cat[1])
        out = P.new_act(N, Co, dims, lo=True)
        _emit_up_block(P, blk, a, cat, out, PREC_SEG)             # stand-alone block: the high-precision recipe
        y = P.zeros((N, Co) + tuple(dims), torch.float32)
        P.unpack(out, y)
        P.result = lambda: y.clone()
    return _run_block(blk, [inp, skip], build)


# =========================================================================== DOSE-PYFER
class ViTEncoder(nn.Module):
    def __init__(self, in_channels: int, img_size: Union[Sequence[int], int], feature_size: int = 16,
                 hidden_size: int = 768, mlp_dim: int = 3072, num_heads: int = 12, num_layers: int = 12,
                 pos_embed: str = "conv", norm_name: Union[Tuple, str] = "instance", conv_block: bool = True,
                 res_block: bool = True, dropout_rate: float = 0.0, spatial_dims: int = 3) -> None:
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise ValueError("hidden_size should be divisible by num_heads.")
        if not (conv_block and res_block and norm_name == "instance" and spatial_dims == 3):
            raise NotImplementedError("only conv_block=res_block=True, instance norm, 3-D (the reference config) is built")
        self.num_layers = num_layers
        img_size = _tuple3(img_size)
        self.patch_size = (16, 16, 16)
        self.feat_size = tuple(i // p for i, p in zip(img_size, self.patch_size))
        self.hidden_size = hidden_size
        self.classification = False
        self.vit = ViT(in_channels=in_channels, img_size=img_size, patch_size=self.patch_size, hidden_size=hidden_size,
                       mlp_dim=mlp_dim, num_layers=num_layers, num_heads=num_heads, pos_embed=pos_embed,
                       classification=False, dropout_rate=dropout_rate, spatial_dims=spatial_dims)
        self.skip1 = UnetrBasicBlock(in_channels, feature_size)
        self.skip2 = UnetrPrUpBlock(hidden_size, feature_size * 2, num_layer=2)
        self.skip3 = UnetrPrUpBlock(hidden_size, feature_size * 4, num_layer=1)
        self.skip4 = UnetrPrUpBlock(hidden_size, feature_size * 8, num_layer=0)
        self.proj_axes = (0, spatial_dims + 1) + tuple(d + 1 for d in range(spatial_dims))
        self.proj_view_shape = list(self.feat_size) + [self.hidden_size]
        self.in_ch = in_channels

    def proj_feat(self, x):
        """dose_pyfer.py:118-122: [B,N,C] -> [B,C,h,w,d]."""
        x = x.view([x.size(0)] + self.proj_view_shape)
        return x.permute(self.proj_axes).contiguous()

    def forward(self, x_in):
        """dose_pyfer.py:124-144: x_in [B,C,S,S,S] -> [enc1 [B,fs,S^3], enc2 [B,2fs,(S/2)^3], enc3, enc4, enc5 [B,768,(S/16)^3]]."""
        x = _check_tensor(self, x_in, self.in_ch)
        enc = self

        def build(P, shapes):
            N, dims = shapes[0][0], tuple(shapes[0][2:])
            P.inputs = [P.zeros(shapes[0], torch.float32)]
            x_act = P.new_act(N, enc.in_ch, dims, lo=False)
            P.pack_input(P.inputs[0], x_act)
            fs = enc.skip1.layer.conv1.conv.weight.shape[0]
            i = enc.num_layers // 4
            slots = [P.new_act(N, fs << l, tuple(d >> l for d in dims)) for l in range(4)]
            z = _emit_unetr_encoder(P, enc.vit, (enc.skip1, enc.skip2, enc.skip3, enc.skip4), [x_act], (i, 2 * i, 3 * i),
                                    [PREC_NET_B] * 4, slots)
            outs = []
            for a in slots:
                y = P.zeros((N, a.C) + a.dims, torch.float32)
                P.unpack(a, y)
                outs.append(y)
            P.result = lambda: [o.clone() for o in outs] + [enc.proj_feat(z.t.float())]
        return _run_block(self, [x], build)


class PyMSCDecoder(nn.Module):
    def __init__(self, feature_size: int = 16, hidden_size: int = 768, norm_name: Union[Tuple, str] = "instance",
                 spatial_dims: int = 3, mode_multi: bool = False, act="relu", multiS_conv=True) -> None:
        super().__init__()
        chans = [hidden_size, feature_size * 8, feature_size * 4, feature_size * 2, feature_size]
        for i, name in enumerate(("decoder4", "decoder3", "decoder2", "decoder1")):
            if mode_multi:
                blk = ModifiedUnetrUpBlock(spatial_dims=spatial_dims, in_channels=chans[i], out_channels=chans[i + 1],
                                           upsample_kernel_size=2, act=act, multiS_conv=multiS_conv)
            else:
                blk = UnetrUpBlock(chans[i], chans[i + 1])
            setattr(self, name, blk)

    def forward(self, out_encoder):
        """dose_pyfer.py:232-239: [enc1, enc2, enc3, enc4, enc5] (NCDHW fp32 CUDA) -> [dec1, dec2, dec3, dec4]."""
        if len(out_encoder) != 5:
            raise ValueError("PyMSCDecoder.forward expects the five encoder outputs")
        ts = [_check_tensor(self, t) for t in out_encoder]
        dec = self

        def build(P, shapes):
            N = shapes[0][0]
            P.inputs = [P.zeros(sh, torch.float32) for sh in shapes]
            cats = []
            for l in range(4):
                C, dims = shapes[l][1], tuple(shapes[l][2:])
                cat = P.new_concat(N, [C, C], dims)
                P.pack_input(P.inputs[l], cat[1])
                cats.append(cat)
            hidden, grid = shapes[4][1], tuple(shapes[4][2:])
            tok = P.zeros((N, grid[0] * grid[1] * grid[2], hidden), torch.float16)       # proj_feat inverted: tokens
            P.add_py(lambda: tok.copy_(P.inputs[4].flatten(2).transpose(1, 2)))
            decs = _emit_unetr_decoder(P, (dec.decoder4, dec.decoder3, dec.decoder2, dec.decoder1), Tokens(tok, grid), cats,
                                       [PREC_NET_B] * 4)
            outs = []
            for a in decs:
                y = P.zeros((N, a.C) + a.dims, torch.float32)
                P.unpack(a, y)
                outs.append(y)
            P.result = lambda: [o.clone() for o in outs]
        return _run_block(self, ts, build)


class MainSubsetModel(nn.Module):
    def __init__(self, in_ch, out_ch, img_size, feature_size: int = 16, hidden_size: int = 768, mlp_dim: int = 3072,
                 num_heads: int = 12, num_layers: int = 12, conv_block: bool = True, res_block: bool = True,
                 dropout_rate: float = 0.0, mode_multi_dec=False, act="relu", multiS_conv=True):
        super().__init__()
        self.in_ch, self.out_ch = in_ch, out_ch
        self.encoder = ViTEncoder(in_channels=in_ch, img_size=img_size, feature_size=feature_size, hidden_size=hidden_size,
                                  mlp_dim=mlp_dim, num_heads=num_heads, num_layers=num_layers, pos_embed="perceptron",
                                  norm_name="instance", res_block=res_block, conv_block=conv_block, dropout_rate=dropout_rate)
        self.decoder = PyMSCDecoder(feature_size=feature_size, hidden_size=hidden_size, mode_multi=mode_multi_dec, act=act,
                                    multiS_conv=multiS_conv)

        def to_out(in_feature):
            return nn.Sequential(nn.Conv3d(in_feature, out_ch, kernel_size=1, padding=0, bias=True))

        self.dose_convertors = nn.ModuleList([to_out(feature_size)])
        for i in range(1, 4):
            self.dose_convertors.append(to_out(int(feature_size * np.power(2, i))))
        self.out = nn.Sequential(nn.Conv3d(feature_size, out_ch, kernel_size=1, padding=0, bias=True))

    def update_config(self, config_hparam):
        self.encoder.hidden_size = config_hparam["hidden_size"]
        self.encoder.num_layers = config_hparam["hidden_size"]

    def forward(self, x):
        return _run_single(self, x, _plan_main_subset)


def _emit_main_subset(P, net, parts):
    enc, dec = net.encoder, net.decoder
    i = enc.num_layers // 4
    decs = _emit_unetr(P, enc.vit, (enc.skip1, enc.skip2, enc.skip3, enc.skip4),
                       (dec.decoder4, dec.decoder3, dec.decoder2, dec.decoder1), parts, (i, 2 * i, 3 * i), PREC_NET_B)
    outs = []
    for d, conv in zip(decs, net.dose_convertors):
        y = P.zeros((d.N, net.out_ch) + d.dims, torch.float32)
        P.head(d, conv[0].weight, conv[0].bias, y)
        outs.append(y)
    return outs


def _plan_main_subset(net, shape, device):
    P = Plan(device)
    N, dims = shape[0], tuple(shape[2:])
    P.x_in = P.zeros(tuple(shape), torch.float32)
    x_act = P.new_act(N, net.in_ch, dims, lo=False)
    P.pack_input(P.x_in, x_act)
    outs = _emit_main_subset(P, net, [x_act])
    P.result = lambda: [o.clone() for o in outs]
    return P


class Model(nn.Module):
    """DOSE-PYFER cascade: C3D net_A -> cat -> ViT/multi-scale-conv net_B (dose_pyfer.py:325-360)."""

    def __init__(self, in_ch, out_ch, list_ch_A, feature_size=16, img_size=(128, 128, 128), num_layers=8, num_heads=6,
                 act="mish", mode_multi_dec=True, multiS_conv=True):
        super().__init__()
        self.in_ch, self.out_ch = in_ch, out_ch
        self.net_A = BaseUNet(in_ch, list_ch_A)
        self.net_B = MainSubsetModel(in_ch=in_ch + list_ch_A[1], out_ch=out_ch, feature_size=feature_size,
                                     img_size=img_size, num_layers=num_layers, num_heads=num_heads, act=act,
                                     mode_multi_dec=mode_multi_dec, multiS_conv=multiS_conv)
        self.conv_out_A = nn.Conv3d(list_ch_A[1], out_ch, kernel_size=1, padding=0, bias=True)

    def forward(self, x):
        """x [B,9,S,S,S] fp32 CUDA -> [output_A [B,1,S^3], [dose_S, dose_S/2, dose_S/4, dose_S/8]].
        eval(): the inference plan.  train(): the same call as one torch.autograd node (training.autograd_forward), so
        `loss.backward()` / `optimizer.step()` of Pyfer.training_step (train_light_pyfer.py:122-143) work unchanged."""
        if self.training:
            from . import training
            return training.autograd_forward(self, x)
        return _run_single(self, x, lambda m, shape, dev: plan_dose_pyfer(m, shape, dev))


def emit_dose_pyfer(P, model, x_act, a_out):
    """dose_pyfer.py:355-360 on an already-packed input; (a_out | x_act) share one concat buffer."""
    _emit_base_unet(P, model.net_A, x_act, a_out)
    N, dims = x_act.N, x_act.dims
    out_A = P.zeros((N, model.out_ch) + dims, torch.float32)
    P.head(a_out, model.conv_out_A.weight, model.conv_out_A.bias, out_A)
    outs = _emit_main_subset(P, model.net_B, [a_out, x_act])
    return out_A, outs


def plan_dose_pyfer(model, shape, device, external_input=False):
    P = Plan(device)
    N, dims = shape[0], tuple(shape[2:])
    a_out, x_act = P.new_concat(N, [model.net_A.list_ch[1], model.in_ch], dims, lo=True)
    P.x_act = x_act
    if not external_input:
        P.x_in = P.zeros(tuple(shape), torch.float32)
        P.pack_input(P.x_in, x_act)
    out_A, outs = emit_dose_pyfer(P, model, x_act, a_out)
    P.outputs = (out_A, outs)
    P.result = lambda: [out_A.clone(), [o.clone() for o in outs]]
    return P


def create_pretrained_unet(ckpt_file, in_ch, out_ch, list_ch_A, feature_size, img_size, num_layers=8, num_heads=6,
                           act="mish", mode_multi_dec=True, multiS_conv=True):
    """dose_pyfer.py:363-407: load the C3D checkpoint's matching keys (net_A.*, conv_out_A.*), strict=False."""
    pretrain = torch.load(ckpt_file, map_location="cpu")
    net = Model(in_ch, out_ch, list_ch_A, feature_size=feature_size, img_size=img_size, num_layers=num_layers,
                num_heads=num_heads, act=act, mode_multi_dec=mode_multi_dec, multiS_conv=multiS_conv)
    net_dict = net.state_dict()
    sd = pretrain["network_state_dict"]
    missing = tuple({k for k in net_dict.keys() if k not in sd})
    print(f"missing in pretrained: {len(missing)}")
    inside = tuple({k for k in sd if k in net_dict.keys()})
    print(f"inside pretrained: {len(inside)}")
    unused = tuple({k for k in sd if k not in net_dict.keys()})
    print(f"unused pretrained: {len(unused)}")
    net.load_state_dict({k: v for k, v in sd.items() if k in net_dict.keys()}, strict=False)
    return net, inside


# =========================================================================== OAR-TRANSEG
class OARTranseg(nn.Module):
    """OARSegmentation/Models/Networks/oar_transeg.py:14 `Model` (exported as oar_transeg.Model)."""

    def __init__(self, in_channels: int, out_channels: int, img_size: Union[Sequence[int], int], feature_size: int = 16,
                 hidden_size: int = 768, mlp_dim: int = 3072, num_heads: int = 12, pos_embed: str = "conv",
                 norm_name: Union[Tuple, str] = "instance", conv_block: bool = True, res_block: bool = True,
                 dropout_rate: float = 0.0, spatial_dims: int = 3, _old_blocks: bool = False) -> None:
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise ValueError("hidden_size should be divisible by num_heads.")
        if not (conv_block and res_block and norm_name == "instance" and spatial_dims == 3):
            raise NotImplementedError("only conv_block=res_block=True, instance norm, 3-D (the reference config) is built")
        self.in_ch, self.out_channels = in_channels, out_channels
        self.num_layers = 12
        img_size = _tuple3(img_size)
        self.patch_size = (16, 16, 16)
        self.feat_size = tuple(i // p for i, p in zip(img_size, self.patch_size))
        self.hidden_size = hidden_size
        self.classification = False
        self.vit = ViT(in_channels=in_channels, img_size=img_size, patch_size=self.patch_size, hidden_size=hidden_size,
                       mlp_dim=mlp_dim, num_layers=self.num_layers, num_heads=num_heads, pos_embed=pos_embed,
                       classification=False, dropout_rate=dropout_rate, spatial_dims=spatial_dims)
        self.encoder1 = UnetrBasicBlock(in_channels, feature_size)
        self.encoder2 = UnetrPrUpBlock(hidden_size, feature_size * 2, num_layer=2)
        self.encoder3 = UnetrPrUpBlock(hidden_size, feature_size * 4, num_layer=1)
        self.encoder4 = UnetrPrUpBlock(hidden_size, feature_size * 8, num_layer=0)
        self.decoder5 = ModifiedUnetrUpBlock(spatial_dims, hidden_size, feature_size * 8, 2, old=_old_blocks)
        self.decoder4 = ModifiedUnetrUpBlock(spatial_dims, feature_size * 8, feature_size * 4, 2, old=_old_blocks)
        self.decoder3 = ModifiedUnetrUpBlock(spatial_dims, feature_size * 4, feature_size * 2, 2, old=_old_blocks)
        self.decoder2 = ModifiedUnetrUpBlock(spatial_dims, feature_size * 2, feature_size, 2, old=_old_blocks)
        self.out = ModifiedUnetOutBlock(spatial_dims=spatial_dims, in_channels=feature_size, out_channels=out_channels)
        self.proj_axes = (0, spatial_dims + 1) + tuple(d + 1 for d in range(spatial_dims))
        self.proj_view_shape = list(self.feat_size) + [self.hidden_size]

    def forward(self, x_in):
        """x_in [B,in_channels,S,S,S] fp32 CUDA -> logits [B,out_channels,S,S,S] fp32.  train(): one torch.autograd node
        (training.autograd_forward_seg), so Transeg.training_step's loss.backward() works unchanged."""
        if self.training:
            from . import training
            return training.autograd_forward_seg(self, x_in)
        return _run_single(self, x_in, lambda m, shape, dev: plan_oar_transeg(m, shape, dev))


class TRANSEG(OARTranseg):
    """OARSegmentation/OldModels/Networks/oar_transeg.py:14 — the variant train_light_transeg.py:20,110 and the
    cascade LinkedNet (train_light_linked_model.py:89) instantiate; BatchNorm in both decoder branches, bare 1^3."""

    def __init__(self, in_channels: int, out_channels: int, img_size: Union[Sequence[int], int], feature_size: int = 16,
                 hidden_size: int = 768, mlp_dim: int = 3072, num_heads: int = 12, pos_embed: str = "conv",
                 norm_name: Union[Tuple, str] = "instance", conv_block: bool = True, res_block: bool = True,
                 dropout_rate: float = 0.0, spatial_dims: int = 3) -> None:
        super().__init__(in_channels, out_channels, img_size, feature_size, hidden_size, mlp_dim, num_heads, pos_embed,
                         norm_name, conv_block, res_block, dropout_rate, spatial_dims, _old_blocks=True)


def emit_oar_transeg(P, model, x_act, x_planar=None):
    """x_planar: the same input as a planar fp32 [N,1,D,H,W] tensor, when the caller has one (lets the one-channel
    encoder1 convs run in exact fp32 straight from it)."""
    decs = _emit_unetr(P, model.vit, (model.encoder1, model.encoder2, model.encoder3, model.encoder4),
                       (model.decoder5, model.decoder4, model.decoder3, model.decoder2), [x_act], (3, 6, 9), PREC_SEG,
                       prec_deep=PREC_SEG_DEEP, x_planar=x_planar if model.in_ch == 1 else None)
    d = decs[0]
    logits = P.zeros((d.N, model.out_channels) + d.dims, torch.float32)
    P.head(d, model.out.conv.conv.weight, model.out.conv.conv.bias, logits)
    return logits


def plan_oar_transeg(model, shape, device):
    P = Plan(device)
    N, dims = shape[0], tuple(shape[2:])
    P.x_in = P.zeros(tuple(shape), torch.float32)
    x_act = P.new_act(N, model.in_ch, dims, lo=True)
    P.pack_input(P.x_in, x_act)
    logits = emit_oar_transeg(P, model, x_act, x_planar=P.x_in)
    P.outputs = logits
    P.result = lambda: logits.clone()
    return P
