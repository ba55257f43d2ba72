"""TEST INFRASTRUCTURE — fp32 functional restatement of the reference hot path.

A state_dict in the REFERENCE layout plus an input goes in; the tensors the
reference modules would return come out.  Plain torch.nn.functional on CPU, no modules, so it
travels to the GPU box (where /root/reference does not exist).  Pinned against the reference's own
modules by tests/test_oracle.py (build container) and tests/golden/*.npz (everywhere).

Each function cites the reference lines it restates (paths relative to the reference root;
"monai:" = monai==0.7.0, an un-vendored dependency, see oracle/monai_compat).
"""
import math
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
EPS = 1e-5

# --------------------------------------------------------------------------- operand-precision emulation
# EMU = None reproduces the reference in exact fp32.  Setting EMU to a callable
#   EMU(tag: str, kind: str) -> (act_mode, weight_mode),  modes in {"f32", "f16", "hilo"}
# rounds the operands of every contraction the way the CUDA path stores/feeds them (fp16, or an
# fp16 hi+lo pair ~ 22 bits) while keeping fp32 accumulation; used by oracle/precision_probe.py to
# freeze the per-layer precision recipe on CPU before any GPU time is spent.
EMU = None


def _round(x, mode):
    if mode == "f32":
        return x
    hi = x.half().float()
    if mode == "f16":
        return hi
    if mode == "hilo":
        return hi + (x - hi).half().float()
    raise ValueError(mode)


def _q(x, w, tag, kind):
    if EMU is None:
        return x, w
    am, wm = EMU(tag, kind)
    return _round(x, am), _round(w, wm)


# --------------------------------------------------------------------------- small helpers
def _act(x, name):
    if name == "relu":
        return F.relu(x)
    if name == "mish":
        return F.mish(x)
    if name == "lrelu":
        return F.leaky_relu(x, 0.01)
    raise ValueError(name)


def _inorm(x, w=None, b=None):
    """nn.InstanceNorm3d: per (n,c) biased variance, eps 1e-5, instance statistics in eval too."""
    return F.instance_norm(x, weight=w, bias=b, eps=EPS)


# BN_TRAIN = None: eval-mode BatchNorm3d (running statistics).  BN_TRAIN = {}: train mode (batch statistics,
# momentum 0.1); the updated running statistics are collected in the dict under the state_dict keys.
BN_TRAIN = None


def _bn_eval(sd, p, x):
    if BN_TRAIN is not None:
        rm, rv = sd[p + "running_mean"].clone(), sd[p + "running_var"].clone()
        y = F.batch_norm(x, rm, rv, sd[p + "weight"], sd[p + "bias"], training=True, momentum=0.1, eps=EPS)
        BN_TRAIN[p + "running_mean"], BN_TRAIN[p + "running_var"] = rm, rv
        return y
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"],
                        sd[p + "bias"], training=False, eps=EPS)


def _conv(sd, p, x, stride=1, padding=0, dilation=1):
    w = sd[p + "weight"]
    kind = "conv%d%s" % (w.shape[2], "s2" if stride != 1 else "")
    x, w = _q(x, w, p, kind)
    return F.conv3d(x, w, sd.get(p + "bias"), stride=stride, padding=padding, dilation=dilation)


def _linear(sd, p, x, bias=True):
    x, w = _q(x, sd[p + "weight"], p, "linear")
    return F.linear(x, w, sd[p + "bias"] if bias else None)


def _deconv2(sd, p, x):
    """ConvTranspose3d k=2 s=2 no bias (monai get_conv_layer(is_transposed=True))."""
    x, w = _q(x, sd[p + "weight"], p, "deconv")
    return F.conv_transpose3d(x, w, None, stride=2)


# --------------------------------------------------------------------------- net_A (C3D U-Net)
def _single_conv(sd, p, x, stride=1):
    """c3d.py:11-22 SingleConv / :25-33 UpConv.conv: conv3(+bias) -> IN(affine) -> ReLU."""
    y = _conv(sd, p + "0.", x, stride=stride, padding=1)
    return F.relu(_inorm(y, sd[p + "1.weight"], sd[p + "1.bias"]))


def c3d_base_unet(sd: SD, p: str, x):
    """c3d.py:118-149 BaseUNet.forward = Decoder(Encoder(x)); Encoder :41-72, Decoder :75-115."""
    enc = []
    h = x
    for s in range(1, 6):
        q = f"{p}encoder.encoder_{s}."
        h = _single_conv(sd, q + "0.single_conv.", h, stride=1 if s == 1 else 2)
        h = _single_conv(sd, q + "1.single_conv.", h)
        enc.append(h)
    h = enc[4]
    for s in (4, 3, 2, 1):
        h, _ = _q(h, h, f"{p}decoder.upconv_{s}.", "upsample")
        up = F.interpolate(h, scale_factor=2, mode="trilinear", align_corners=True)  # c3d.py:36
        up = _single_conv(sd, f"{p}decoder.upconv_{s}.conv.", up)
        h = torch.cat((up, enc[s - 1]), dim=1)
        h = _single_conv(sd, f"{p}decoder.decoder_conv_{s}.0.single_conv.", h)
        if s != 1:
            h = _single_conv(sd, f"{p}decoder.decoder_conv_{s}.1.single_conv.", h)
    return h


# --------------------------------------------------------------------------- monai ViT
def vit(sd: SD, p: str, x, num_layers: int, num_heads: int, patch: int = 16):
    """monai: nets/vit.py ViT.forward with pos_embed='perceptron' (dose_pyfer.py:55-67,
    oar_transeg.py:79-91).  Returns (LN(x_L), [x_1..x_L])."""
    b, c = x.shape[:2]
    g = [s // patch for s in x.shape[2:]]
    # einops 'b c (h p1) (w p2) (d p3) -> b (h w d) (p1 p2 p3 c)'
    t = x.reshape(b, c, g[0], patch, g[1], patch, g[2], patch)
    t = t.permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(b, g[0] * g[1] * g[2], patch ** 3 * c)
    q = p + "patch_embedding."
    if q + "patch_embeddings.weight" in sd:       # pos_embed="conv": Conv3d(kernel=stride=patch), flatten(2).transpose
        xq, wq = _q(x, sd[q + "patch_embeddings.weight"], q + "patch_embeddings.", "linear")
        t = F.conv3d(xq, wq, sd[q + "patch_embeddings.bias"], stride=patch).flatten(2).transpose(-1, -2)
    else:
        t = _linear(sd, q + "patch_embeddings.1.", t)
    t = t + sd[q + "position_embeddings"]
    hidden = t.shape[-1]
    d = hidden // num_heads
    hs = []
    for i in range(num_layers):
        q = f"{p}blocks.{i}."
        y = F.layer_norm(t, (hidden,), sd[q + "norm1.weight"], sd[q + "norm1.bias"], EPS)
        qkv = _linear(sd, q + "attn.qkv.", y, bias=False)                # no bias
        # 'b h (qkv l d) -> qkv b l h d'
        qkv = qkv.reshape(b, -1, 3, num_heads, d).permute(2, 0, 3, 1, 4)
        if EMU is None:
            att = torch.softmax(qkv[0] @ qkv[1].transpose(-1, -2) * d ** -0.5, dim=-1)
            y = (att @ qkv[2]).permute(0, 2, 1, 3).reshape(b, -1, hidden)
        else:   # the CUDA path folds the scale into q before rounding q/k/v/p to fp16
            qq, kk = _q(qkv[0] * d ** -0.5, qkv[1], q + "attn.qk.", "attn")
            att = torch.softmax(qq @ kk.transpose(-1, -2), dim=-1)
            att, vv = _q(att, qkv[2], q + "attn.pv.", "attn")
            y = (att @ vv).permute(0, 2, 1, 3).reshape(b, -1, hidden)
        t = t + _linear(sd, q + "attn.out_proj.", y)
        y = F.layer_norm(t, (hidden,), sd[q + "norm2.weight"], sd[q + "norm2.bias"], EPS)
        y = F.gelu(_linear(sd, q + "mlp.linear1.", y))
        t = t + _linear(sd, q + "mlp.linear2.", y)
        hs.append(t)
    return F.layer_norm(t, (hidden,), sd[p + "norm.weight"], sd[p + "norm.bias"], EPS), hs


def proj_feat(t, grid: Sequence[int]):
    """dose_pyfer.py:115-122 / oar_transeg.py:165-169: [B,N,C] -> [B,C,h,w,d]."""
    b, _, c = t.shape
    return t.view(b, grid[0], grid[1], grid[2], c).permute(0, 4, 1, 2, 3).contiguous()


# --------------------------------------------------------------------------- monai UNETR blocks
def unet_res_block(sd: SD, p: str, x):
    """monai: dynunet_block.UnetResBlock.forward, k=3 s=1, instance norm (no affine), LeakyReLU 0.01;
    conv3/norm3 on the residual only when in_channels != out_channels."""
    w1 = sd[p + "conv1.conv.weight"]
    y = _act(_inorm(_conv(sd, p + "conv1.conv.", x, padding=1)), "lrelu")
    y = _inorm(_conv(sd, p + "conv2.conv.", y, padding=1))
    r = x
    if w1.shape[0] != w1.shape[1]:
        r = _inorm(_conv(sd, p + "conv3.conv.", x))
    elif EMU is not None:
        r = _q(x, x, p + "residual.", "store")[0]
    return _act(y + r, "lrelu")


def unet_basic_block(sd: SD, p: str, x):
    """monai: dynunet_block.UnetBasicBlock.forward (k3 s1, instance norm, LeakyReLU 0.01)."""
    y = _act(_inorm(_conv(sd, p + "conv1.conv.", x, padding=1)), "lrelu")
    return _act(_inorm(_conv(sd, p + "conv2.conv.", y, padding=1)), "lrelu")


def unetr_pr_up_block(sd: SD, p: str, x, num_layer: int):
    """monai: unetr_block.UnetrPrUpBlock.forward (conv_block=True, res_block=True)."""
    x = _deconv2(sd, p + "transp_conv_init.conv.", x)
    for i in range(num_layer):
        x = _deconv2(sd, f"{p}blocks.{i}.0.conv.", x)
        x = unet_res_block(sd, f"{p}blocks.{i}.1.", x)
    return x


# --------------------------------------------------------------------------- multi-scale decoder
def conv_3_1(sd: SD, p: str, x, act: str):
    """OARSegmentation/Models/Nets/blocks_MDUNet.py:132-157 (conv_block_3 :64-78 always ReLU inside,
    conv_block_7 :98-112 BatchNorm+ReLU inside)."""
    q = p + "conv_3.0.conv."
    y = F.relu(_inorm(_conv(sd, q + "0.", x, padding=1)))
    y = F.relu(_inorm(_conv(sd, q + "3.", y, padding=1)))
    x3 = _act(_inorm(y), act)
    q = p + "conv_7.0.conv."
    y = F.relu(_bn_eval(sd, q + "1.", _conv(sd, q + "0.", x, padding=3)))
    y = F.relu(_bn_eval(sd, q + "4.", _conv(sd, q + "3.", y, padding=3)))
    x7 = _act(_inorm(y), act)
    y = _conv(sd, p + "conv.0.", torch.cat((x3, x7), dim=1))
    return _act(_inorm(y), act)


def dual_dilated_block(sd: SD, p: str, x, act: str):
    """blocks_MDUNet.py:194-215 DualDilatedBlock (multiS_conv=False): dilation 1/2/3 k3 branches."""
    outs = []
    for name, dil in (("conv_3", 1), ("conv_5", 2), ("conv_7", 3)):
        q = f"{p}{name}.conv."
        y = _act(_inorm(_conv(sd, q + "0.", x, padding=dil, dilation=dil)), act)
        y = _act(_inorm(_conv(sd, q + "3.", y, padding=dil, dilation=dil)), act)
        outs.append(y)
    y = _conv(sd, p + "conv.0.", torch.cat(outs, dim=1))
    return _act(_inorm(y), act)


def conv_3_1_old(sd: SD, p: str, x):
    """OARSegmentation/OldModels/Nets/blocks_MDUNet.py:132-147 (conv_block_3 :64-78 and conv_block_7 :98-112 both
    conv->BN->ReLU twice; bare 1^3 conv, no norm/activation after it)."""
    outs = []
    for name, pad in (("conv_3", 1), ("conv_7", 3)):
        q = f"{p}{name}.conv."
        y = F.relu(_bn_eval(sd, q + "1.", _conv(sd, q + "0.", x, padding=pad)))
        y = F.relu(_bn_eval(sd, q + "4.", _conv(sd, q + "3.", y, padding=pad)))
        outs.append(y)
    return _conv(sd, p + "conv.", torch.cat(outs, dim=1))


def modified_unetr_up_block(sd: SD, p: str, inp, skip, act: str, multiS_conv: bool = True, old: bool = False):
    """OARSegmentation/Models/Nets/base_blocks.py:136-141 (OldModels/Nets/base_blocks.py:127-132 when old)."""
    out = torch.cat((_deconv2(sd, p + "transp_conv.conv.", inp), skip), dim=1)
    if p + "conv_block.conv1.conv.weight" in sd:   # monai UnetrUpBlock (PyMSCDecoder mode_multi=False, dose_pyfer.py:164-171)
        return unet_basic_block(sd, p + "conv_block.", out)
    if old:
        return conv_3_1_old(sd, p + "conv_block.cov_.", out)
    blk = conv_3_1 if multiS_conv else dual_dilated_block
    return blk(sd, p + "conv_block.cov_.", out, act)


# --------------------------------------------------------------------------- whole networks
def oar_transeg_forward(sd: SD, x, num_heads: int = 12, num_layers: int = 12, old: bool = False):
    """OARSegmentation/Models/Networks/oar_transeg.py:171-185 (old=True: OldModels/Networks/oar_transeg.py TRANSEG)."""
    grid = [s // 16 for s in x.shape[2:]]
    z, hs = vit(sd, "vit.", x, num_layers, num_heads)
    enc1 = unet_res_block(sd, "encoder1.layer.", x)
    enc2 = unetr_pr_up_block(sd, "encoder2.", proj_feat(hs[3], grid), 2)
    enc3 = unetr_pr_up_block(sd, "encoder3.", proj_feat(hs[6], grid), 1)
    enc4 = unetr_pr_up_block(sd, "encoder4.", proj_feat(hs[9], grid), 0)
    d = modified_unetr_up_block(sd, "decoder5.", proj_feat(z, grid), enc4, "relu", old=old)
    d = modified_unetr_up_block(sd, "decoder4.", d, enc3, "relu", old=old)
    d = modified_unetr_up_block(sd, "decoder3.", d, enc2, "relu", old=old)
    d = modified_unetr_up_block(sd, "decoder2.", d, enc1, "relu", old=old)
    return _conv(sd, "out.conv.conv.", d)


def main_subset_forward(sd: SD, p: str, x, num_layers: int, num_heads: int, act: str,
                        multiS_conv: bool = True):
    """dose_pyfer.py:311-319 MainSubsetModel.forward (ViTEncoder :124-144, PyMSCDecoder :232-239)."""
    grid = [s // 16 for s in x.shape[2:]]
    e = p + "encoder."
    i = num_layers // 4
    z, hs = vit(sd, e + "vit.", x, num_layers, num_heads)
    enc1 = unet_res_block(sd, e + "skip1.layer.", x)
    enc2 = unetr_pr_up_block(sd, e + "skip2.", proj_feat(hs[i], grid), 2)
    enc3 = unetr_pr_up_block(sd, e + "skip3.", proj_feat(hs[2 * i], grid), 1)
    enc4 = unetr_pr_up_block(sd, e + "skip4.", proj_feat(hs[3 * i], grid), 0)
    d = p + "decoder."
    dec4 = modified_unetr_up_block(sd, d + "decoder4.", proj_feat(z, grid), enc4, act, multiS_conv)
    dec3 = modified_unetr_up_block(sd, d + "decoder3.", dec4, enc3, act, multiS_conv)
    dec2 = modified_unetr_up_block(sd, d + "decoder2.", dec3, enc2, act, multiS_conv)
    dec1 = modified_unetr_up_block(sd, d + "decoder1.", dec2, enc1, act, multiS_conv)
    return [_conv(sd, f"{p}dose_convertors.{j}.0.", t) for j, t in enumerate((dec1, dec2, dec3, dec4))]


def dose_pyfer_forward(sd: SD, x, num_layers: int = 8, num_heads: int = 6, act: str = "mish",
                       multiS_conv: bool = True):
    """dose_pyfer.py:355-360 Model.forward -> [output_A, [dose_S, dose_S/2, dose_S/4, dose_S/8]]."""
    a = c3d_base_unet(sd, "net_A.", x)
    outs = main_subset_forward(sd, "net_B.", torch.cat((a, x), dim=1), num_layers, num_heads, act,
                               multiS_conv)
    return [_conv(sd, "conv_out_A.", a), outs]


# --------------------------------------------------------------------------- cascade hand-off
def handoff(logits, ptv, ct):
    """train_light_linked_model.py:156-167 + OARSegmentation/config.py:70:
    argmax over 8 classes -> one-hot -> permute (0,3,2,1) -> drop background -> cat(ptv, oars, ct^T).
    logits [1,8,X,Y,Z], ptv/ct [1,1,X,Y,Z] -> structures [1,9,...]."""
    assert logits.shape[0] == 1
    idx = torch.argmax(logits[0], dim=0, keepdim=True)
    onehot = torch.zeros_like(logits[0]).scatter_(0, idx, 1.0)
    oars = onehot.permute(0, 3, 2, 1).unsqueeze(0)[:, 1:]
    ct_t = ct.permute(0, 1, 4, 3, 2)
    return torch.cat((ptv, oars, ct_t), dim=1)


def sliding_window_logits(sd: SD, ct, roi: int = 96, sw_batch: int = 4, overlap: float = 0.25, **kw):
    """monai: inferers.sliding_window_inference (constant blending) around the seg net, as called at
    train_light_linked_model.py:152-154."""
    s = ct.shape[2:]
    iv = [roi if roi == n else max(int(roi * (1 - overlap)), 1) for n in s]
    starts = []
    for n, step in zip(s, iv):
        cnt = int(math.ceil((n - roi) / step)) + 1 if n > roi else 1
        starts.append([min(k * step, n - roi) for k in range(cnt)])
    wins = [(a, b, c) for a in starts[0] for b in starts[1] for c in starts[2]]
    out = cnt_map = None
    for g in range(0, len(wins), sw_batch):
        chunk = wins[g:g + sw_batch]
        data = torch.cat([ct[:, :, a:a + roi, b:b + roi, c:c + roi] for a, b, c in chunk])
        prob = oar_transeg_forward(sd, data, **kw)
        if out is None:
            out = torch.zeros((1, prob.shape[1]) + tuple(s))
            cnt_map = torch.zeros_like(out)
        for j, (a, b, c) in enumerate(chunk):
            out[:, :, a:a + roi, b:b + roi, c:c + roi] += prob[j]
            cnt_map[:, :, a:a + roi, b:b + roi, c:c + roi] += 1.0
    return out / cnt_map


# --------------------------------------------------------------------------- training loss
def gen_loss(predictions, gt, delta1=10.0, delta2=8.0, freeze=True):
    """DosePrediction/Train/loss.py:69-119 GenLoss.forward(mode='train', casecade=True, freez=freeze,
    huber=False), with downSample :57-67 (trilinear align_corners GT, nearest-exact mask); freez=False adds
    0.5 * L1 of net_A's own prediction over the possible-dose mask (:114-115)."""
    gt_dose, mask = gt[:, 0:1], gt[:, 1:]
    preds = predictions[1]
    size = gt.shape[-1]
    l_ds = 0.0
    for i, p_i in enumerate(preds[1:], start=1):
        dim = size // 2 ** i
        g_i = F.interpolate(gt_dose, size=(dim,) * 3, mode="trilinear", align_corners=True)
        m_i = F.interpolate(mask, size=(dim,) * 3, mode="nearest-exact")
        sel = m_i > 0
        l_ds = l_ds + F.l1_loss(p_i[sel], g_i[sel])
    l_ds = l_ds / (len(preds) - 1)
    sel = mask > 0
    loss = delta1 * F.l1_loss(preds[0][sel], gt_dose[sel]) + delta2 * l_ds
    if not freeze:
        loss = loss + 0.5 * F.l1_loss(predictions[0][sel], gt_dose[sel])
    return loss


def dose_pyfer_train_step(sd: SD, x, gt, lr=1e-4, weight_decay=1e-4, delta1=10.0, delta2=8.0, betas=(0.9, 0.999),
                          eps=1e-8, probe=None, freeze=True, **kw):
    """Pyfer.training_step + one optimizer step (DosePrediction/Train/train_light_pyfer.py:85-88,122-143,194-197):
    train-mode forward (freeze=True: net_A.* / conv_out_A.* get no gradient; freeze=False, the other value of the
    constructor flag :61-88: every parameter trains and GenLoss gains its net_A term), GenLoss, autograd, AdamW with fp32
    state (the reference's bnb Adam8bit quantises the same update's state to 8 bit; not restated).
    Returns (loss, {name: grad}, {name: updated parameter or running statistic}, forward outputs)."""
    global BN_TRAIN
    is_buffer = lambda k: k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked")
    leaf = {k: v.detach().clone() for k, v in sd.items()}
    train_keys = [k for k in leaf if (not freeze or not (k.startswith("net_A") or k.startswith("conv_out_A")))
                  and not is_buffer(k) and leaf[k].is_floating_point()]
    for k in train_keys:
        leaf[k].requires_grad_(True)
    BN_TRAIN = {}
    try:
        out = dose_pyfer_forward(leaf, x, **kw)
        loss = gen_loss(out, gt, delta1, delta2, freeze=freeze)
        if probe is not None:      # linear loss sum <pred_i, R_i>: smooth gradients for backward-pass parity tests
            lin = sum((o * r).sum() for o, r in zip(out[1], probe))
            if len(probe) > len(out[1]):          # freeze=False: a fifth probe tensor for net_A's own prediction
                lin = lin + (out[0] * probe[len(out[1])]).sum()
            lin.backward()
        else:
            loss.backward()
        new_stats = dict(BN_TRAIN)
    finally:
        BN_TRAIN = None
    # parameters the forward never touches keep grad None: torch.optim.AdamW (and bnb Adam8bit) skip them entirely
    # (no weight decay either); they are reported with a zero gradient
    grads = {k: (leaf[k].grad if leaf[k].grad is not None else torch.zeros_like(leaf[k])) for k in train_keys}
    params = [leaf[k] for k in train_keys]
    torch.optim.AdamW(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay).step()
    new = {k: leaf[k].detach() for k in train_keys}
    new.update(new_stats)
    outs = [out[0].detach(), [o.detach() for o in out[1]]]
    return loss.detach(), grads, new, outs


def dice_ce_loss(logits, label, smooth=1e-5):
    """monai==0.7.0 losses.DiceCELoss(to_onehot_y=True, softmax=True) as called at
    OARSegmentation/train_light_transeg.py:148,186 (un-vendored dependency, restated from its published source):
    DiceLoss(include_background=True, softmax over channels, reduction over the spatial dims per (n,c),
    smooth_nr = smooth_dr = 1e-5, mean over (n,c)) + nn.CrossEntropyLoss (mean over voxels); label [B,1,...] holds
    class indices."""
    n_cls = logits.shape[1]
    idx = label[:, 0].long()
    onehot = F.one_hot(idx, n_cls).permute(0, 4, 1, 2, 3).to(logits.dtype)
    p = torch.softmax(logits, dim=1)
    dims = (2, 3, 4)
    inter = (onehot * p).sum(dims)
    denom = onehot.sum(dims) + p.sum(dims)
    dice = (1.0 - (2.0 * inter + smooth) / (denom + smooth)).mean()
    return dice + F.cross_entropy(logits, idx)


def oar_transeg_train_step(sd: SD, x, label, lr=1e-4, weight_decay=1e-5, betas=(0.9, 0.999), eps=1e-8, probe=None, **kw):
    """Transeg.training_step + configure_optimizers (OARSegmentation/train_light_transeg.py:184-198): train-mode
    forward (BatchNorm3d batch statistics in conv_block_7), DiceCELoss, autograd, torch.optim.AdamW(1e-4, wd 1e-5).
    Returns (loss, {name: grad}, {name: updated parameter or running statistic}, logits)."""
    global BN_TRAIN
    is_buffer = lambda k: k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked")
    leaf = {k: v.detach().clone() for k, v in sd.items()}
    train_keys = [k for k in leaf if not is_buffer(k) and leaf[k].is_floating_point()]
    for k in train_keys:
        leaf[k].requires_grad_(True)
    BN_TRAIN = {}
    try:
        logits = oar_transeg_forward(leaf, x, **kw)
        loss = dice_ce_loss(logits, label)
        ((logits * probe).sum() if probe is not None else loss).backward()
        new_stats = dict(BN_TRAIN)
    finally:
        BN_TRAIN = None
    grads = {k: (leaf[k].grad if leaf[k].grad is not None else torch.zeros_like(leaf[k])) for k in train_keys}
    params = [leaf[k] for k in train_keys]                   # grad None (unused parameters): skipped by AdamW
    torch.optim.AdamW(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay).step()
    new = {k: leaf[k].detach() for k in train_keys}
    new.update(new_stats)
    return loss.detach(), grads, new, logits.detach()


def sample_idx(numel, k=64):
    """k evenly spaced flat indices (integer arithmetic; used by the training fixtures)."""
    k = min(k, numel)
    return torch.arange(k, dtype=torch.long) * ((numel - 1) // max(k - 1, 1))


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
