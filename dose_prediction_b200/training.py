"""DOSE-PYFER training step on the C ABI.

Mirrors `Pyfer.training_step` + `configure_optimizers` (DosePrediction/Train/train_light_pyfer.py:85-88,122-143,
194-197) with `GenLoss(mode='train', casecade=True, freez=True)` (DosePrediction/Train/loss.py:69-119):

    train-mode forward (net_A frozen; net_B BatchNorm3d layers use batch statistics and update their running
    statistics) -> deep-supervised masked L1 -> backward through net_B -> [grad all-reduce over ranks] -> AdamW.

`TrainPlan` extends the inference `Plan` with a tape: every forward emitter also registers the emitter of its
backward launches; the tape is unrolled in reverse once, so a training step is one static launch list like the
inference plans.  Convolution data gradients run through the same tcgen05 kernels as the forward pass (flipped /
transposed weights), token-side gradients through the tcgen05 GEMM; everything else is in csrc/train.cu.
PyTorch is used for memory, the per-step re-packing of parameters into kernel layouts and the NCCL all-reduce.

Numerics: forward as the inference path of net_B (fp16 operands, fp32 accumulation); gradients are carried as
fp32 between kernels and as fp16 (static loss scale) into the tensor-core dgrad / wgrad operands.
Biases of convolutions that feed a normalisation layer have a mathematically zero gradient; it is written as 0
(the reference's autograd produces rounding noise there).
"""
import ctypes
import math
import os

import torch
import torch.nn as nn

from . import _lib
from .engine import ACT_ID, Act, LiveWeight, Plan, Raw, Tokens, blocks16, ceil_div
from . import networks as nw

ARENA_DOUBLES = 1 << 22
DDP_OVERLAP = os.environ.get("DP_DDP_OVERLAP", "1") != "0"  # all-reduce the decoder/head gradients while the encoder backward runs
WGRAD_TC = os.environ.get("DP_WGRAD_TC", "1") != "0"      # bring-up switch: tcgen05 vs CUDA-core conv weight gradient


class TrainPlan(Plan):
    def __init__(self, device, loss_scale):
        super().__init__(device)
        self.training = True
        self.loss_scale = float(loss_scale)
        self.tape = []
        self.act_grads = {}      # (buffer ptr, cb_off) -> [(fp32 c8 tensor, cb_total, cb_off)]
        self.raw_grad = {}       # raw tensor ptr -> Act holding the fp16 gradient of that pre-norm tensor
        self.tok_grads = {}      # token tensor ptr -> [fp32 [B,T,C] gradients]
        self.planar_grad = {}    # planar output ptr -> fp32 gradient
        self.arena = self.zeros((ARENA_DOUBLES,), torch.float64)
        self.arena_used = 0
        self.grad_of = {}        # id(parameter) -> fp32 view into the flat gradient buffer
        self.finalizers = []

    # ------------------------------------------------------------------ bookkeeping
    def get_raw(self, N, C, dims, with_stats=True):
        if not self.training:                 # frozen sub-network: pooled like the inference plans
            return super().get_raw(N, C, dims, with_stats)
        t = self.zeros((N, blocks16(C)) + tuple(dims) + (8,), torch.float32)    # kept for the backward pass
        return Raw(t, C, self.new_stats(N, C) if with_stats else None)

    def release(self, raw):
        if not self.training:
            super().release(raw)

    def new_graw(self, N, C, dims):
        return self.get_raw(N, C, dims, with_stats=False)

    @staticmethod
    def _key(a):
        return (a.buf.data_ptr(), a.cb_off)

    def add_act_grad(self, a, src):
        self.act_grads.setdefault(self._key(a), []).append(src)

    def arena_alloc(self, n):
        n = (n + 1) // 2 * 2
        if self.arena_used + n > ARENA_DOUBLES:
            raise RuntimeError("gradient accumulation arena exhausted")
        v = self.arena[self.arena_used:self.arena_used + n]
        self.arena_used += n
        return v

    def grad(self, p):
        return self.grad_of[id(p)]

    def small_grad(self, p):
        """fp64 accumulator for a small parameter; converted into the flat gradient at the end of backward."""
        acc = self.arena_alloc(p.numel())
        g = self.grad(p)
        self.finalizers.append((acc, g, p.numel()))
        return acc

    def _gsum(self, srcs, f16=None):
        assert len(srcs) <= 3, "a tensor with more than three gradient contributions"
        ptrs = _lib.ptr_array([t.data_ptr() for t, _, _ in srcs] or [0])
        cbt = _lib.int_array([c for _, c, _ in srcs] or [0])
        cbo = _lib.int_array([o for _, _, o in srcs] or [0])
        self.keep.append((ptrs, cbt, cbo))
        if f16 is None:
            return (len(srcs), ptrs, cbt, cbo, None, 0, 0)
        return (len(srcs), ptrs, cbt, cbo, f16.buf.data_ptr(), f16.cb_total, f16.cb_off)

    def ones(self, n):
        return self.zeros((n,), torch.float32).fill_(1.0)

    # ------------------------------------------------------------------ differentiable emitters
    def t_conv(self, parts, conv, k, dil=1, need_dgrad=True, mode="p1"):
        """nn.Conv3d (stride 1) on fp16 operands (mode p3: the 3-term hi/lo operand split of the inference recipe, for
        net_A) -> Raw (+ instance statistics).  Backward: wgrad + dgrad, both on fp16 operands."""
        a0 = parts[0]
        N, dims = a0.N, a0.dims
        w = conv.weight
        Co, Ci = w.shape[0], w.shape[1]
        raw = self.get_raw(N, Co, dims)
        shift = conv.bias if conv.bias is not None else self.zeros((Co,), torch.float32)
        self.conv_tc(parts, LiveWeight(w), k, dil, mode, self.ones(Co), shift.detach(), False, out_raw=raw)

        def bwd():
            g16 = self.raw_grad.pop(raw.t.data_ptr())
            self.conv_wgrad(parts, g16, w, k, dil)
            if need_dgrad and Ci % 16 == 0:
                graw = self.new_graw(N, Ci, dims)
                self.conv_tc([g16], LiveWeight(w, transpose_flip=True), k, dil, "p1", self.ones(Ci),
                             self.zeros((Ci,), torch.float32), False, out_raw=graw)
                off = 0
                for a in parts:
                    self.add_act_grad(a, (graw.t, graw.cb_total, off // 8))
                    off += a.C
            elif need_dgrad:
                # cat(net_A output, 9 network-input channels): only the leading parts that fill whole 16-channel chunks
                # carry a gradient (the network input needs none) — a dgrad conv over that slice of the weight
                lead, c0 = [], 0
                for a in parts:
                    if a.C % 16:
                        break
                    lead.append(a)
                    c0 += a.C
                assert lead, "data gradient requested for an input without whole 16-channel parts"
                graw = self.new_graw(N, c0, dims)
                self.conv_tc([g16], lambda: w.detach()[:, :c0].flip(2, 3, 4).transpose(0, 1), k, dil, "p1", self.ones(c0),
                             self.zeros((c0,), torch.float32), False, out_raw=graw)
                off = 0
                for a in lead:
                    self.add_act_grad(a, (graw.t, graw.cb_total, off // 8))
                    off += a.C
        self.tape.append(bwd)
        return raw

    def conv_wgrad(self, parts, g16, w, k, dil, out=None):
        """dW of a stride-1 conv into the flat gradient slot of `w`, or (out given: a [Co, Ci, k, k, k] fp32 tensor whose
        shape stands in for the weight's) into `out` — the space-to-depth form of a stride-2 conv."""
        a0 = parts[0]
        D, H, W = a0.dims
        if out is not None:
            w = out
        Co, Ci = w.shape[0], w.shape[1]
        cbs, ci0, nci, base = [], [], [], 0
        for a in parts:
            assert a.buf is a0.buf
            for j in range(ceil_div(a.C, 16)):
                cbs.append(a.cb_off + 2 * j)
                ci0.append(base + 16 * j)
                nci.append(min(16, a.C - 16 * j))
            base += a.C
        assert base == Ci
        arrs = ((ctypes.c_uint8 * len(cbs))(*cbs), _lib.int_array(ci0), _lib.int_array(nci))
        self.keep.append(arrs)
        flops = 2.0 * a0.N * D * H * W * k ** 3 * Ci * Co
        if WGRAD_TC and dil == 1 and k in (3, 7):
            types = len(cbs) * (Co // 16) * (4 if k == 7 else 1)
            row_blocks = a0.N * D * ceil_div(H, 12) * ceil_div(W, 64)
            splits = max(1, min(row_blocks // 2, 148 // types if types <= 148 else 1))
            ws = self.zeros((splits, w.numel()), torch.float32)
            self.count_flops("dp_conv3d_wgrad_tc", flops)
            self.add("dp_conv3d_wgrad_tc", a0.buf.data_ptr(), a0.cb_total, *arrs, len(cbs), g16.buf.data_ptr(), g16.cb_total,
                     g16.cb_off, a0.N, D, H, W, Ci, Co, k, ws.data_ptr(), splits, self.err.data_ptr())
        else:
            blocks = k * k * len(cbs) * (Co // 16)
            rows = a0.N * D * H
            splits = max(1, min(ceil_div(rows, 64), (3 * 148) // blocks))
            ws = self.zeros((splits, w.numel()), torch.float32)
            self.count_flops("dp_conv3d_wgrad", flops)
            self.add("dp_conv3d_wgrad", a0.buf.data_ptr(), a0.cb_total, *arrs, len(cbs), g16.buf.data_ptr(), g16.cb_total,
                     g16.cb_off, a0.N, D, H, W, Ci, Co, k, dil, ws.data_ptr(), splits)
        self.add("dp_splitk_reduce", ws.data_ptr(), splits, 1, w.numel(), None, None, 0,
                 (out if out is not None else self.grad(w)).data_ptr())

    def t_conv_s2(self, s2d_in, src_act, conv, mode="p3"):
        """nn.Conv3d k3 s2 p1 (c3d.py:49-61) as the tap-masked stride-1 conv over the space-to-depth copy `s2d_in` (8*C
        channels at half resolution, written by the producing norm pass) of `src_act` -> Raw.  Backward: the weight
        gradient of the virtual [Co, 8C, 3,3,3] weight on the tensor cores, of which the 27 real taps are gathered; the data
        gradient as a dense conv producing the space-to-depth gradient, scattered back to full resolution."""
        w = conv.weight
        Co, C = w.shape[0], w.shape[1]
        N, dims = s2d_in.N, s2d_in.dims
        vox = dims[0] * dims[1] * dims[2]
        _, class_masks = nw._s2d_weight(w.detach())
        raw = self.get_raw(N, Co, dims)
        shift = conv.bias if conv.bias is not None else self.zeros((Co,), torch.float32)
        fn = nw._S2DMask(C, class_masks, 2.0 * N * vox * 27 * C * Co)
        self.conv_tc([s2d_in], lambda: nw._s2d_weight(w.detach())[0], 3, 1, mode, self.ones(Co), shift.detach(), False,
                     out_raw=raw, tap_mask_fn=fn)

        def bwd():
            g16 = self.raw_grad.pop(raw.t.data_ptr())
            dws = self.zeros((Co, 8 * C, 3, 3, 3), torch.float32)
            self.conv_wgrad([s2d_in], g16, None, 3, 1, out=dws)
            gw = self.grad(w)
            taps = [(cls, td, th, tw, kd, kh, kw) for cls in range(8)
                    for td, kd in nw._S2D_TAPS[(cls >> 2) & 1] for th, kh in nw._S2D_TAPS[(cls >> 1) & 1]
                    for tw, kw in nw._S2D_TAPS[cls & 1]]

            def gather():
                v = dws.view(Co, 8, C, 3, 3, 3)
                for cls, td, th, tw, kd, kh, kw in taps:
                    gw[:, :, kd, kh, kw].copy_(v[:, cls, :, td, th, tw])
            self.add_py(gather)
            # the dense dgrad conv has 8*C output channels; the tcgen05 conv takes at most 256 per launch
            nsl = max(1, (8 * C) // 256)
            per = 8 * C // nsl
            cps = per // C                                     # parity classes per slice
            gss = []
            for j in range(nsl):
                gs = self.new_graw(N, per, dims)
                self.conv_tc([g16], lambda j=j: nw._s2d_weight(w.detach())[0].flip(2, 3, 4).transpose(0, 1)[j * per:(j + 1) * per],
                             3, 1, "p1", self.ones(per), self.zeros((per,), torch.float32), False, out_raw=gs)
                gss.append(gs)
            gr = self.new_graw(N, C, src_act.dims)

            def depth_to_space():
                dst = gr.t.view(N, C // 8, dims[0], 2, dims[1], 2, dims[2], 2, 8)
                for j, gs in enumerate(gss):
                    v = gs.t.view(N, cps, C // 8, dims[0], dims[1], dims[2], 8)
                    for q in range(cps):
                        cls = j * cps + q
                        dst[:, :, :, (cls >> 2) & 1, :, (cls >> 1) & 1, :, cls & 1].copy_(v[:, q])
            self.add_py(depth_to_space)
            self.add_act_grad(src_act, (gr.t, gr.cb_total, 0))
        self.tape.append(bwd)
        return raw

    def t_upsample(self, src, out):
        """F.interpolate(scale_factor=2, mode='trilinear', align_corners=True) (c3d.py:36); backward = its transpose, three
        separable gather passes (dp_lerp2x_bwd)."""
        self.upsample2x(src, out)

        def bwd():
            dy = self.act_grads.pop(self._key(out))
            N, ncb = src.N, blocks16(src.C)
            D, H, W = src.dims
            g = self.sum_sources(dy, N, ncb, 8 * src.vox)
            t1 = self.zeros((N, ncb, D, 2 * H, 2 * W, 8), torch.float32)
            t2 = self.zeros((N, ncb, D, H, 2 * W, 8), torch.float32)
            gr = self.new_graw(N, src.C, src.dims)
            self.add("dp_lerp2x_bwd", g.data_ptr(), N * ncb, D, 4 * H * W, t1.data_ptr())
            self.add("dp_lerp2x_bwd", t1.data_ptr(), N * ncb * D, H, 2 * W, t2.data_ptr())
            self.add("dp_lerp2x_bwd", t2.data_ptr(), N * ncb * D * H, W, 1, gr.t.data_ptr())
            self.add_act_grad(src, (gr.t, gr.cb_total, 0))
        self.tape.append(bwd)

    def sum_sources(self, srcs, N, ncb, vox):
        """the gradient contributions [(fp32 c8 tensor, cb_total, cb_off)] of one ncb-block activation as ONE contiguous
        fp32 tensor [N][ncb][vox][8] (a single full-tensor contribution is used in place)"""
        t0, cbt0, off0 = srcs[0]
        if len(srcs) == 1 and off0 == 0 and cbt0 == ncb:
            return t0
        views = [t.view(N, cbt, -1)[:, off:off + ncb] for t, cbt, off in srcs]
        dst = self.zeros((N, ncb, vox * 8), torch.float32)

        def run():
            torch.add(views[0], views[1], out=dst) if len(views) > 1 else dst.copy_(views[0])
            for v in views[2:]:
                dst.add_(v)
        self.add_py(run)
        return dst

    def t_norm(self, src, out, act=None, res=None, act_after_res=None, bn=None, stats=None, stats_out=None, identity=False,
               affine=None, s2d=None):
        """InstanceNorm3d (affine: an nn.InstanceNorm3d(affine=True) whose weight / bias train, c3d.py:17) or train-mode
        BatchNorm3d `bn` + activation (+ residual) -> out Act (+ its space-to-depth copy s2d).
        identity=True: no normalisation (the block output is the raw conv output, OldModels conv_3_1)."""
        if isinstance(src, Raw):
            N, C = src.t.shape[0], src.C
            vox = src.t.shape[2] * src.t.shape[3] * src.t.shape[4]
            dims = tuple(src.t.shape[2:5])
            st = None if identity else src.stats
        else:
            N, C, vox, dims, st = src.N, src.C, src.vox, src.dims, stats
        gamma = beta = None
        if bn is not None:
            self.add("dp_batch_combine", st.data_ptr(), N, C, 2, vox, bn.running_mean.data_ptr(), bn.running_var.data_ptr(),
                     float(bn.momentum))
            gamma, beta = bn.weight.detach(), bn.bias.detach()
        elif affine is not None:
            gamma, beta = affine.weight.detach(), affine.bias.detach()
        self.norm_act(src, out, stats=st, gamma=gamma, beta=beta, act=act, res=res, act_after_res=act_after_res,
                      stats_out=stats_out, identity=identity, s2d=s2d)

        def bwd():
            dy = self.act_grads.pop(self._key(out))
            if len(dy) > 3:                   # net_A's output with freeze=False: head, res-block convs, patch embedding
                dy = [(self.sum_sources(dy, N, blocks16(C), vox), blocks16(C), 0)]
            bsum = self.zeros((N * C * 6,), torch.float64)
            self.add_zero(bsum)
            if isinstance(src, Raw):
                fwd = (src.t.data_ptr(), None, None, src.cb_total, 0)
            else:
                fwd = (None, src.hi_ptr, src.lo_ptr, src.cb_total, src.cb_off)
            eh = el = er = es = None
            ecb = eoff = 0
            if res is not None:
                if isinstance(res, Raw):
                    er, ecb, es = res.t.data_ptr(), res.cb_total, res.stats.data_ptr()
                else:
                    eh, el, ecb, eoff = res.hi_ptr, res.lo_ptr, res.cb_total, res.cb_off
            common = fwd + (st.data_ptr() if st is not None else None,
                            gamma.data_ptr() if gamma is not None else None, beta.data_ptr() if beta is not None else None,
                            ACT_ID[act], eh, el, er, es, ecb, eoff, ACT_ID[act_after_res]) + self._gsum(dy)
            # outputs of the apply pass
            dxh = dxf = drh = drf = None
            dxcb = dxoff = drcb = droff = 0
            if isinstance(src, Raw):
                g16 = self.new_act(N, C, dims)
                self.raw_grad[src.t.data_ptr()] = g16
                dxh, dxcb = g16.buf.data_ptr(), g16.cb_total
            else:
                gr = self.new_graw(N, C, dims)
                self.add_act_grad(src, (gr.t, gr.cb_total, 0))
                dxf, dxcb = gr.t.data_ptr(), gr.cb_total
            if res is not None:
                if isinstance(res, Raw):
                    r16 = self.new_act(N, C, dims)
                    self.raw_grad[res.t.data_ptr()] = r16
                    drh, drcb = r16.buf.data_ptr(), r16.cb_total
                else:
                    gr = self.new_graw(N, C, dims)
                    self.add_act_grad(res, (gr.t, gr.cb_total, 0))
                    drf, drcb = gr.t.data_ptr(), gr.cb_total
            tail = (dxh, dxf, dxcb, dxoff, drh, drf, drcb, droff, N, C, vox)
            self.add("dp_norm_act_bwd", *common, bsum.data_ptr(), 0, *tail)
            if bn is not None:
                self.add("dp_batch_combine", bsum.data_ptr(), N, C, 6, vox, None, None, 0.0)
            self.add("dp_norm_act_bwd", *common, bsum.data_ptr(), 1, *tail)
            if bn is not None or affine is not None:
                mod = bn if bn is not None else affine
                self.add("dp_affine_grad", bsum.data_ptr(), N, C, self.grad(mod.weight).data_ptr(),
                         self.grad(mod.bias).data_ptr(), 1.0)
        self.tape.append(bwd)

    def t_pointwise(self, srcs, conv, need_dgrad=True):
        """nn.Conv3d k=1 over cat(srcs) -> Raw."""
        a0 = srcs[0]
        N, dims = a0.N, a0.dims
        w, b = conv.weight, conv.bias
        Co, Ci = w.shape[0], w.shape[1]
        raw = self.get_raw(N, Co, dims)
        self.pointwise([(a, None, None) for a in srcs], w.detach(), b.detach() if b is not None else None, out_raw=raw)

        def bwd():
            g16 = self.raw_grad.pop(raw.t.data_ptr())
            if need_dgrad:
                graw = self.new_graw(N, Ci, dims)
                wT = self.derived(lambda: w.detach().reshape(Co, Ci).t().contiguous())
                self.pointwise([(g16, None, None)], wT, None, out_raw=graw)
            dw = self.small_grad(w)
            db = self.small_grad(b) if b is not None else None
            off = 0
            for i, a in enumerate(srcs):
                if need_dgrad:
                    self.add_act_grad(a, (graw.t, graw.cb_total, off // 8))
                self.add("dp_small_wgrad", *self._gsum([], g16), Co, a.hi_ptr, a.lo_ptr, a.cb_total, a.cb_off, a.C, None,
                         N, dims[0], dims[1], dims[2], 0, dw.data_ptr() + off * 8, Ci, 1, 0,
                         db.data_ptr() if (db is not None and i == 0) else None)
                off += a.C
        self.tape.append(bwd)
        return raw

    def t_head(self, a, conv):
        """dose head: 1x1x1 conv C -> 1, planar fp32 output."""
        w, b = conv.weight, conv.bias
        y = self.zeros((a.N, w.shape[0]) + a.dims, torch.float32)
        self.pointwise([(a, None, None)], w.detach(), b.detach(), out_planar=y)

        def bwd():
            g = self.planar_grad.pop(y.data_ptr())
            graw = self.new_graw(a.N, a.C, a.dims)
            self.add("dp_head_bwd", g.data_ptr(), a.hi_ptr, a.lo_ptr, a.cb_total, a.cb_off, a.C, w.data_ptr(), a.N, a.vox,
                     graw.t.data_ptr(), graw.cb_total, 0, self.small_grad(w).data_ptr(), self.small_grad(b).data_ptr())
            self.add_act_grad(a, (graw.t, graw.cb_total, 0))
        self.tape.append(bwd)
        return y

    def t_deconv(self, src, w, out):
        """ConvTranspose3d k2 s2 (no bias): src Act or Tokens -> out Act."""
        Ci, Co = w.shape[0], w.shape[1]
        self.deconv2x(src, w, out)

        def bwd():
            dy = self.act_grads.pop(self._key(out))
            wb = self.derived(lambda: w.detach().permute(2, 3, 4, 1, 0).reshape(8, Co, Ci).contiguous())
            dw = self.small_grad(w)
            if isinstance(src, Tokens):
                B, T, C = src.t.shape
                D, H, W = src.grid
                dtok = self.zeros((B, T, C), torch.float32)
                self.add("dp_deconv2x_bwd_data", *self._gsum(dy), Co, wb.data_ptr(), Ci, B, D, H, W, None, 0, 0, dtok.data_ptr())
                self.tok_grads.setdefault(src.t.data_ptr(), []).append(dtok)
                self.add("dp_small_wgrad", *self._gsum(dy), Co, None, None, 0, 0, C, src.t.data_ptr(), B, D, H, W, 1,
                         dw.data_ptr(), 8, Co * 8, 1, None)
            else:
                D, H, W = src.dims
                graw = self.new_graw(src.N, Ci, src.dims)
                self.add("dp_deconv2x_bwd_data", *self._gsum(dy), Co, wb.data_ptr(), Ci, src.N, D, H, W, graw.t.data_ptr(),
                         graw.cb_total, 0, None)
                self.add_act_grad(src, (graw.t, graw.cb_total, 0))
                self.add("dp_small_wgrad", *self._gsum(dy), Co, src.hi_ptr, src.lo_ptr, src.cb_total, src.cb_off, src.C, None,
                         src.N, D, H, W, 1, dw.data_ptr(), 8, Co * 8, 1, None)
        self.tape.append(bwd)

    # ------------------------------------------------------------------ token-side helpers
    def transpose(self, src, R, C, dst, batch=1, src_bs=0, dst_bs=0, ld_src=None, ld_dst=None, scale=1.0):
        self.add("dp_transpose", src.data_ptr(), int(src.dtype == torch.float32), src_bs, ld_src or C, R, C, dst.data_ptr(),
                 dst_bs, ld_dst or R, batch, float(scale))

    def heads(self, src, dst, B, T, heads, hd, ld, col0, merge, scale=1.0):
        self.add("dp_heads", src.data_ptr(), int(src.dtype == torch.float32), dst.data_ptr(),
                 int(dst.dtype == torch.float32), B, T, heads, hd, ld, col0, int(merge), float(scale))

    def colsum(self, a, rows, cols, p):
        self.add("dp_colsum", a.data_ptr(), rows, cols, self.small_grad(p).data_ptr())

    def linear_bwd(self, lin, x16, dy32, M, want_dx=True, dx_f16=False):
        """y = x W^T (+b): returns dx (fp32, or fp16 if dx_f16); writes dW, db.  x16 [M,in] fp16, dy32 [M,out] fp32."""
        w = lin.weight
        out_f, in_f = w.shape
        dy16 = self.zeros((M, out_f), torch.float16)
        self.add("dp_add", dy32.data_ptr(), None, M * out_f, None, dy16.data_ptr())
        dyT = self.zeros((out_f, M), torch.float16)
        self.transpose(dy32, M, out_f, dyT)
        xT = self.zeros((in_f, M), torch.float16)
        self.transpose(x16, M, in_f, xT)
        self.gemm(dyT, xT, out_f, in_f, M, out_f32=self.grad(w))
        if lin.bias is not None:
            self.colsum(dy32, M, out_f, lin.bias)
        if not want_dx:
            return None
        wT = self.derived_f16(w, transpose=True)          # [in, out]
        if dx_f16:
            dx = self.zeros((M, in_f), torch.float16)
            self.gemm(dy16, wT, M, in_f, out_f, out_f16=dx)
        else:
            dx = self.zeros((M, in_f), torch.float32)
            self.gemm(dy16, wT, M, in_f, out_f, out_f32=dx)
        return dx

    def t_vit(self, vit, parts, N, S, taps, dgrad_first=False):
        """monai ViT forward (perceptron patch embedding) with every intermediate kept + its backward.
        dgrad_first: also the data gradient of the patch embedding w.r.t. parts[0] (net_A's output when it trains)."""
        P = self
        hidden, heads, L = vit.hidden_size, vit.num_heads, vit.num_layers
        hd = hidden // heads
        grid = tuple(s // 16 for s in S)
        T = grid[0] * grid[1] * grid[2]
        M = N * T
        Tp = ceil_div(T, 8) * 8
        first = parts[0]
        ncb = sum(ceil_div(a.C, 8) if i == len(parts) - 1 else blocks16(a.C) for i, a in enumerate(parts))
        slots, base = [], 0
        for a in parts:
            assert a.cb_off == first.cb_off + base // 8 and a.buf is first.buf, "ViT input parts must be adjacent"
            slots += [base + c for c in range(a.C)]
            base += blocks16(a.C) * 8
        K = ncb * 4096 * 8
        Cin = len(slots)
        pe = vit.patch_embedding
        assert pe.pos_embed == "perceptron", "training path: perceptron patch embedding (dose_pyfer.py:55-67)"
        lin = pe.patch_embeddings[1]
        slot_idx = torch.tensor(slots, device=P.device)

        def pack_pe():
            w = lin.weight.detach().view(hidden, 16, 16, 16, Cin)
            wfull = torch.zeros((hidden, 16, 16, 16, ncb * 8), device=P.device)
            wfull[..., slot_idx] = w
            return wfull.view(hidden, 16, 16, 16, ncb, 8).permute(0, 4, 1, 2, 3, 5).reshape(hidden, K).half()
        wpe = P.derived(pack_pe)
        A = P.zeros((M, K), torch.float16)
        P.patchify(first, ncb, A)
        x0 = P.zeros((M, hidden), torch.float32)
        pos = pe.position_embeddings.detach().reshape(T, hidden)
        tiles = ceil_div(M, 128) * ceil_div(hidden, 128)
        total_kb = K // 64
        split_k = max(1, min(total_kb, (2 * 148) // tiles))
        split_k = ceil_div(total_kb, ceil_div(total_kb, split_k))
        P.gemm_splitk(A, wpe, M, hidden, K, split_k, x0, bias=lin.bias.detach(), rowvec=pos, row_period=T)
        scores = P.zeros((N * heads, T, T), torch.float32)
        saved = []
        hs = {}
        x = x0
        f16 = lambda *shape: P.zeros(shape, torch.float16)
        for i, blk in enumerate(vit.blocks):
            ln1 = f16(M, hidden)
            P.layernorm(x, blk.norm1.weight.detach(), blk.norm1.bias.detach(), M, hidden, out_f16=ln1)
            q, k, vt = f16(N * heads, T, hd), f16(N * heads, T, hd), f16(N * heads, hd, Tp)
            wqkv = P.derived_f16(blk.attn.qkv.weight)
            P.gemm(ln1, wqkv, M, 3 * hidden, hidden, qkv=(heads, hd, T, q, k, vt, hd ** -0.5))
            P.gemm(q, k, T, T, hd, batch=N * heads, a_batch_rows=T, b_batch_rows=T, c_batch_stride=T * T, ldc=T, out_f32=scores)
            probs = f16(N * heads, T, Tp)
            P.softmax(scores, N * heads * T, T, probs)
            o = f16(M, hidden)
            P.gemm(probs, vt, T, hd, Tp, batch=N * heads, a_batch_rows=T, b_batch_rows=hd, c_batch_stride=T * hidden,
                   c_batch_period=heads, c_batch_stride2=hd, ldc=hidden, out_f16=o)
            xm = P.zeros((M, hidden), torch.float32)
            wo = P.derived_f16(blk.attn.out_proj.weight)
            P.gemm(o, wo, M, hidden, hidden, bias=blk.attn.out_proj.bias.detach(), resid=x, out_f32=xm)
            ln2 = f16(M, hidden)
            P.layernorm(xm, blk.norm2.weight.detach(), blk.norm2.bias.detach(), M, hidden, out_f16=ln2)
            u = P.zeros((M, vit.mlp_dim), torch.float32)
            w1 = P.derived_f16(blk.mlp.linear1.weight)
            P.gemm(ln2, w1, M, vit.mlp_dim, hidden, bias=blk.mlp.linear1.bias.detach(), out_f32=u)
            h = f16(M, vit.mlp_dim)
            P.add("dp_act_fwd", u.data_ptr(), M * vit.mlp_dim, ACT_ID["gelu"], h.data_ptr())
            xo = P.zeros((M, hidden), torch.float32)
            tap = None
            if i in taps:
                tap = f16(N, T, hidden)
                hs[i] = Tokens(tap, grid)
            w2 = P.derived_f16(blk.mlp.linear2.weight)
            P.gemm(h, w2, M, hidden, vit.mlp_dim, bias=blk.mlp.linear2.bias.detach(), resid=xm, out_f32=xo, out_f16=tap)
            saved.append(dict(x_in=x, ln1=ln1, q=q, k=k, vt=vt, probs=probs, o=o, xm=xm, ln2=ln2, u=u, h=h, tap=tap))
            x = xo
        z = f16(N, T, hidden)
        P.layernorm(x, vit.norm.weight.detach(), vit.norm.bias.detach(), M, hidden, out_f16=z)
        x_last = x

        def sum_tok(ts):
            g = ts[0]
            for t in ts[1:]:
                s = P.zeros((M, hidden), torch.float32)
                P.add("dp_add", g.data_ptr(), t.data_ptr(), M * hidden, s.data_ptr(), None)
                g = s
            return g

        def bwd():
            f32 = lambda *shape: P.zeros(shape, torch.float32)
            dz = sum_tok(P.tok_grads.pop(z.data_ptr()))
            dx = f32(M, hidden)
            P.add("dp_layernorm_bwd", x_last.data_ptr(), vit.norm.weight.data_ptr(), dz.data_ptr(), None, M, hidden,
                  dx.data_ptr(), P.small_grad(vit.norm.weight).data_ptr(), P.small_grad(vit.norm.bias).data_ptr())
            BH = N * heads
            for i in reversed(range(L)):
                blk, sv = vit.blocks[i], saved[i]
                if sv["tap"] is not None:
                    dx = sum_tok([dx] + P.tok_grads.pop(sv["tap"].data_ptr()))
                # ---- MLP branch: xo = xm + linear2(gelu(linear1(LN2(xm))))
                dh = P.linear_bwd(blk.mlp.linear2, sv["h"], dx, M)
                du = f32(M, vit.mlp_dim)
                P.add("dp_act_bwd", sv["u"].data_ptr(), dh.data_ptr(), M * vit.mlp_dim, ACT_ID["gelu"], du.data_ptr(), None)
                dln2 = P.linear_bwd(blk.mlp.linear1, sv["ln2"], du, M)
                dxm = f32(M, hidden)
                P.add("dp_layernorm_bwd", sv["xm"].data_ptr(), blk.norm2.weight.data_ptr(), dln2.data_ptr(), dx.data_ptr(), M,
                      hidden, dxm.data_ptr(), P.small_grad(blk.norm2.weight).data_ptr(), P.small_grad(blk.norm2.bias).data_ptr())
                # ---- attention branch: xm = x_in + out_proj(softmax(q k^T) v)
                do16 = P.linear_bwd(blk.attn.out_proj, sv["o"], dxm, M, dx_f16=True)
                doh = P.zeros((BH, T, hd), torch.float16)
                P.heads(do16, doh, N, T, heads, hd, hidden, 0, merge=False)
                v16 = P.zeros((BH, T, hd), torch.float16)
                P.transpose(sv["vt"], hd, T, v16, batch=BH, src_bs=hd * Tp, dst_bs=T * hd, ld_src=Tp, ld_dst=hd)
                dP = f32(BH, T, T)
                P.gemm(doh, v16, T, T, hd, batch=BH, a_batch_rows=T, b_batch_rows=T, c_batch_stride=T * T, ldc=T, out_f32=dP)
                dS = P.zeros((BH, T, Tp), torch.float16)
                P.add("dp_softmax_bwd", sv["probs"].data_ptr(), Tp, dP.data_ptr(), T, BH * T, T, dS.data_ptr(), Tp)
                kT = P.zeros((BH, hd, Tp), torch.float16)
                P.transpose(sv["k"], T, hd, kT, batch=BH, src_bs=T * hd, dst_bs=hd * Tp, ld_src=hd, ld_dst=Tp)
                qT = P.zeros((BH, hd, Tp), torch.float16)
                P.transpose(sv["q"], T, hd, qT, batch=BH, src_bs=T * hd, dst_bs=hd * Tp, ld_src=hd, ld_dst=Tp)
                dST = P.zeros((BH, T, Tp), torch.float16)
                P.transpose(dS, T, T, dST, batch=BH, src_bs=T * Tp, dst_bs=T * Tp, ld_src=Tp, ld_dst=Tp)
                PT = P.zeros((BH, T, Tp), torch.float16)
                P.transpose(sv["probs"], T, T, PT, batch=BH, src_bs=T * Tp, dst_bs=T * Tp, ld_src=Tp, ld_dst=Tp)
                dohT = P.zeros((BH, hd, Tp), torch.float16)
                P.transpose(doh, T, hd, dohT, batch=BH, src_bs=T * hd, dst_bs=hd * Tp, ld_src=hd, ld_dst=Tp)
                dq, dk, dv = f32(BH, T, hd), f32(BH, T, hd), f32(BH, T, hd)
                bat = dict(batch=BH, a_batch_rows=T, b_batch_rows=hd, c_batch_stride=T * hd, ldc=hd)
                P.gemm(dS, kT, T, hd, Tp, out_f32=dq, **bat)
                P.gemm(dST, qT, T, hd, Tp, out_f32=dk, **bat)
                P.gemm(PT, dohT, T, hd, Tp, out_f32=dv, **bat)
                dqkv = f32(M, 3 * hidden)
                P.heads(dq, dqkv, N, T, heads, hd, 3 * hidden, 0, merge=True, scale=hd ** -0.5)
                P.heads(dk, dqkv, N, T, heads, hd, 3 * hidden, hidden, merge=True)
                P.heads(dv, dqkv, N, T, heads, hd, 3 * hidden, 2 * hidden, merge=True)
                dln1 = P.linear_bwd(blk.attn.qkv, sv["ln1"], dqkv, M)
                dxi = f32(M, hidden)
                P.add("dp_layernorm_bwd", sv["x_in"].data_ptr(), blk.norm1.weight.data_ptr(), dln1.data_ptr(), dxm.data_ptr(),
                      M, hidden, dxi.data_ptr(), P.small_grad(blk.norm1.weight).data_ptr(), P.small_grad(blk.norm1.bias).data_ptr())
                dx = dxi
            # ---- patch embedding: x0 = A Wpe^T + b + pos
            P.colsum(dx, M, hidden, lin.bias)
            P.add("dp_colsum", dx.data_ptr(), N, T * hidden, P.small_grad(pe.position_embeddings).data_ptr())
            dxT = P.zeros((hidden, M), torch.float16)
            P.transpose(dx, M, hidden, dxT)
            AT = P.zeros((K, M), torch.float16)
            P.transpose(A, M, K, AT)
            dwp = f32(hidden, K)
            P.gemm(dxT, AT, hidden, K, M, out_f32=dwp)
            gw = P.grad(lin.weight)

            def unpack_pe():
                g = dwp.view(hidden, ncb, 16, 16, 16, 8).permute(0, 2, 3, 4, 1, 5).reshape(hidden, 16, 16, 16, ncb * 8)
                gw.view(hidden, 16, 16, 16, Cin).copy_(g[..., slot_idx])
            P.add_py(unpack_pe)
            if dgrad_first:
                # dA = dx Wpe restricted to the K columns of parts[0] (its channel blocks lead the (block, p1, p2, p3, e)
                # flatten order), un-patchified into a c8 fp32 gradient
                a0 = parts[0]
                nb0 = blocks16(a0.C)
                K0 = nb0 * 4096 * 8
                wT = P.derived(lambda: pack_pe()[:, :K0].t().contiguous())          # [K0, hidden] fp16
                dx16 = P.zeros((M, hidden), torch.float16)
                P.add("dp_add", dx.data_ptr(), None, M * hidden, None, dx16.data_ptr())
                dA = f32(M, K0)
                P.gemm(dx16, wT, M, K0, hidden, out_f32=dA)
                gr = P.new_graw(N, a0.C, tuple(S))

                def unpatchify():
                    v = dA.view(N, grid[0], grid[1], grid[2], nb0, 16, 16, 16, 8).permute(0, 4, 1, 5, 2, 6, 3, 7, 8)
                    gr.t.view(N, nb0, grid[0], 16, grid[1], 16, grid[2], 16, 8).copy_(v)
                P.add_py(unpatchify)
                P.add_act_grad(a0, (gr.t, gr.cb_total, 0))
        P.tape.append(bwd)
        return Tokens(z, grid), hs


# --------------------------------------------------------------------------- train-mode network emitters
def _t_res_block(P, blk, parts, out, need_dgrad=True):
    """monai UnetResBlock.forward (k3 s1, InstanceNorm, LeakyReLU 0.01) with backward."""
    N, dims = parts[0].N, parts[0].dims
    Co = blk.conv1.conv.weight.shape[0]
    raw1 = P.t_conv(parts, blk.conv1.conv, 3, need_dgrad=need_dgrad)
    a1 = P.new_act(N, Co, dims)
    P.t_norm(raw1, a1, act="lrelu")
    raw2 = P.t_conv([a1], blk.conv2.conv, 3)
    if blk.downsample:
        raw3 = P.t_pointwise(parts, blk.conv3.conv, need_dgrad=need_dgrad)
        P.t_norm(raw2, out, res=raw3, act_after_res="lrelu")
    else:
        assert len(parts) == 1
        P.t_norm(raw2, out, res=parts[0], act_after_res="lrelu")


def _t_pr_up(P, blk, tokens, out):
    N = tokens.t.shape[0]
    Co = blk.transp_conv_init.conv.weight.shape[1]
    dims = tuple(2 * g for g in tokens.grid)
    n_layers = len(blk.blocks)
    h = out if n_layers == 0 else P.new_act(N, Co, dims)
    P.t_deconv(tokens, blk.transp_conv_init.conv.weight, h)
    for i, seq in enumerate(blk.blocks):
        dims = tuple(2 * d for d in dims)
        u = P.new_act(N, Co, dims)
        P.t_deconv(h, seq[0].conv.weight, u)
        dst = out if i == n_layers - 1 else P.new_act(N, Co, dims)
        _t_res_block(P, seq[1], [u], dst)
        h = dst


def _t_conv_3_1(P, blk, parts, out):
    """blocks_MDUNet.py:132-157 in train mode: the BatchNorm3d layers of conv_block_7 use batch statistics."""
    N, dims = parts[0].N, parts[0].dims
    act = blk.act
    c3, c7 = blk.conv_3[0].conv, blk.conv_7[0].conv
    C = c3[0].weight.shape[0]
    z3, z7 = P.new_concat(N, [C, C], dims)
    raw = P.t_conv(parts, c3[0], 3)
    a = P.new_act(N, C, dims)
    P.t_norm(raw, a, act="relu")
    raw = P.t_conv([a], c3[3], 3)
    y3, st3 = P.new_act(N, C, dims), P.new_stats(N, C)
    P.t_norm(raw, y3, act="relu", stats_out=st3)
    P.t_norm(y3, z3, act=act, stats=st3)
    raw = P.t_conv(parts, c7[0], 7)
    a7 = P.new_act(N, C, dims)
    P.t_norm(raw, a7, act="relu", bn=c7[1])
    raw = P.t_conv([a7], c7[3], 7)
    y7, st7 = P.new_act(N, C, dims), P.new_stats(N, C)
    P.t_norm(raw, y7, act="relu", bn=c7[4], stats_out=st7)
    P.t_norm(y7, z7, act=act, stats=st7)
    raw = P.t_pointwise([z3, z7], blk.conv[0])
    P.t_norm(raw, out, act=act)


def _t_conv_3_1_old(P, blk, parts, out):
    """OARSegmentation/OldModels/Nets/blocks_MDUNet.py:132-147 in train mode: both branches conv -> BN -> ReLU twice
    (batch statistics), then a bare 1^3 conv."""
    N, dims = parts[0].N, parts[0].dims
    C = blk.conv.weight.shape[0]
    ys = P.new_concat(N, [C, C], dims)
    for (br, k), y in zip(((blk.conv_3, 3), (blk.conv_7, 7)), ys):
        c = br.conv
        raw = P.t_conv(parts, c[0], k)
        a = P.new_act(N, C, dims)
        P.t_norm(raw, a, act="relu", bn=c[1])
        raw = P.t_conv([a], c[3], k)
        P.t_norm(raw, y, act="relu", bn=c[4])
    raw = P.t_pointwise(ys, blk.conv)
    P.t_norm(raw, out, identity=True)


def _t_dual_dilated(P, blk, parts, out):
    """blocks_MDUNet.py:194-215 (multiS_conv=False) in train mode: dilation 1 / 2 / 3 branches (conv -> IN -> act twice),
    cat, 1^3 conv, IN, act.  Dilated weight gradients run through the CUDA-core dp_conv3d_wgrad."""
    N, dims = parts[0].N, parts[0].dims
    act = blk.act
    C = blk.conv_3.conv[0].weight.shape[0]
    zs = P.new_concat(N, [C, C, C], dims)
    for br, z in zip((blk.conv_3, blk.conv_5, blk.conv_7), zs):
        c = br.conv
        raw = P.t_conv(parts, c[0], 3, dil=br.dil)
        a = P.new_act(N, C, dims)
        P.t_norm(raw, a, act=act)
        raw = P.t_conv([a], c[3], 3, dil=br.dil)
        P.t_norm(raw, z, act=act)
    raw = P.t_pointwise(zs, blk.conv[0])
    P.t_norm(raw, out, act=act)


def _t_basic_block(P, blk, parts, out):
    """monai 0.7.0 UnetBasicBlock.forward (the conv block of UnetrUpBlock, mode_multi_dec=False) in train mode:
    conv3 -> IN -> LeakyReLU(0.01) twice."""
    N, dims = parts[0].N, parts[0].dims
    Co = blk.conv1.conv.weight.shape[0]
    raw = P.t_conv(parts, blk.conv1.conv, 3)
    a = P.new_act(N, Co, dims)
    P.t_norm(raw, a, act="lrelu")
    raw = P.t_conv([a], blk.conv2.conv, 3)
    P.t_norm(raw, out, act="lrelu")


def _t_unetr(P, vit, enc_blocks, dec_blocks, parts, taps, input_grad=False):
    """UNETR-shaped body shared by MainSubsetModel.forward (dose_pyfer.py:311-319) and oar_transeg Model.forward
    (oar_transeg.py:171-185) in train mode; returns the decoder outputs [full res, /2, /4, /8].
    input_grad: the network input carries a gradient (parts[0] = net_A's output with freeze=False)."""
    N, dims = parts[0].N, parts[0].dims
    fs = enc_blocks[0].layer.conv1.conv.weight.shape[0]
    z, hs = P.t_vit(vit, parts, N, dims, taps, dgrad_first=input_grad)
    sizes = [dims, tuple(d // 2 for d in dims), tuple(d // 4 for d in dims), tuple(d // 8 for d in dims)]
    cats = [P.new_concat(N, [fs << l, fs << l], sizes[l]) for l in range(4)]
    _t_res_block(P, enc_blocks[0].layer, parts, cats[0][1], need_dgrad=input_grad)
    _t_pr_up(P, enc_blocks[1], hs[taps[0]], cats[1][1])
    _t_pr_up(P, enc_blocks[2], hs[taps[1]], cats[2][1])
    _t_pr_up(P, enc_blocks[3], hs[taps[2]], cats[3][1])
    P.tape.append(("mark", "decoders"))          # everything recorded after this point runs its backward before it
    decs, inp = [], z
    for lvl, blk in zip((3, 2, 1, 0), dec_blocks):
        out = P.new_act(N, fs << lvl, sizes[lvl])
        P.t_deconv(inp, blk.transp_conv.conv.weight, cats[lvl][0])
        cov = getattr(getattr(blk, "conv_block", None), "cov_", None)
        if isinstance(blk, nw.UnetrUpBlock):                   # mode_multi_dec=False: monai UnetrUpBlock
            _t_basic_block(P, blk.conv_block, cats[lvl], out)
        elif isinstance(cov, nw.conv_3_1_old):
            _t_conv_3_1_old(P, cov, cats[lvl], out)
        elif isinstance(cov, nw.DualDilatedBlock):             # multiS_conv=False
            _t_dual_dilated(P, cov, cats[lvl], out)
        elif isinstance(cov, nw.conv_3_1):
            _t_conv_3_1(P, cov, cats[lvl], out)
        else:
            raise RuntimeError(f"training path: unknown decoder block {type(blk).__name__}")
        decs.append(out)
        inp = out
    return decs[::-1]


def _t_base_unet(P, net, x_act, out_act):
    """c3d.py:118-149 BaseUNet in train mode (freeze=False): Encoder (:41-72, stride-2 SingleConvs on the space-to-depth
    copy), Decoder (:75-115, trilinear UpConvs), InstanceNorm3d(affine=True) + ReLU after every conv; forward with the
    3-term operand split of the inference recipe, backward on fp16 operands.  The final tensor lands in out_act."""
    N, dims0 = x_act.N, x_act.dims
    ch = net.list_ch
    if any(d % 16 for d in dims0) or any(c % 16 for c in ch[1:]):
        raise RuntimeError("training net_A (freeze=False): sizes must be multiples of 16 (space-to-depth stride-2 convs)")
    dims = [tuple(d >> s for d in dims0) for s in range(5)]
    cat = {s: P.new_concat(N, [ch[s], ch[s]], dims[s - 1], lo=True) for s in (1, 2, 3, 4)}

    def conv_norm(parts, single, out, need_dgrad=True, s2d_out=None):
        raw = P.t_conv(parts, single[0], 3, need_dgrad=need_dgrad, mode="p3")
        P.t_norm(raw, out, act="relu", affine=single[1], s2d=s2d_out)
    h, s2d = x_act, None
    for s in range(1, 6):
        enc = getattr(net.encoder, f"encoder_{s}")
        a = P.new_act(N, ch[s], dims[s - 1], lo=True)
        if s == 1:
            conv_norm([h], enc[0].single_conv, a, need_dgrad=False)          # the network input needs no gradient
        else:
            raw = P.t_conv_s2(s2d, h, enc[0].single_conv[0])
            P.t_norm(raw, a, act="relu", affine=enc[0].single_conv[1])
        dst = cat[s][1] if s <= 4 else P.new_act(N, ch[s], dims[s - 1], lo=True)
        s2d = P.new_act(N, 8 * ch[s], dims[s], lo=True) if s <= 4 else None
        conv_norm([a], enc[1].single_conv, dst, s2d_out=s2d)
        h = dst
    for s in (4, 3, 2, 1):
        up = P.new_act(N, ch[s + 1], dims[s - 1], lo=True)
        P.t_upsample(h, up)
        conv_norm([up], getattr(net.decoder, f"upconv_{s}").conv, cat[s][0])
        dec = getattr(net.decoder, f"decoder_conv_{s}")
        last = (s == 1)
        d0 = out_act if last else P.new_act(N, ch[s], dims[s - 1], lo=True)
        conv_norm(cat[s], dec[0].single_conv, d0)
        h = d0
        if not last:
            d1 = P.new_act(N, ch[s], dims[s - 1], lo=True)
            conv_norm([h], dec[1].single_conv, d1)
            h = d1
    return out_act


def _t_main_subset(P, net, parts, input_grad=False):
    """MainSubsetModel.forward (dose_pyfer.py:311-319) in train mode; returns the four planar dose outputs."""
    enc, dec = net.encoder, net.decoder
    i = enc.num_layers // 4
    decs = _t_unetr(P, enc.vit, (enc.skip1, enc.skip2, enc.skip3, enc.skip4),
                    (dec.decoder4, dec.decoder3, dec.decoder2, dec.decoder1), parts, (i, 2 * i, 3 * i), input_grad=input_grad)
    return [P.t_head(d, conv[0]) for d, conv in zip(decs, net.dose_convertors)]


def allreduce_mean_(flat, group=None):
    """Data-parallel gradient exchange: ONE all-reduce (sum) of the flat gradient buffer, then / world size — what
    Lightning's DDP strategy does for the reference (train_light_pyfer.py trainer setup).  No-op without a group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return flat
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat, group=group)
        flat.div_(world)
    return flat


def unused_parameter_names(model):
    """parameters that live in state_dict() but that no forward touches: the ViT cls_token, MainSubsetModel.out
    (dose_pyfer.py:301-305) and conv3 of monai-0.7 UnetResBlocks whose input and output channels agree."""
    out = set()
    for mn, m in model.named_modules():
        pre = mn + "." if mn else ""
        if isinstance(m, nw._PatchEmbedding):
            out.add(pre + "cls_token")
        if isinstance(m, nw.MainSubsetModel):
            out.update(pre + "out." + n for n, _ in m.out.named_parameters())
        if isinstance(m, nw.UnetResBlock) and not m.downsample:
            out.update(pre + "conv3." + n for n, _ in m.conv3.named_parameters())
    return out


class _Trainer:
    """Shared host side of the training steps: re-homes the trainable parameters into one flat fp32 buffer
    (+ flat gradient / Adam moment buffers), builds the static launch list once, replays it per step."""

    def _setup(self, model, trainable, lr, weight_decay, betas, eps, loss_scale, process_group):
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("dose_prediction_b200 trains on CUDA devices only (no CPU fallback)")
        self.model, self.device = model, dev
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.loss_scale = float(loss_scale)
        self.group = process_group
        self.step_count = 0                     # steps submitted; the bias-correction step lives on the device (opt_state)
        self.opt_state = torch.zeros(3, dtype=torch.int32, device=dev)      # {completed, consecutive skips, total skips}
        # parameters the forward never touches get grad None in the reference, so torch.optim.AdamW / bnb Adam skip them
        # entirely (no decay either): keep them out of the flat optimizer range
        unused = unused_parameter_names(model)
        for n, p in model.named_parameters():
            if getattr(self, "external", False):
                # autograd shim: the caller owns requires_grad; the static backward list covers every used parameter of
                # the trained sub-network, so partial freezing inside it is not built
                if bool(trainable(n)) and not p.requires_grad and n not in unused:
                    raise NotImplementedError(f"train-mode forward: parameter {n} is frozen; only whole-network "
                                              "(net_A / conv_out_A) freezing is built")
                continue
            p.requires_grad_(bool(trainable(n)))
        self.params = [(n, p) for n, p in model.named_parameters() if p.requires_grad and n not in unused]
        self.unused = [(n, p) for n, p in model.named_parameters() if p.requires_grad and n in unused]
        # BatchNorm3d bookkeeping the reference's modules do in train mode: num_batches_tracked += 1 per forward
        self._bn_counters = [m.num_batches_tracked for m in model.modules()
                             if isinstance(m, nn.BatchNorm3d) and m.num_batches_tracked is not None
                             and any(q.requires_grad for q in m.parameters())]
        total = sum((p.numel() + 3) // 4 * 4 for _, p in self.params)
        self.flat_p = torch.zeros(total, device=dev)
        self.flat_g = torch.zeros(total, device=dev)
        self.flat_m = torch.zeros(total, device=dev)
        self.flat_v = torch.zeros(total, device=dev)
        self.found_inf = torch.zeros(1, dtype=torch.int32, device=dev)
        P = TrainPlan(dev, loss_scale)
        off = 0
        self.offsets = {}
        with torch.no_grad():
            for n, p in self.params:
                k = p.numel()
                view = self.flat_p[off:off + k].view(p.shape)
                view.copy_(p.data.to(dev, torch.float32))
                p.data = view
                P.grad_of[id(p)] = self.flat_g[off:off + k].view(p.shape)
                self.offsets[n] = (off, k)
                off += (k + 3) // 4 * 4
        self.total = total
        self.P = P
        self.loss = P.zeros((1,), torch.float32)

    def _finish_emit(self):
        """unroll the tape in reverse; small-parameter gradients are finalised right after the op that produced them,
        so that at the "decoders done" mark the tail of the flat gradient buffer is complete and its all-reduce can
        start while the encoder / ViT backward is still running."""
        P = self.P
        done = 0
        for entry in reversed(P.tape):
            if isinstance(entry, tuple) and entry[0] == "mark":
                if entry[1] == "decoders" and self.tail_off is not None:
                    P.add_py(self._start_tail_allreduce, capturable=False)
                continue
            entry()
            for acc64, g, n in P.finalizers[done:]:
                P.add("dp_grad_finalize", acc64.data_ptr(), g.data_ptr(), n, 1.0)
            done = len(P.finalizers)

    def _set_tail(self, prefixes):
        """flat-buffer offset where the parameters of the modules whose backward runs FIRST begin (they must form the
        tail of the buffer: named_parameters() lists encoders before decoders and heads)."""
        names = [n for n, _ in self.params]
        first = next((i for i, n in enumerate(names) if n.startswith(prefixes)), None)
        self.tail_off = None
        if first is not None and all(n.startswith(prefixes) for n in names[first:]) and DDP_OVERLAP:
            self.tail_off = self.offsets[names[first]][0]
        self._tail_work = None

    skip_allreduce = False      # measurement only (bench.py): time the step without its collective

    def _start_tail_allreduce(self):
        import torch.distributed as dist
        if self.skip_allreduce:
            return
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            self._tail_work = dist.all_reduce(self.flat_g[self.tail_off:], group=self.group, async_op=True)

    # ------------------------------------------------------------------ CUDA-graph replay of the whole step
    graphs = None

    def capture(self):
        """Record the training step into CUDA graphs: {weight re-packing + launch list up to the first host-side step},
        {the rest of the launch list}, {gradient check + AdamW}.  The only host-side steps are the collectives of the
        data-parallel gradient exchange (they stay eager NCCL calls between the graph segments), so on one GPU the step is
        three graph launches instead of ~950 kernel launches.  Call after at least one eager step (warm-up)."""
        P = self.P
        bounds = [i for i, (fn, _, name) in enumerate(P.steps) if fn is None and name == "py_host"]
        ranges, a = [], 0
        for b in bounds:
            ranges.append((a, b))
            a = b + 1
        ranges.append((a, len(P.steps)))
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            segs = []
            for i, (a, b) in enumerate(ranges):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    if i == 0:
                        P._refresh_weights()
                    P._run(a, b, zero_stats=(i == 0))
                    if i == len(ranges) - 1 and self._bn_counters:
                        torch._foreach_add_(self._bn_counters, 1)
                segs.append(g)
            opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(opt):
                self._optimizer_launches()
            self.graphs = (segs, [P.steps[b][1][0] for b in bounds], opt)
        return self

    def _run(self):
        P = self.P
        if self.graphs is not None:
            segs, hosts, _ = self.graphs
            with torch.cuda.device(self.device):
                for i, g in enumerate(segs):
                    g.replay()
                    if i < len(hosts):
                        hosts[i]()
        else:
            P.refresh_weights()
            P.run()
            if self._bn_counters:
                torch._foreach_add_(self._bn_counters, 1)
        if self._tail_work is not None:            # bucket 2 (decoders + heads) has been in flight since the mark
            import torch.distributed as dist
            dist.all_reduce(self.flat_g[:self.tail_off], group=self.group)
            self._tail_work.wait()
            self._tail_work = None
            self.flat_g.div_(dist.get_world_size(self.group))
        elif not self.skip_allreduce:
            allreduce_mean_(self.flat_g, self.group)
        return self.loss

    def _optimizer_step(self):
        self.step_count += 1
        if self.graphs is not None:
            with torch.cuda.device(self.device):
                self.graphs[2].replay()
        else:
            self._optimizer_launches()
        # the parameters (and BatchNorm running statistics) just changed through raw pointers: no tensor version
        # counter moved, so tell every cached inference plan of this model that its packed weights are stale
        nw.invalidate_plans(self.model)
        if self.step_count % self.HEALTH_EVERY == 0:
            self.check_health()

    def _optimizer_launches(self):
        s = torch.cuda.current_stream(self.device).cuda_stream
        lib = self.P.lib
        self.found_inf.zero_()
        _lib.check(lib.dp_grad_check(self.flat_g.data_ptr(), self.total, self.found_inf.data_ptr(), s), "dp_grad_check")
        _lib.check(lib.dp_adamw_dev(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.flat_m.data_ptr(),
                                    self.flat_v.data_ptr(), self.total, self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
                                    1.0 / self.loss_scale, self.found_inf.data_ptr(), self.opt_state.data_ptr(), s),
                   "dp_adamw_dev")

    HEALTH_EVERY = 64
    MAX_CONSECUTIVE_SKIPS = 16

    def check_health(self):
        """host sync: raises if a tcgen05 pipeline fault was flagged or if the static loss scale keeps overflowing
        (every step skipped: training would otherwise stall silently)."""
        self.P.check_device_errors()
        done, consec, total = (int(x) for x in self.opt_state.tolist())
        if consec >= self.MAX_CONSECUTIVE_SKIPS:
            raise RuntimeError(f"{consec} consecutive optimizer steps skipped for non-finite gradients (loss_scale="
                               f"{self.loss_scale:g} overflows fp16): rebuild the trainer with a smaller loss_scale")
        return {"completed_steps": done, "consecutive_skips": consec, "skipped_steps": total}

    def _external_flat_grad(self):
        """autograd shims: the unscaled flat gradient as a FRESH tensor (autograd may keep it as .grad); all-zero when the
        loss-scaled fp16 backward overflowed (the optimizer step is then a no-op, like a GradScaler skip).  No
        all-reduce here: under the shim the caller's DDP wrapper owns the gradient exchange."""
        s = torch.cuda.current_stream(self.device).cuda_stream
        self.found_inf.zero_()
        _lib.check(self.P.lib.dp_grad_check(self.flat_g.data_ptr(), self.total, self.found_inf.data_ptr(), s), "dp_grad_check")
        flat = torch.where(self.found_inf.bool(), torch.zeros((), device=self.device), self.flat_g * (1.0 / self.loss_scale))
        nw.invalidate_plans(self.model)          # BatchNorm running statistics moved in the forward half
        return flat

    def _param_grads(self, flat):
        return tuple(flat[o:o + k].view(p.shape) for (n, p), (o, k) in ((np_, self.offsets[np_[0]]) for np_ in self.params))

    def grads(self):
        """{parameter name: unscaled fp32 gradient} (copies; for tests / inspection)."""
        out = {n: (self.flat_g[o:o + k] / self.loss_scale).view(p.shape).clone()
               for (n, p), (o, k) in ((np_, self.offsets[np_[0]]) for np_ in self.params)}
        out.update({n: torch.zeros_like(p) for n, p in self.unused})      # autograd leaves these None (never used)
        return out


class DoseTrainer(_Trainer):
    """One DOSE-PYFER training step per call: `loss = trainer.step(input_[B,9,S,S,S], gt[B,2,S,S,S])`.

    model: dose_prediction_b200.networks.Model on a CUDA device (parameters are re-homed into one flat fp32
    buffer; state_dict() keeps working).  freeze=True as in the reference's default (net_A / conv_out_A get no gradient);
    freeze=False (the other value of Pyfer's constructor flag, train_light_pyfer.py:61-88) trains every parameter: the
    backward pass continues through net_A (strided convs, trilinear up-sampling, affine InstanceNorm) and GenLoss gains
    0.5 * L1(net_A's own prediction) (loss.py:114-115)."""

    def __init__(self, model, batch, size, lr=1e-4, weight_decay=1e-4, delta1=10.0, delta2=8.0, freeze=True,
                 betas=(0.9, 0.999), eps=1e-8, loss_scale=4096.0, process_group=None, probe=None, external_grads=False):
        """probe (tests only): list of four tensors R_i shaped like the dose outputs; the backward pass then starts
        from dL/dpred_i = R_i (a linear loss sum <pred_i, R_i>) instead of the GenLoss gradient, whose sign()
        makes gradient parity ill-conditioned."""
        self.freeze = bool(freeze)
        self.batch, self.size = batch, size
        self.delta1, self.delta2, self.probe = delta1, delta2, probe
        # external_grads (the autograd shim, autograd_forward below): the loss lives in the caller's torch code; the
        # backward half of the launch list starts from dL/dpred_i written into self.up_grads by the autograd engine
        self.external = bool(external_grads)
        self.up_grads = []
        self._setup(model, lambda n: not self.freeze or not (n.startswith("net_A") or n.startswith("conv_out_A")), lr,
                    weight_decay, betas, eps, loss_scale, process_group)
        self._set_tail(("net_B.decoder.", "net_B.dose_convertors.", "net_B.out."))
        self._emit()

    def _emit(self):
        P, m = self.P, self.model
        N, S = self.batch, self.size
        dims = (S, S, S)
        P.x_in = P.zeros((N, m.in_ch) + dims, torch.float32)
        P.gt = P.zeros((N, 2) + dims, torch.float32)
        a_out, x_act = P.new_concat(N, [m.net_A.list_ch[1], m.in_ch], dims, lo=True)
        P.add_zero(self.flat_g)
        P.add_zero(P.arena)
        P.pack_input(P.x_in, x_act)
        if self.freeze:
            P.training = False      # frozen net_A: inference emitter, weights packed once (InstanceNorm: no running stats)
            nw._emit_base_unet(P, m.net_A, x_act, a_out)
            self.out_A = P.zeros((N, m.out_ch) + dims, torch.float32)
            P.pointwise([(a_out, None, None)], m.conv_out_A.weight, m.conv_out_A.bias, out_planar=self.out_A)
            P.training = True
        else:
            _t_base_unet(P, m.net_A, x_act, a_out)
            self.out_A = P.t_head(a_out, m.conv_out_A)
        outs = _t_main_subset(P, m.net_B, [a_out, x_act], input_grad=not self.freeze)
        self.outs = outs
        self.fwd_end = len(P.steps)              # steps[:fwd_end] = the forward pass, steps[fwd_end:] = loss + backward
        # ---- GenLoss forward
        acc = P.zeros((2 * len(outs) + 2,), torch.float64)
        acc_a = None if self.freeze else acc[2 * len(outs):]          # net_A's own L1 term (freeze=False)
        sizes = [o.shape[2] for o in outs]
        if not self.external:
            P.add_zero(acc)
            for i, o in enumerate(outs):
                P.add("dp_masked_l1", o.data_ptr(), P.gt.data_ptr(), N, S, sizes[i], acc[2 * i:].data_ptr(), 0, 0.0, None)
            if acc_a is not None:
                P.add("dp_masked_l1", self.out_A.data_ptr(), P.gt.data_ptr(), N, S, S, acc_a.data_ptr(), 0, 0.0, None)
            P.add("dp_genloss_finalize", acc.data_ptr(), len(outs), float(self.delta1), float(self.delta2),
                  acc_a.data_ptr() if acc_a is not None else None, 0.5, self.loss.data_ptr())
        # ---- backward
        heads = list(outs) + ([] if self.freeze else [self.out_A])     # freeze=False: net_A's prediction is a fifth head
        for i, o in enumerate(heads):
            g = P.zeros(tuple(o.shape), torch.float32)
            if i == len(outs):
                coef, sz, a_i = self.loss_scale * 0.5, S, acc_a
            else:
                coef, sz, a_i = self.loss_scale * (self.delta1 if i == 0 else self.delta2 / (len(outs) - 1)), sizes[i], acc[2 * i:]
            if self.external:
                self.up_grads.append(g)
            elif self.probe is not None and i < len(self.probe):
                r = (self.probe[i].to(self.device, torch.float32) * self.loss_scale).contiguous()
                P.add_py(lambda g=g, r=r: g.copy_(r))
            elif self.probe is not None:
                pass                                  # probe without an entry for this head: zero gradient
            else:
                P.add("dp_masked_l1", o.data_ptr(), P.gt.data_ptr(), N, S, sz, a_i.data_ptr(), 1, float(coef), g.data_ptr())
            P.planar_grad[o.data_ptr()] = g
        self._finish_emit()
        P.act_grads.pop(P._key(x_act), None)          # gradients w.r.t. the network input are not needed

    def forward_backward(self, x, gt):
        """forward + loss + backward; gradients (times loss_scale) are left in the flat gradient buffer."""
        self.P.x_in.copy_(x.to(torch.float32), non_blocking=True)
        self.P.gt.copy_(gt.to(torch.float32), non_blocking=True)
        return self._run()

    def step(self, x, gt):
        loss = self.forward_backward(x, gt)
        self._optimizer_step()
        return loss

    def outputs(self):
        return [self.out_A.clone(), [o.clone() for o in self.outs]]

    # ---- the two halves of the launch list, for the autograd shim
    def run_forward(self, x):
        self.P.x_in.copy_(x.to(torch.float32), non_blocking=True)
        self.P.refresh_weights()
        self.P.run_range(0, self.fwd_end, True)
        if self._bn_counters:
            torch._foreach_add_(self._bn_counters, 1)

    def run_backward(self, grads):
        """grads: dL/dpred_i (or None) for the four dose outputs -> flat fp32 gradient of the trainable parameters
        (a fresh buffer, unscaled; all-zero when the loss-scaled fp16 backward overflowed: the step is then a no-op)"""
        for g, up in zip(grads, self.up_grads):
            if g is None:
                up.zero_()
            else:
                torch.mul(g.to(self.device, torch.float32), self.loss_scale, out=up)
        self.P.run_range(self.fwd_end, None, False)
        return self._external_flat_grad()


class _DoseTrainFunction(torch.autograd.Function):
    """train-mode `model(x)` as ONE autograd node: forward = the forward half of the static launch list, backward = the
    backward half, started from whatever dL/dpred the caller's loss (GenLoss, train_light_pyfer.py:131) produced."""

    @staticmethod
    def forward(ctx, tr, x, *params):
        tr.run_forward(x)
        ctx.tr = tr
        out_A = tr.out_A.clone()
        if tr.freeze:
            ctx.mark_non_differentiable(out_A)
        return (out_A,) + tuple(o.clone() for o in tr.outs)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_a, *g_outs):
        tr = ctx.tr
        return (None, None) + tr._param_grads(tr.run_backward(list(g_outs) + ([] if tr.freeze else [g_a])))


class _SegTrainFunction(torch.autograd.Function):
    """train-mode OAR-TRANSEG `model(x)` as one autograd node (Transeg.training_step, train_light_transeg.py:184-198)."""

    @staticmethod
    def forward(ctx, tr, x, *params):
        tr.run_forward(x)
        ctx.tr = tr
        return tr.logits()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        tr = ctx.tr
        return (None, None) + tr._param_grads(tr.run_backward(g))


def autograd_forward_seg(model, x):
    """networks.OARTranseg / TRANSEG .forward in train mode: logits wired into torch autograd."""
    if not (x.is_cuda and x.dim() == 5):
        raise RuntimeError("expected a CUDA tensor [B,C,D,H,W]; dose_prediction_b200 has no CPU fallback")
    if x.shape[1] != model.in_ch:
        raise ValueError(f"expected {model.in_ch} input channels, got {x.shape[1]}")
    flags = tuple(p.requires_grad for p in model.parameters())
    cache = model.__dict__.setdefault("_train_ctx", {})
    key = (tuple(x.shape), x.device.index)
    tr = cache.get(key)
    if tr is None or tr._flags != flags:
        cache.clear()
        tr = SegTrainer(model, x.shape[0], x.shape[2], external_grads=True)
        tr._flags = flags
        cache[key] = tr
    return _SegTrainFunction.apply(tr, x.detach(), *[p for _, p in tr.params])


def autograd_forward(model, x):
    """networks.Model.forward in train mode (Pyfer.training_step, train_light_pyfer.py:122-143: `output = self(input_)`,
    then Lightning's loss.backward() and optimizer.step()): returns [output_A, [dose_S, .., dose_S/8]] wired into torch
    autograd; .grad lands on the module's own parameters, any torch / bitsandbytes optimizer can step them."""
    if not (x.is_cuda and x.dim() == 5):
        raise RuntimeError("expected a CUDA tensor [B,C,D,H,W]; dose_prediction_b200 has no CPU fallback")
    if x.shape[1] != model.in_ch:
        raise ValueError(f"expected {model.in_ch} input channels, got {x.shape[1]}")
    flags = tuple(p.requires_grad for p in model.parameters())
    a_flags = [p.requires_grad for n, p in model.named_parameters() if n.startswith("net_A") or n.startswith("conv_out_A")]
    if any(a_flags) and not all(a_flags):
        raise NotImplementedError("train-mode forward: net_A / conv_out_A must be frozen as a whole (Pyfer(freeze=True), "
                                  "train_light_pyfer.py:85-88) or trained as a whole (freeze=False)")
    freeze = not any(a_flags)
    cache = model.__dict__.setdefault("_train_ctx", {})
    key = (tuple(x.shape), x.device.index)
    tr = cache.get(key)
    if tr is None or tr._flags != flags:
        cache.clear()                              # one resident training plan per module
        tr = DoseTrainer(model, x.shape[0], x.shape[2], external_grads=True, freeze=freeze)
        tr._flags = tuple(p.requires_grad for p in model.parameters())
        cache[key] = tr
    outs = _DoseTrainFunction.apply(tr, x.detach(), *[p for _, p in tr.params])
    return [outs[0], list(outs[1:])]


class SegTrainer(_Trainer):
    """One OAR-TRANSEG training step per call (`Transeg.training_step` + `configure_optimizers`,
    OARSegmentation/train_light_transeg.py:184-198): `loss = trainer.step(ct[B,1,S,S,S], label[B,1,S,S,S])` with
    DiceCELoss(to_onehot_y=True, softmax=True) and AdamW(1e-4, weight_decay 1e-5); every parameter is trained.
    model: dose_prediction_b200.networks.OARTranseg (Models/, mode_model=0) or networks.TRANSEG (OldModels/, mode_model=1)."""

    def __init__(self, model, batch, size, lr=1e-4, weight_decay=1e-5, betas=(0.9, 0.999), eps=1e-8, loss_scale=4096.0,
                 process_group=None, probe=None, external_grads=False):
        self.batch, self.size, self.probe = batch, size, probe
        self.external = bool(external_grads)
        self._setup(model, lambda n: True, lr, weight_decay, betas, eps, loss_scale, process_group)
        self._set_tail(("decoder5.", "decoder4.", "decoder3.", "decoder2.", "out."))
        self._emit()

    def _emit(self):
        P, m = self.P, self.model
        N, S = self.batch, self.size
        dims = (S, S, S)
        vox = S ** 3
        C = m.out_channels
        P.x_in = P.zeros((N, m.in_ch) + dims, torch.float32)
        P.label = P.zeros((N, 1) + dims, torch.float32)
        x_act = P.new_act(N, m.in_ch, dims)
        P.add_zero(self.flat_g)
        P.add_zero(P.arena)
        P.pack_input(P.x_in, x_act)
        decs = _t_unetr(P, m.vit, (m.encoder1, m.encoder2, m.encoder3, m.encoder4),
                        (m.decoder5, m.decoder4, m.decoder3, m.decoder2), [x_act], (3, 6, 9))
        raw = P.t_pointwise([decs[0]], m.out.conv.conv)          # logits, c8 fp32 (C classes in block 0)
        self.logits_raw = raw
        self.fwd_end = len(P.steps)
        acc = P.zeros((N * 24 + 2,), torch.float64)
        if not self.external:
            P.add_zero(acc)
            P.add("dp_dice_ce", raw.t.data_ptr(), raw.cb_total, P.label.data_ptr(), N, C, vox, acc.data_ptr(), 0, 0.0, None, 0)
            P.add("dp_dice_ce_finalize", acc.data_ptr(), N, C, vox, self.loss.data_ptr())
        g16 = P.new_act(N, C, dims)
        if self.external:
            self.up_grad = P.zeros((N, C) + dims, torch.float32)          # loss_scale * dL/dlogits, filled by autograd
            P.pack_input(self.up_grad, g16)
        elif self.probe is not None:
            r = (self.probe.to(self.device, torch.float32) * self.loss_scale).contiguous()
            P.keep.append(r)
            P.pack_input(r, g16)
        else:
            P.add("dp_dice_ce", raw.t.data_ptr(), raw.cb_total, P.label.data_ptr(), N, C, vox, acc.data_ptr(), 1,
                  self.loss_scale, g16.buf.data_ptr(), g16.cb_total)
        P.raw_grad[raw.t.data_ptr()] = g16
        self._finish_emit()

    def forward_backward(self, ct, label):
        self.P.x_in.copy_(ct.to(torch.float32), non_blocking=True)
        self.P.label.copy_(label.to(torch.float32), non_blocking=True)
        return self._run()

    def step(self, ct, label):
        loss = self.forward_backward(ct, label)
        self._optimizer_step()
        return loss

    def run_forward(self, ct):
        self.P.x_in.copy_(ct.to(torch.float32), non_blocking=True)
        self.P.refresh_weights()
        self.P.run_range(0, self.fwd_end, True)
        if self._bn_counters:
            torch._foreach_add_(self._bn_counters, 1)

    def run_backward(self, g):
        torch.mul(g.to(self.device, torch.float32), self.loss_scale, out=self.up_grad)
        self.P.run_range(self.fwd_end, None, False)
        return self._external_flat_grad()

    def logits(self):
        """[B,C,S,S,S] fp32 logits of the last forward (copy)."""
        t, C = self.logits_raw.t, self.model.out_channels
        n, cb, d, h, w, _ = t.shape
        return t.permute(0, 1, 5, 2, 3, 4).reshape(n, cb * 8, d, h, w)[:, :C].clone()
