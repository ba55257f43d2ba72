"""Cascaded OAR-TRANSEG -> DOSE-PYFER inference (reference: LinkedNet.test_step,
DosePrediction/Train/train_light_linked_model.py:138-176) as one static launch schedule per GPU, plus
the per-rank volume sharding used for multi-GPU inference (volumes are independent: no collective).
"""
import torch

from . import engine
from .engine import Plan
from .networks import emit_dose_pyfer, emit_oar_transeg


def shard_volumes(num_volumes: int, rank: int, world_size: int):
    """indices of the patient volumes rank `rank` processes: volumes[rank::world_size]."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    return list(range(rank, num_volumes, world_size))


def sliding_window_starts(size, roi, overlap=0.25):
    """window origins of monai.inferers.sliding_window_inference (dense_patch_slices / _get_scan_interval) in the
    reference's iteration order (first spatial dim slowest)."""
    if size < roi:
        raise ValueError("volumes smaller than the ROI (padding path of sliding_window_inference) are not built")
    if size == roi:
        per_dim = [0]
    else:
        interval = max(int(roi * (1 - overlap)), 1)
        num = -(-(size - roi) // interval) + 1
        per_dim = [min(k * interval, size - roi) for k in range(num)]
    return [(a, b, c) for a in per_dim for b in per_dim for c in per_dim]


class CascadePlan:
    """seg(ct) -> argmax/one-hot hand-off -> dose(cat(ptv, oars, ct^T)) for a fixed (batch, size).

    sw_roi=None runs the seg net directly on the full volume (its img_size must equal `size`); sw_roi=R
    reproduces the reference's sliding-window call (train_light_linked_model.py:152-154: ROI R, overlap 0.25,
    constant blending) with the seg net built for R^3, `sw_batch` windows per predictor pass."""

    def __init__(self, seg_model, dose_model, batch, size, device, keep_structures=False, graph=False, sw_roi=None,
                 sw_batch=8, overlap=0.25):
        if seg_model.training or dose_model.training:
            raise RuntimeError("cascade inference needs both networks in eval() mode")
        P = Plan(device)
        dims = (size, size, size)
        self.ct = P.zeros((batch, 1) + dims, torch.float32)
        self.ptv = P.zeros((batch, 1) + dims, torch.float32)
        if sw_roi is None:
            seg_x = P.new_act(batch, seg_model.in_ch, dims, lo=True)
            P.pack_input(self.ct, seg_x)
            self.logits = emit_oar_transeg(P, seg_model, seg_x, x_planar=self.ct)
        else:
            self.logits = self._emit_sliding_window(P, seg_model, batch, size, sw_roi, sw_batch, overlap)
        a_out, dose_x = P.new_concat(batch, [dose_model.net_A.list_ch[1], dose_model.in_ch], dims, lo=True)
        self.structures = P.zeros((batch, 9) + dims, torch.float32) if keep_structures else None
        P.handoff(self.logits, self.ptv, self.ct, dose_x, self.structures)
        self.out_A, self.outs = emit_dose_pyfer(P, dose_model, dose_x, a_out)
        self.plan = P
        if engine.COMPACT:
            P.compact()
        if graph:
            P.capture()

    def _emit_sliding_window(self, P, seg_model, batch, size, roi, sw_batch, overlap):
        starts = sliding_window_starts(size, roi, overlap)
        windows = [(b,) + st for b in range(batch) for st in starts]          # reference order: volume-major
        rdims = (roi, roi, roi)
        seg_x = P.new_act(sw_batch, seg_model.in_ch, rdims, lo=True)
        first_step, stats0 = len(P.steps), P.stats_used
        win_logits = emit_oar_transeg(P, seg_model, seg_x)                     # emitted once, replayed per pass
        seg_steps, seg_flops, seg_kernels = P.steps[first_step:], P.step_flops[first_step:], P.step_kernels[first_step:]
        del P.steps[first_step:]
        del P.step_flops[first_step:]
        del P.step_kernels[first_step:]
        fam0 = dict(P.flops)
        seg_stats = P.stats[stats0:P.stats_used]
        total = P.zeros((batch, win_logits.shape[1], size, size, size), torch.float32)
        cnt1 = torch.zeros(size)
        for s0 in sorted({st[0] for st in starts}):
            cnt1[s0:s0 + roi] += 1
        count = P.dev((cnt1[:, None, None] * cnt1[None, :, None] * cnt1[None, None, :]).contiguous())
        P.add_zero(total)
        for g in range(0, len(windows), sw_batch):
            chunk = windows[g:g + sw_batch]
            P.crop_pack(self.ct, roi, chunk, seg_x)
            if seg_stats.numel():
                P.add_zero(seg_stats)
            P.steps.extend(seg_steps)
            P.step_flops.extend(seg_flops)
            P.step_kernels.extend(seg_kernels)
            if g > 0:                                  # the emission above already counted one pass
                for name, fl in fam0.items():
                    P.flops[name] = P.flops.get(name, 0.0) + fl
            P.window_add(win_logits, roi, chunk, total)
        P.div_count(total, count)
        self.windows = windows
        return total

    @property
    def dose(self):
        return self.outs[0]

    def run(self):
        self.plan.replay()

    def __call__(self, ct, ptv):
        """ct, ptv: [B,1,S,S,S] fp32 (CUDA or pinned host) -> dose [B,1,S,S,S] (static output buffer)."""
        self.ct.copy_(ct, non_blocking=True)
        self.ptv.copy_(ptv, non_blocking=True)
        self.plan.replay()
        return self.outs[0]


class CascadeStream:
    """Double-buffered host<->device pipeline around a CascadePlan: the H2D copy of batch i+1 and the D2H copy of
    result i-1 run on side streams while batch i computes, so PCIe time disappears behind the kernels.

        stream = CascadeStream(casc)
        for ct, ptv in batches:            # pinned host tensors [B,1,S,S,S]
            out = stream.submit(ct, ptv)   # returns the pinned host dose of the PREVIOUS submit (or None)
        last = stream.flush()

    Buffer lifetimes: submit() returns only after its H2D copies have completed, so the caller may refill ct / ptv
    right away (the usual loader pattern).  The tensor submit() / flush() return is one of two internal pinned
    buffers: it stays valid until the second submit() after the one that returned it — copy it out if it must live
    longer.  A tcgen05 pipeline fault flagged by any kernel of a batch raises when that batch's result is handed out.
    """

    def __init__(self, casc):
        self.casc = casc
        dev = casc.plan.device
        self.dev = dev
        self.h2d, self.d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        shape = tuple(casc.ct.shape)
        self.stage = [(torch.empty(shape, device=dev), torch.empty(shape, device=dev)) for _ in range(2)]
        self.out_dev = [torch.empty(tuple(casc.dose.shape), device=dev) for _ in range(2)]
        self.out_host = [torch.empty(tuple(casc.dose.shape)).pin_memory() for _ in range(2)]
        self.err_host = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in range(2)]
        self.staged = [torch.cuda.Event() for _ in range(2)]
        self.computed = [torch.cuda.Event() for _ in range(2)]
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.i = 0
        self.pending = None

    def submit(self, ct_host, ptv_host):
        k = self.i & 1
        main = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.h2d):
            if self.i >= 2:
                self.h2d.wait_event(self.consumed[k])       # staging buffer k was read two submits ago
            self.stage[k][0].copy_(ct_host, non_blocking=True)
            self.stage[k][1].copy_(ptv_host, non_blocking=True)
            self.staged[k].record(self.h2d)
        main.wait_event(self.staged[k])
        self.casc.ct.copy_(self.stage[k][0], non_blocking=True)      # device-to-device, ~50 us
        self.casc.ptv.copy_(self.stage[k][1], non_blocking=True)
        self.consumed[k].record(main)
        self.casc.plan.replay()
        if self.i >= 2:
            main.wait_event(self.copied[k])                 # result buffer k must have left the device
        self.out_dev[k].copy_(self.casc.dose, non_blocking=True)
        self.computed[k].record(main)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(self.computed[k])
            self.out_host[k].copy_(self.out_dev[k], non_blocking=True)
            self.err_host[k].copy_(self.casc.plan.err, non_blocking=True)
            self.copied[k].record(self.d2h)
        prev, self.pending = self.pending, k
        self.i += 1
        self.staged[k].synchronize()          # the caller's pinned inputs have been read: they may be refilled now
        if prev is None:
            return None
        return self._collect(prev)

    def _collect(self, k):
        self.copied[k].synchronize()
        if int(self.err_host[k][0]) != 0:
            raise RuntimeError("libdose_b200: an in-kernel mbarrier wait timed out (pipeline protocol error)")
        return self.out_host[k]

    def flush(self):
        if self.pending is None:
            return None
        k, self.pending = self.pending, None
        return self._collect(k)


def postprocess_dose(prediction, possible_dose_mask):
    """train_light_linked_model.py:171-173: zero outside the mask / negative values, scale to Gy."""
    prediction = prediction.clone()
    prediction[(possible_dose_mask < 1) | (prediction < 0)] = 0
    return 70.0 * prediction
