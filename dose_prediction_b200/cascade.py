"""Cascaded OAR-TRANSEG -> DOSE-PYFER inference (reference: LinkedNet.test_step,
DosePrediction/Train/train_light_linked_model.py:138-176) as one static launch schedule per GPU, plus
the per-rank volume sharding used for multi-GPU inference (volumes are independent: no collective).
"""
import torch

from .engine import Plan
from .networks import emit_dose_pyfer, emit_oar_transeg


def shard_volumes(num_volumes: int, rank: int, world_size: int):
    """indices of the patient volumes rank `rank` processes: volumes[rank::world_size]."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    return list(range(rank, num_volumes, world_size))


class CascadePlan:
    """seg(ct) -> argmax/one-hot hand-off -> dose(cat(ptv, oars, ct^T)) for a fixed (batch, size)."""

    def __init__(self, seg_model, dose_model, batch, size, device, keep_structures=False, graph=False):
        if seg_model.training or dose_model.training:
            raise RuntimeError("cascade inference needs both networks in eval() mode")
        P = Plan(device)
        dims = (size, size, size)
        self.ct = P.zeros((batch, 1) + dims, torch.float32)
        self.ptv = P.zeros((batch, 1) + dims, torch.float32)
        seg_x = P.new_act(batch, seg_model.in_ch, dims, lo=True)
        P.pack_input(self.ct, seg_x)
        self.logits = emit_oar_transeg(P, seg_model, seg_x)
        a_out, dose_x = P.new_concat(batch, [dose_model.net_A.list_ch[1], dose_model.in_ch], dims, lo=True)
        self.structures = P.zeros((batch, 9) + dims, torch.float32) if keep_structures else None
        P.handoff(self.logits, self.ptv, self.ct, dose_x, self.structures)
        self.out_A, self.outs = emit_dose_pyfer(P, dose_model, dose_x, a_out)
        self.plan = P
        if graph:
            P.capture()

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


def postprocess_dose(prediction, possible_dose_mask):
    """train_light_linked_model.py:171-173: zero outside the mask / negative values, scale to Gy."""
    prediction = prediction.clone()
    prediction[(possible_dose_mask < 1) | (prediction < 0)] = 0
    return 70.0 * prediction
