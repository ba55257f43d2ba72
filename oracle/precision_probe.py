"""TEST INFRASTRUCTURE — operand-precision emulation of the CUDA path on CPU.

    python -m oracle.precision_probe [--size 64] [--seeds 2]

Runs oracle.torch_ref in exact fp32 and under RECIPE (the storage / operand precision of every
contraction exactly as dose_prediction_b200 feeds its kernels: fp16, fp16 hi+lo pairs, fp32 SIMT
weights; fp32 accumulation everywhere) and prints the parity metrics north_star asks for:
rel-L2(dose) <= 1e-2, rel-L2(logits) <= 1e-2, argmax agreement >= 99.9 %.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dose_prediction_b200 import synth  # noqa: E402
from oracle import synth_ckpt, torch_ref  # noqa: E402


def recipe_dose(tag, kind):
    """dose_prediction_b200/networks precision plan for DOSE-PYFER (kept in sync by tests)."""
    if tag.startswith("net_A.") or tag.startswith("conv_out_A."):
        if kind in ("conv3s2", "conv1"):
            return "hilo", "f32"          # SIMT kernels: fp32 weights, hi+lo activations
        if kind == "upsample":
            return "hilo", "f32"
        return "hilo", "hilo"             # tensor-core conv, 3-term split
    if kind in ("conv1", "deconv", "store"):
        return "f16", "f32"
    return "f16", "f16"


SEG_DEEP = ("decoder5.", "decoder4.", "encoder4.", "encoder3.")     # the 16^3 / 32^3 levels (at 128^3 input)


def recipe_seg(tag, kind):
    if kind == "conv7" or kind in ("linear", "attn"):
        return "f16", "f16"
    if tag.startswith(SEG_DEEP) and kind in ("conv3", "conv1", "store", "deconv"):
        return "f16", ("f16" if kind == "conv3" else "f32")      # deep levels: plain fp16 operands / storage
    if tag.startswith("decoder3.transp_conv"):                    # reads decoder4's fp16-stored output
        return "f16", "f32"
    if kind == "deconv":
        tokens = "transp_conv_init" in tag or tag.startswith("decoder5.transp_conv")
        return ("f16" if tokens else "hilo"), "f32"
    if kind in ("conv1", "store"):
        return "hilo", "f32"
    return "hilo", "hilo"


def _manifest(name, tokens):
    with open(os.path.join(ROOT, "tests", "golden", f"manifest_{name}.json")) as f:
        man = json.load(f)
    return [(k, ([1, tokens, s[2]] if k.endswith("position_embeddings") else s)) for k, s, *_ in man]


def run(size=64, seeds=2, dose_recipe=recipe_dose, seg_recipe=recipe_seg, verbose=True):
    tokens = (size // 16) ** 3
    out = []
    for sd_seed in range(seeds):
        dsd = synth_ckpt.make_state_dict(_manifest("dose_pyfer", tokens), seed=10 + sd_seed)
        ssd = synth_ckpt.make_state_dict(_manifest("oar_transeg", tokens), seed=20 + sd_seed)
        vol = synth.make_volume(size, seed=1234 + sd_seed)
        with torch.no_grad():
            torch_ref.EMU = None
            logits = torch_ref.oar_transeg_forward(ssd, vol["ct"])
            dose = torch_ref.dose_pyfer_forward(dsd, vol["dose_input"])[1][0]
            torch_ref.EMU = seg_recipe
            logits_q = torch_ref.oar_transeg_forward(ssd, vol["ct"])
            torch_ref.EMU = dose_recipe
            dose_q = torch_ref.dose_pyfer_forward(dsd, vol["dose_input"])[1][0]
            torch_ref.EMU = None
        res = {"seed": sd_seed, "logits_rel_l2": torch_ref.rel_l2(logits_q, logits),
               "argmax_agree": (logits_q.argmax(1) == logits.argmax(1)).float().mean().item(),
               "dose_rel_l2": torch_ref.rel_l2(dose_q, dose)}
        out.append(res)
        if verbose:
            print(res, flush=True)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--seeds", type=int, default=2)
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    run(a.size, a.seeds)
