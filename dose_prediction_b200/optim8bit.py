"""bitsandbytes `Adam8bit`-compatible optimizer state (reference: `bnb.optim.Adam8bit(self.model_.parameters(),
lr, weight_decay)`, DosePrediction/Train/train_light_pyfer.py:194-197 — so the `optimizer_states` of a reference Lightning
checkpoint are in this layout).

The trainers keep fp32 Adam moments on the device; this module converts them to / from what
`bnb.optim.Adam8bit.state_dict()` holds (bitsandbytes 0.40.2, `Optimizer2State.init_state` with optim_bits=8,
block_wise=True, min_8bit_size=4096):

    state_dict = {"state": {i: {...}}, "param_groups": [{"params": [0, 1, ...], "lr", "betas", "eps", "weight_decay", ...}]}
    state[i] for a parameter of >= 4096 elements:
        "step"                     int
        "state1", "state2"         uint8, the parameter's shape      (first / second moment codes)
        "qmap1",  "qmap2"          fp32 [256]                         (signed / unsigned "dynamic" code books)
        "absmax1", "absmax2"       fp32 [ceil(numel / 2048)]          (per-block scales)
    smaller parameters keep fp32 "state1" / "state2" (bnb never quantises them).

bitsandbytes is an un-vendored dependency and absent offline: layout, block size and `create_dynamic_map` are restated from
its published sources; parity against the real package is UNPINNED (nothing here can import it).  What the tests pin:
code-book properties, quantise/dequantise round trips against a numpy restatement (oracle/bnb_ref.py), export -> import ->
identical continuation of training up to the 8-bit rounding.
"""
import torch

from . import _lib

BLOCK = 2048
MIN_8BIT_SIZE = 4096


def create_dynamic_map(signed=True, max_exponent_bits=7, total_bits=8):
    """bitsandbytes.functional.create_dynamic_map: the 'dynamic tree' 8-bit data type — values m * 10^-e with the number
    of fraction items doubling per exponent, plus 0 and 1, sorted ascending (256 entries)."""
    data = []
    non_sign_bits = total_bits - 1            # (bnb writes `total_bits - (1 if signed else 1)`: one bit less in both cases)
    additional_items = 2 ** (non_sign_bits - max_exponent_bits) - 1
    i = 0
    for i in range(max_exponent_bits):
        fraction_items = int(2 ** (i + non_sign_bits - max_exponent_bits) + 1 if signed
                             else 2 ** (i + non_sign_bits - max_exponent_bits + 1) + 1)
        boundaries = torch.linspace(0.1, 1, fraction_items)
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        data += ((10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
        if signed:
            data += (-(10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
    if additional_items > 0:
        boundaries = torch.linspace(0.1, 1, additional_items + 1)
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        data += ((10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
        if signed:
            data += (-(10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
    data.append(0)
    data.append(1.0)
    assert len(data) <= 2 ** total_bits
    data += [0] * (2 ** total_bits - len(data))
    data.sort()
    return torch.tensor(data, dtype=torch.float32)


def quantize_blockwise(x, qmap):
    """x fp32 CUDA (any shape) -> (codes uint8 of x's shape, absmax fp32 [ceil(numel / 2048)])."""
    if not x.is_cuda:
        raise RuntimeError("dose_prediction_b200 quantises on CUDA devices only (no CPU fallback)")
    x = x.contiguous().float()
    n = x.numel()
    codes = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    absmax = torch.empty((n + BLOCK - 1) // BLOCK, dtype=torch.float32, device=x.device)
    q = qmap.to(x.device, torch.float32).contiguous()
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().dp_quantize_blockwise(x.data_ptr(), n, q.data_ptr(), codes.data_ptr(), absmax.data_ptr(),
                                                    torch.cuda.current_stream(x.device).cuda_stream), "dp_quantize_blockwise")
    return codes, absmax


def dequantize_blockwise(codes, absmax, qmap, out=None):
    if not codes.is_cuda:
        raise RuntimeError("dose_prediction_b200 dequantises on CUDA devices only (no CPU fallback)")
    codes = codes.contiguous()
    n = codes.numel()
    out = torch.empty(codes.shape, dtype=torch.float32, device=codes.device) if out is None else out
    q = qmap.to(codes.device, torch.float32).contiguous()
    a = absmax.to(codes.device, torch.float32).contiguous()
    with torch.cuda.device(codes.device):
        _lib.check(_lib.lib().dp_dequantize_blockwise(codes.data_ptr(), a.data_ptr(), q.data_ptr(), n, out.data_ptr(),
                                                      torch.cuda.current_stream(codes.device).cuda_stream), "dp_dequantize_blockwise")
    return out


def export_state(trainer, parameters=None):
    """The trainer's optimizer state as `bnb.optim.Adam8bit.state_dict()` would hold it.  `parameters`: the parameter list the
    reference optimizer was built over (default: `model.parameters()`, as configure_optimizers does); parameters the trainer
    does not train (frozen / never used) get no state entry, exactly like parameters whose grad stayed None under bnb."""
    model = trainer.model
    params = list(model.parameters()) if parameters is None else list(parameters)
    name_of = {id(p): n for n, p in model.named_parameters()}
    qmap1, qmap2 = create_dynamic_map(True), create_dynamic_map(False)
    step = int(trainer.opt_state[0].item())
    state = {}
    for i, p in enumerate(params):
        n = name_of.get(id(p))
        if n not in trainer.offsets:
            continue
        o, k = trainer.offsets[n]
        m, v = trainer.flat_m[o:o + k].view(p.shape), trainer.flat_v[o:o + k].view(p.shape)
        if k < MIN_8BIT_SIZE:
            state[i] = {"step": step, "state1": m.clone(), "state2": v.clone()}
            continue
        c1, a1 = quantize_blockwise(m, qmap1)
        c2, a2 = quantize_blockwise(v, qmap2)
        state[i] = {"step": step, "state1": c1, "qmap1": qmap1.to(p.device), "absmax1": a1,
                    "state2": c2, "qmap2": qmap2.to(p.device), "absmax2": a2}
    group = {"lr": trainer.lr, "betas": tuple(trainer.betas), "eps": trainer.eps, "weight_decay": trainer.wd,
             "params": list(range(len(params)))}
    return {"state": state, "param_groups": [group]}


def import_state(trainer, state_dict, parameters=None):
    """Resume from a `bnb.optim.Adam8bit.state_dict()` (e.g. `ckpt["optimizer_states"][0]` of a reference Lightning
    checkpoint): dequantise the moments into the trainer's fp32 buffers with the code books stored IN the checkpoint, set the
    step count.  fp32 states (small parameters, or a torch.optim.Adam(W) state with exp_avg / exp_avg_sq) are copied."""
    model = trainer.model
    params = list(model.parameters()) if parameters is None else list(parameters)
    name_of = {id(p): n for n, p in model.named_parameters()}
    steps = []
    trainer.flat_m.zero_()
    trainer.flat_v.zero_()
    for i, st in state_dict["state"].items():
        p = params[int(i)]
        n = name_of.get(id(p))
        if n not in trainer.offsets:
            continue
        o, k = trainer.offsets[n]
        m, v = trainer.flat_m[o:o + k].view(p.shape), trainer.flat_v[o:o + k].view(p.shape)
        s1 = st["state1"] if "state1" in st else st["exp_avg"]
        s2 = st["state2"] if "state2" in st else st["exp_avg_sq"]
        if s1.dtype == torch.uint8:
            dequantize_blockwise(s1.to(trainer.device), st["absmax1"], st["qmap1"], out=m)
            dequantize_blockwise(s2.to(trainer.device), st["absmax2"], st["qmap2"], out=v)
        else:
            m.copy_(s1.to(trainer.device, torch.float32))
            v.copy_(s2.to(trainer.device, torch.float32))
        steps.append(int(st["step"]))
    step = max(steps) if steps else 0
    trainer.opt_state.zero_()
    trainer.opt_state[0] = step
    trainer.step_count = step
    g = state_dict.get("param_groups", [{}])[0]
    trainer.lr = g.get("lr", trainer.lr)
    trainer.wd = g.get("weight_decay", trainer.wd)
    trainer.betas = tuple(g.get("betas", trainer.betas))
    trainer.eps = g.get("eps", trainer.eps)
    if getattr(trainer, "graphs", None) is not None:
        trainer.graphs = None              # lr / betas are baked into the captured optimizer launches: re-capture
    return step
