#!/usr/bin/env python
"""Headline benchmark: 128^3 cascade volumes/sec (OAR-TRANSEG -> hand-off -> DOSE-PYFER inference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--size S] [--impl reference]

One "step" = one cascade pass over a batch of B synthetic OpenKBP-shaped volumes per GPU.  Prints ONE
JSON line (see DESIGN.md "Measurement"):
  value     whole-job volumes/s with inputs resident in HBM (device-timed, CUDA events, max over ranks)
  e2e       same metric through CascadePlan.__call__ with pinned HOST inputs (H2D + D2H inside the timing)
  roofline  the dominant kernel family (tcgen05 implicit-GEMM conv): algorithmic conv FLOPs / summed
            launch time measured live with CUDA events, against the measured bf16 peak
  cpu_baseline  the reference's own nn.Modules (oracle/_ref, kind "reference"; the oracle port oracle/torch_ref.py, kind
            "port", when that copy is absent), fp32, all host threads, on one volume, rank 0, N=1
  parity    the timed batch-8 plan's volume 0 (logits, argmax, dose) against that CPU run; outside north_star's
            tolerance the run FAILS (exit code 3)
  train     BASELINE.json configs[3] measured in the same run: DOSE-PYFER training step, batch 2 per GPU, with the
            NCCL gradient all-reduce when N > 1 (samples/s, ms/step, all-reduce ms, overlap)
`--impl reference` times the CPU implementation alone (reference modules over the monai 0.7.0 restatement).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cascade_volumes_per_sec_128cubed"
UNIT = "volumes/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops": float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1400.0))), "hbm_gbs": float(p["hbm_gbs"]),
                "tflops_burst": float(p.get("bf16_tflops", 0.0)) or None,
                "source": "measured (MEASURED_PEAKS.json, sustained bf16: the launch is timed inside a long step)"}
    return {"tflops": 1400.0, "hbm_gbs": 6650.0, "tflops_burst": None, "source": "fallback (B200_PROFILING.md, sustained)"}


_JSON_OUT = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line, but libraries write to fd 1 too (NCCL prints its version banner
    there at communicator creation).  Keep a private handle on the real stdout and point fd 1 at stderr."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def _build_once(entry, dist, world, local):
    """compile (if stale) on local rank 0 only; the other ranks wait, then just load the library."""
    if world > 1:
        if local == 0:
            entry.build()
        dist.barrier()
    entry.build()


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                c = [x.strip() for x in line.split(",")]
                if len(c) < 8:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2]))
                except ValueError:
                    continue
                for n, v in zip(names, c[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        os.unlink(self.path)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class CpuCascade:
    """The CPU implementation of the path: the reference's own nn.Modules (oracle/_ref or /root/reference, imported
    unmodified over oracle/monai_compat) when available — kind "reference" — else the oracle port — kind "port".
    The argmax / one-hot hand-off (LinkedNet.test_step needs Lightning) is the oracle's restatement in both cases."""

    def __init__(self, seg_sd, dose_sd, size):
        import torch

        from oracle import ref_loader, torch_ref
        self.torch_ref, self.kind = torch_ref, "port"
        self.seg_sd, self.dose_sd = seg_sd, dose_sd
        torch.set_num_threads(os.cpu_count() or 1)
        if ref_loader.available() and os.environ.get("DP_BENCH_PORT", "0") == "0":
            try:
                self.seg = ref_loader.build_seg(size).eval()
                self.dose = ref_loader.build_dose(size).eval()
                self.seg.load_state_dict(seg_sd, strict=True)
                self.dose.load_state_dict(dose_sd, strict=True)
                self.kind = "reference"
            except Exception as e:                      # pragma: no cover - reported in the JSON line
                self.kind, self.why = "port", f"reference modules unusable: {e!r}"
        self.what = ("reference nn.Modules (dose_pyfer.Model, oar_transeg.Model) over oracle/monai_compat"
                     if self.kind == "reference" else "oracle/torch_ref.py")

    def __call__(self, vol, structures=None):
        """-> (logits, structures, dose) for one volume dict; structures given = dose net only on those."""
        import torch
        tr = self.torch_ref
        with torch.no_grad():
            logits = None
            if structures is None:
                logits = self.seg(vol["ct"]) if self.kind == "reference" else tr.oar_transeg_forward(self.seg_sd, vol["ct"])
                structures = tr.handoff(logits, vol["ptv"], vol["ct"])
            dose = (self.dose(structures) if self.kind == "reference" else tr.dose_pyfer_forward(self.dose_sd, structures))[1][0]
        return logits, structures, dose


def cpu_baseline(cpu, size, volumes=3, seed=1234):
    """the CPU implementation timed on the host cores (reported baseline, not the optimisation target)."""
    import torch

    from dose_prediction_b200 import synth
    vol = synth.make_volume(size, seed=seed)
    t0 = time.perf_counter()
    for _ in range(volumes):
        out = cpu(vol)
    dt = time.perf_counter() - t0
    return {"value": volumes / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": cpu.kind,
            "sample": f"{volumes} x {size}^3 cascade volume(s), {cpu.what}, fp32, {dt:.1f} s"}, vol, out


def parity_check(cpu, vol, cpu_out, got_logits, got_dose, ptv, ct):
    """the benchmarked plan's volume 0 against the CPU implementation (north_star tolerances).  The timed plan does not
    write the 9-channel structures tensor out (nothing downstream needs it); the hand-off of the GPU's own logits is
    recomputed here with the oracle's handoff(), which the GPU kernel matches bit for bit (tests/test_kernels_gpu.py)."""
    from oracle import torch_ref
    logits, st, dose = cpu_out
    got_structures = torch_ref.handoff(got_logits, ptv, ct)
    _, _, dose_same = cpu(vol, structures=got_structures)       # dose net alone on the GPU path's own structures
    rep = {"logits_rel_l2": torch_ref.rel_l2(got_logits, logits),
           "argmax_agree": float((got_logits.argmax(1) == logits.argmax(1)).float().mean()),
           "structures_agree": float((got_structures == st).float().mean()),
           "dose_rel_l2": torch_ref.rel_l2(got_dose, dose_same),
           "dose_rel_l2_end_to_end": torch_ref.rel_l2(got_dose, dose),
           "against": cpu.kind, "volume": "batch entry 0 of the timed plan",
           "tolerance": {"logits_rel_l2": 1e-2, "argmax_agree": 0.999, "dose_rel_l2": 1e-2}}
    rep["ok"] = bool(rep["logits_rel_l2"] <= 1e-2 and rep["argmax_agree"] >= 0.999 and rep["dose_rel_l2"] <= 1e-2)
    return rep


def build_models(size, device=None, seg_size=None):
    import torch

    from dose_prediction_b200 import networks
    torch.manual_seed(0)
    dose = networks.Model(9, 1, [-1, 16, 32, 64, 128, 256], feature_size=16, img_size=(size,) * 3, num_layers=8,
                          num_heads=6, act="mish", mode_multi_dec=True, multiS_conv=True).eval()
    seg = networks.OARTranseg(1, 8, (seg_size or size,) * 3, feature_size=16, hidden_size=768, mlp_dim=3072, num_heads=12,
                              pos_embed="perceptron", norm_name="instance", res_block=True, conv_block=True).eval()
    if device is not None:
        dose, seg = dose.to(device), seg.to(device)
    return seg, dose


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch

    from dose_prediction_b200 import synth
    seg, dose = build_models(args.size)
    cpu = CpuCascade(seg.state_dict(), dose.state_dict(), args.size)
    vol = synth.make_volume(args.size, seed=1234)
    for _ in range(args.warmup):
        cpu(vol)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu(vol)
    dt = time.perf_counter() - t0
    val = args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"cascade OAR-TRANSEG->DOSE-PYFER inference, {args.size}^3, 1 volume per step (bounded sample)",
                       "batch_per_gpu": 1},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": cpu.kind,
                             "sample": f"{args.steps} x {args.size}^3 cascade volumes, {cpu.what}, fp32"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def run_train(args):
    """--workload train: BASELINE.json configs[3], the DOSE-PYFER training step (forward + GenLoss + backward through
    net_B + AdamW), batch 2 per GPU, data-parallel with one NCCL all-reduce of the flat gradient buffer."""
    import torch
    import torch.distributed as dist

    import __graft_entry__
    from dose_prediction_b200 import synth
    from dose_prediction_b200.training import DoseTrainer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    B, S = (args.batch if args.batch_given else 2), args.size
    metric, unit = "dose_pyfer_train_samples_per_sec_128cubed", "samples/s"
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import ref_loader, torch_ref
        torch.set_num_threads(os.cpu_count() or 1)
        size = min(S, 64)
        _, dose = build_models(size)
        sd = dose.state_dict()
        vol = synth.make_batch(1, size, seed=1234)
        kind, what = "port", "oracle/torch_ref.py autograd"
        if ref_loader.available() and os.environ.get("DP_BENCH_PORT", "0") == "0":
            # the reference's own module + GenLoss + torch.optim.AdamW (train_light_pyfer.py:85-88,122-143,194-197)
            tm = ref_loader.build_dose(size).train()
            tm.load_state_dict(sd, strict=True)
            for n, p_ in tm.named_parameters():
                if "net_A" in n or "conv_out_A" in n:
                    p_.requires_grad = False
            opt = torch.optim.AdamW([p_ for p_ in tm.parameters() if p_.requires_grad], lr=1e-4, weight_decay=1e-4)
            loss_mod = ref_loader.loss().GenLoss(im_size=size)
            kind, what = "reference", "reference dose_pyfer.Model + loss.GenLoss + torch.optim.AdamW (autograd)"

            def one_step():
                opt.zero_grad(set_to_none=True)
                loss_mod(tm(vol["dose_input"]), vol["gt"], casecade=True, freez=True, delta1=10, delta2=8).backward()
                opt.step()
        else:
            def one_step():
                torch_ref.dose_pyfer_train_step(sd, vol["dose_input"], vol["gt"])
        t0 = time.perf_counter()
        for _ in range(max(1, args.steps)):
            one_step()
        dt = (time.perf_counter() - t0) / max(1, args.steps)
        val = (size / S) ** 3 / dt
        sample = (f"{max(1, args.steps)} training step(s), batch 1, {size}^3, {what}, fp32; value scaled by "
                  f"({size}/{S})^3 to {S}^3-equivalent samples/s")
        _emit({"impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": 0, "ms_per_step": 1e3 * dt, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": f"DOSE-PYFER training step, bounded CPU sample ({sample})"},
                          "cpu_baseline": {"value": val, "unit": unit, "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
                          "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _build_once(__graft_entry__, dist, world, local)
    t = measure_train(B, S, args.steps, args.warmup, dev, world, rank, local, detail=True)
    if rank == 0:
        line = {"metric": metric, "value": t["value"], "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": t["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f16", "data": "synthetic", "config": t["config"], "e2e": t["e2e"],
                "gpu_launches": t["launches_per_step"] * args.steps, "launches_per_step": t["launches_per_step"],
                "roofline": t["roofline"], "roofline_families": t["roofline_families"],
                "kernel_ms_per_step": t["kernel_ms_per_step"], "allreduce": t["allreduce"], "clocks": t["clocks"]}
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_train(B, S, steps, warmup, dev, world, rank, local, detail=False):
    """BASELINE.json configs[3]: DOSE-PYFER training step (train-mode forward, GenLoss, backward through net_B, AdamW),
    batch B per GPU, data-parallel over `world` ranks with one NCCL all-reduce (two buckets) of the flat gradient.
    Returns the measurements as a dict (rank 0 gets clocks); device-timed with CUDA events, max over ranks."""
    import torch
    import torch.distributed as dist

    from dose_prediction_b200 import synth
    from dose_prediction_b200.training import DoseTrainer
    unit = "samples/s"
    _, dose = build_models(S, dev)
    dose.train()
    tr = DoseTrainer(dose, B, S, lr=1e-4, weight_decay=1e-4)
    vols = synth.make_batch(B, S, seed=1234 + rank * B)
    x_h, gt_h = vols["dose_input"].pin_memory(), vols["gt"].pin_memory()
    x_d, gt_d = x_h.to(dev), gt_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(warmup, 3)):
        tr.step(x_d, gt_d)
    torch.cuda.synchronize(dev)
    tr.check_health()
    graphed = os.environ.get("DP_TRAIN_GRAPH", "1") != "0"
    if graphed:
        tr.capture()                     # the step as CUDA graphs (collectives stay eager NCCL calls between the segments)
        tr.step(x_d, gt_d)
        torch.cuda.synchronize(dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(lambda: tr.step(x_d, gt_d), steps)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * steps / (ms / 1e3)
    losses = []

    def e2e_step():                      # host batch in (pinned), scalar loss out, every step
        losses.append(float(tr.step(x_h.to(dev, non_blocking=True), gt_h.to(dev, non_blocking=True))))
    ms_e2e = timed(e2e_step, steps)
    e2e_val = world * B * steps / (ms_e2e / 1e3)
    # ---- the collective: one all-reduce of the whole flat gradient timed alone, and the step without it
    allreduce = {"collective": "none (single GPU)", "ms_alone": 0.0, "exposed_ms_per_step": 0.0, "overlap_fraction": None,
                 "bytes": tr.total * 4}
    if world > 1:
        n_ar = 5
        ms_ar = timed(lambda: dist.all_reduce(tr.flat_g, group=tr.group), n_ar) / n_ar
        tr.skip_allreduce = True
        ms_compute = timed(lambda: tr.step(x_d, gt_d), steps) / steps
        tr.skip_allreduce = False
        exposed = max(0.0, ms / steps - ms_compute)
        allreduce = {"collective": f"NCCL all-reduce (sum) of the flat fp32 gradient, {tr.total} elements, two buckets "
                                   "(decoders+heads overlapped with the encoder backward, then the rest)",
                     "ms_alone": ms_ar, "bus_GBps": 2.0 * (world - 1) / world * tr.total * 4 / (ms_ar / 1e3) / 1e9,
                     "step_ms_without_collective": ms_compute, "exposed_ms_per_step": exposed,
                     "overlap_fraction": max(0.0, 1.0 - exposed / ms_ar) if ms_ar > 0 else None, "bytes": tr.total * 4}
    P = tr.P
    lps = P.kernels_per_step + 3 + len(P.refresh_launches)
    out = {"metric": "dose_pyfer_train_samples_per_sec_128cubed", "value": value, "unit": unit, "ms_per_step": ms / steps,
           "steps": steps, "n_gpus": world, "batch_per_gpu": B, "size": S, "dtype": "f16",
           "e2e": {"value": e2e_val, "unit": unit, "ms_per_step": ms_e2e / steps,
                   "h2d_bytes_per_step": B * 11 * S ** 3 * 4, "d2h_bytes_per_step": 4, "last_loss": losses[-1]},
           "allreduce": allreduce, "launches_per_step": lps, "clocks": clocks,
           "algorithmic_tflop_per_sample": 7.87 * (S / 128.0) ** 3,
           "achieved_tflops_per_gpu": 7.87 * (S / 128.0) ** 3 * B * steps / (ms / 1e3),
           "optimizer": tr.check_health(),
           "config": {"workload": f"DOSE-PYFER training step (train-mode forward, GenLoss, backward through net_B, AdamW), "
                                  f"{S}^3, batch {B} per GPU (BASELINE.json configs[3])", "batch_per_gpu": B, "size": S,
                      "parallelism": f"data-parallel x{world}, one NCCL all-reduce of the flat fp32 gradient buffer "
                                     f"({tr.total} elements)" if world > 1 else "single GPU",
                      "loss_scale": tr.loss_scale, "trainable_parameters": tr.total, "cuda_graph": graphed,
                      "l2": f"no flush: per-step working set {P.bytes_alloc / 2**30:.1f} GiB >> 126 MB L2"}}
    peaks = _peaks()
    out["frac_of_tensor_peak"] = out["achieved_tflops_per_gpu"] / peaks["tflops"]
    if detail:
        fam = P.profile_families()
        total_fam = sum(v["ms"] for v in fam.values()) or 1.0
        rows = [r for r in P.profile_launches() if r[3] > 0]
        top = max(rows, key=lambda r: r[2])
        ach = top[3] / (top[2] / 1e3) / 1e12
        tc_fams = ("dp_conv3d_wgrad_tc", "dp_conv3d_stack", "dp_conv3d_tc", "dp_gemm_tc")
        out["roofline_families"] = [{"kernel": k, "achieved": P.flops.get(k, 0.0) / (fam[k]["ms"] / 1e3) / 1e12, "unit": "TFLOP/s",
                                     "frac": P.flops.get(k, 0.0) / (fam[k]["ms"] / 1e3) / 1e12 / peaks["tflops"],
                                     "ms_per_step": fam[k]["ms"], "launches_per_step": fam[k]["launches"]}
                                    for k in tc_fams if k in fam]
        out["roofline"] = {"kernel": top[0], "launch": top[1], "bound": "tensor", "achieved": ach, "peak": peaks["tflops"],
                           "unit": "TFLOP/s", "frac": ach / peaks["tflops"], "traffic": None, "peak_source": peaks["source"],
                           "avg_launch_ms": top[2], "algorithmic_flops_per_launch": top[3], "share_of_step": top[2] / total_fam}
        out["kernel_ms_per_step"] = {k: round(v["ms"], 3) for k, v in sorted(fam.items())}
    del tr, dose
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=int(os.environ.get("DP_BENCH_BATCH", "8")))
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step sub-measurement of the default line")
    ap.add_argument("--workload", default="cascade", choices=["cascade", "train"],
                    help="cascade (default, the headline metric) or train (BASELINE.json configs[3]: DOSE-PYFER training step)")
    ap.add_argument("--sw-roi", type=int, default=0,
                    help="run the seg stage as the reference does: sliding 96^3-style windows of this ROI (0 = direct)")
    args = ap.parse_args()
    _claim_stdout()
    args.batch_given = any(a == "--batch" or a.startswith("--batch=") for a in sys.argv[1:])
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.workload == "train":
        return run_train(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import __graft_entry__
    from dose_prediction_b200 import synth
    from dose_prediction_b200.cascade import CascadePlan

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _build_once(__graft_entry__, dist, world, local)

    B, S = args.batch, args.size
    seg, dose = build_models(S, dev, seg_size=args.sw_roi or None)
    casc = CascadePlan(seg, dose, B, S, dev, graph=False, sw_roi=args.sw_roi or None)
    plan = casc.plan
    # synthetic volumes: this rank's shard of a job of world*B volumes (weak scaling), pinned on the host
    vols = synth.make_batch(B, S, seed=1234 + rank * B)
    ct_h, ptv_h = vols["ct"].pin_memory(), vols["ptv"].pin_memory()
    out_h = torch.empty((B, 1, S, S, S), dtype=torch.float32).pin_memory()
    casc.ct.copy_(ct_h)
    casc.ptv.copy_(ptv_h)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        casc.run()
    torch.cuda.synchronize(dev)
    plan.check_device_errors()
    if not args.no_graph:
        plan.capture()
        casc.run()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(casc.run, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end through the public API (CascadeStream): every step copies its inputs from pinned host memory
    # and its dose result back to pinned host memory; copies of neighbouring steps overlap the kernels
    from dose_prediction_b200.cascade import CascadeStream
    pipe = CascadeStream(casc)

    def e2e_step():
        pipe.submit(ct_h, ptv_h)
    for _ in range(2):
        e2e_step()
    pipe.flush()
    torch.cuda.synchronize(dev)

    def e2e_run():
        for _ in range(args.steps):
            e2e_step()
        pipe.flush()                                   # the last result has reached the host inside the timed region
    ms_e2e = timed(e2e_run, 1)
    e2e_val = world * B * args.steps / (ms_e2e / 1e3)

    # ---- per-kernel-family device time (instrumented eager replay, CUDA events on the launch stream)
    fam = plan.profile_families(repeats=2)
    peaks = _peaks()
    total_fam = sum(v["ms"] for v in fam.values()) or 1.0
    names = {"dp_conv3d_stack": "conv3d_stack_kernel (tcgen05 depth-stacked implicit-GEMM conv, C_out 16/32)",
             "dp_conv3d_tc": "conv3d_tc_kernel (tcgen05 implicit-GEMM conv, C_out >= 64 / dilated)",
             "dp_gemm_tc": "gemm_tc_kernel (tcgen05 GEMM: ViT linears, patch embedding, token deconvs)",
             "dp_gemm_patch_embed": "gemm_tc_kernel, patch-embedding mode (A tiles gathered from the c8 activation by 5-D TMA)",
             "dp_attention": "attention_kernel (tcgen05 QK^T / PV with the softmax in TMEM + smem)"}

    def tensor_roofline(key):
        ms_k = fam.get(key, {}).get("ms", 0.0)
        n_k = fam.get(key, {}).get("launches", 0)
        fl = plan.flops.get(key, 0.0)
        ach = fl / (ms_k / 1e3) / 1e12 if ms_k > 0 else 0.0
        return {"kernel": names[key], "bound": "tensor", "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s",
                "frac": ach / peaks["tflops"], "traffic": None, "peak_source": peaks["source"], "launches_per_step": n_k,
                "avg_launch_ms": ms_k / max(n_k, 1), "algorithmic_flops_per_step": fl, "share_of_step": ms_k / total_fam}

    dominant = max(names, key=lambda k: fam.get(k, {}).get("ms", 0.0))
    roofline_family = [tensor_roofline(k) for k in names]
    # gemm_tc_kernel over ALL its modes (round 1 reported the linears and the patch embedding as one family)
    g_ms = fam.get("dp_gemm_tc", {}).get("ms", 0.0) + fam.get("dp_gemm_patch_embed", {}).get("ms", 0.0)
    g_fl = plan.flops.get("dp_gemm_tc", 0.0) + plan.flops.get("dp_gemm_patch_embed", 0.0)
    if g_ms > 0:
        g_ach = g_fl / (g_ms / 1e3) / 1e12
        roofline_family.append({"kernel": "gemm_tc_kernel, all modes (ViT linears + token deconvs + patch embedding)", "bound": "tensor",
                                "achieved": g_ach, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": g_ach / peaks["tflops"],
                                "traffic": None, "peak_source": peaks["source"],
                                "launches_per_step": fam.get("dp_gemm_tc", {}).get("launches", 0) + fam.get("dp_gemm_patch_embed", {}).get("launches", 0),
                                "avg_launch_ms": None, "algorithmic_flops_per_step": g_fl, "share_of_step": g_ms / total_fam})
    family_headline = tensor_roofline(dominant)
    # the dominant KERNEL LAUNCH: heaviest launch shape of the dominant family, timed live (CUDA events)
    per_launch = [r for r in plan.profile_launches() if r[0] == dominant]
    by_shape = {}
    for n_, label, ms_l, fl in per_launch:
        d = by_shape.setdefault(label, {"ms": 0.0, "n": 0, "flops": fl})
        d["ms"] += ms_l
        d["n"] += 1
    top_label, top = max(by_shape.items(), key=lambda kv: kv[1]["ms"])
    avg_ms = top["ms"] / top["n"]
    ach = top["flops"] / (avg_ms / 1e3) / 1e12
    traffic = family_traffic = None
    for tname in ("r3_traffic.json", "r2_traffic.json", "r1_traffic.json"):       # ncu dram__bytes_read.sum + dram__bytes_write.sum (profiles/)
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            traffic = tj.get(f"{dominant} {top_label}")
            family_traffic = tj.get(f"family {dominant} batch {B} size {S}")
            break
    # headline = the dominant kernel over ALL its launches of the step (every shape it runs, the slow 3^3 ones included);
    # its heaviest launch shape is reported beside it as a sub-field
    best_shape = {"launch": top_label, "achieved": ach, "frac": ach / peaks["tflops"], "launches_per_step": top["n"],
                  "avg_launch_ms": avg_ms, "algorithmic_flops_per_launch": top["flops"], "share_of_step": top["ms"] / total_fam,
                  "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write)",
                  "frac_of_burst": (ach / peaks["tflops_burst"]) if peaks["tflops_burst"] else None}
    roofline = dict(family_headline)
    roofline.update({"scope": "all launches of the dominant kernel in one step (algorithmic FLOPs / summed CUDA-event time)",
                     "traffic": family_traffic, "traffic_unit": "DRAM bytes per step over all launches of this kernel (ncu)",
                     "peak_burst": peaks["tflops_burst"],
                     "frac_of_burst": (family_headline["achieved"] / peaks["tflops_burst"]) if peaks["tflops_burst"] else None,
                     "heaviest_launch_shape": best_shape})
    # HBM-bound fused kernels: algorithmic bytes (every input and output moved once) / summed launch time
    hbm_families = [{"kernel": k, "bound": "hbm", "achieved": plan.bytes[k] / (fam[k]["ms"] / 1e3) / 1e9, "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "frac": plan.bytes[k] / (fam[k]["ms"] / 1e3) / 1e9 / peaks["hbm_gbs"],
                     "algorithmic_bytes_per_step": plan.bytes[k], "ms_per_step": fam[k]["ms"], "launches_per_step": fam[k]["launches"]}
                    for k in sorted(plan.bytes, key=lambda k: -fam.get(k, {"ms": 0})["ms"]) if k in fam and fam[k]["ms"] > 0]
    conv_fl = plan.flops.get("dp_conv3d_stack", 0.0) + plan.flops.get("dp_conv3d_tc", 0.0)
    conv_ms = fam.get("dp_conv3d_stack", {}).get("ms", 0.0) + fam.get("dp_conv3d_tc", {}).get("ms", 0.0)
    conv_pct = conv_fl / (conv_ms / 1e3) / 1e12 / peaks["tflops"] if conv_ms > 0 else 0.0
    plan.check_device_errors()
    # ---- outputs of the benchmarked plan for the in-run parity check (volume 0 of this rank's batch)
    casc.run()
    torch.cuda.synchronize(dev)
    got = (casc.logits[:1].float().cpu(), casc.dose[:1].float().cpu(), vols["ptv"][:1], vols["ct"][:1]) if rank == 0 else None
    gpu_launches = plan.kernels_per_step
    bytes_alloc = plan.bytes_alloc
    seg_sd = {k: v.cpu() for k, v in seg.state_dict().items()}
    dose_sd = {k: v.cpu() for k, v in dose.state_dict().items()}
    # ---- BASELINE.json configs[3] in the same run (so the 1/2/4/8-GPU scaling runs carry the NCCL config): free the
    # cascade plan first, then the DOSE-PYFER training step, batch 2 per GPU
    train = None
    if not args.no_train and not args.sw_roi:
        del pipe, casc, plan, seg, dose
        torch.cuda.empty_cache()
        train = measure_train(2, S, min(args.steps, 10), 3, dev, world, rank, local)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f16", "data": "synthetic",
                "config": {"workload": f"cascade OAR-TRANSEG->hand-off->DOSE-PYFER inference, {S}^3, batch {B} per GPU "
                                       "(BASELINE.json configs[2]; configs[1] seg forward is its first half)",
                           "batch_per_gpu": B, "size": S,
                           "seg_stage": (f"sliding window ROI {args.sw_roi}, overlap 0.25 (as train_light_linked_model.py:152)"
                                         if args.sw_roi else "direct full-volume forward"),
                           "parallelism": f"volume-sharded x{world}, no collective",
                           "cuda_graph": not args.no_graph,
                           "l2": f"no flush: per-step working set {bytes_alloc / 2**30:.1f} GiB >> 126 MB L2"},
                "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": 2 * B * S ** 3 * 4, "d2h_bytes_per_step": B * S ** 3 * 4},
                "gpu_launches": gpu_launches * args.steps, "launches_per_step": gpu_launches,
                "roofline": roofline, "roofline_families": roofline_family, "conv_frac_of_tensor_peak": conv_pct,
                "roofline_hbm_families": hbm_families,
                "kernel_ms_per_step": {k: round(v["ms"], 3) for k, v in sorted(fam.items())},
                "clocks": clocks}
        if train is not None:
            line["train"] = train
        failed = False
        if world == 1 and not args.no_cpu_baseline and not args.sw_roi:
            cpu = CpuCascade(seg_sd, dose_sd, S)
            line["cpu_baseline"], vol0, cpu_out = cpu_baseline(cpu, S, seed=1234)
            line["parity"] = parity_check(cpu, vol0, cpu_out, *got)
            failed = not line["parity"]["ok"]
        _emit(line)
        if failed:
            sys.stderr.write("bench.py: PARITY FAILED " + json.dumps(line["parity"]) + "\n")
            sys.exit(3)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
