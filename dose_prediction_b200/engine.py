"""Host-side plan builder / executor over the C ABI (include/dose_b200.h).

A `Plan` is built once per (network, input shape): it walks the network graph symbolically, allocates
every activation buffer statically, packs the module's parameters into the layouts the kernels want
and records one (function, args) tuple per kernel launch.  `Plan.run()` replays the launches on the
current CUDA stream (optionally as one CUDA graph).  PyTorch is used for device memory, streams and
one-time weight re-layout only; every launch inside `run()` is one of our kernels.

Activation layout "c8": fp16 [N][C/8][D][H][W][8]; channel counts are padded to multiples of 16 so a
tensor-core K step (16 channels) is two channel blocks.  A buffer may carry an fp16 "lo" residue part
(x ~ hi + lo, ~22 mantissa bits) for the precision-critical layers.
"""
import ctypes
import math
import os

import torch

from . import _lib

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_MISH, ACT_GELU = 0, 1, 2, 3, 4
ACT_ID = {None: ACT_NONE, "none": ACT_NONE, "relu": ACT_RELU, "lrelu": ACT_LRELU, "mish": ACT_MISH, "gelu": ACT_GELU}
STATS_DOUBLES = 1 << 20
STACK_TILES = int(os.environ.get("DP_STACK_TILES", "0"))
FOLD_P3 = os.environ.get("DP_FOLD_P3", "1") != "0"
CONV_C1 = os.environ.get("DP_CONV_C1", "1") != "0"              # bring-up switch: one-channel 3^3 conv + closed-form residual (seg encoder1)
POINTWISE_TCK = os.environ.get("DP_POINTWISE_TCK", "1") != "0"  # bring-up switch: 1^3 convs on tcgen05 with normalise-on-load
COMPACT = os.environ.get("DP_COMPACT", "1") != "0"             # liveness-packed activation arena for inference plans (Plan.compact)
FUSE_HEADS = os.environ.get("DP_FUSE_HEADS", "1") != "0"         # bring-up switch: 1^3 heads folded into the producing norm_act
STACK_SPLIT_HALF = os.environ.get("DP_STACK_SPLIT_HALF", "1") != "0"   # bring-up switch: split-half folded 3^3 stacked conv
DEPTH_PAIR = os.environ.get("DP_DEPTH_PAIR", "1") != "0"       # plain conv kernel: two output planes per tile for 7^3, C_out = 64
FOLD_TC_MAX = int(os.environ.get("DP_FOLD_TC_MAX", "32"))        # plain tcgen05 conv: fold [W_hi | W_lo] into N up to this C_out
POINTWISE_CW = os.environ.get("DP_POINTWISE_CW", "1") != "0"     # bring-up switch: constant-bank weights for static 1^3 convs
POINTWISE_TC = os.environ.get("DP_POINTWISE_TC", "1") != "0"    # bring-up switch: wide coarse-level 1^3 convs on the tensor cores
DECONV_TC = os.environ.get("DP_DECONV_TC", "1") != "0"          # bring-up switch: k2s2 transposed convs of c8 inputs on the tensor cores
STACKED_CONV = os.environ.get("DP_STACKED_CONV", "1") != "0"    # bring-up switch between the two tcgen05 conv kernels
EPS = 1e-5


def ceil_div(a, b):
    return (a + b - 1) // b


def blocks16(C):
    """channel blocks reserved for C logical channels (padded to a multiple of 16)."""
    return 2 * ceil_div(C, 16)


class Act:
    """C logical channels living in channel blocks [cb_off, cb_off+blocks16(C)) of a c8 fp16 buffer;
    lo_off = first block of the fp16 residue part, or None."""
    __slots__ = ("buf", "cb_off", "C", "lo_off")

    def __init__(self, buf, cb_off, C, lo_off=None):
        self.buf, self.cb_off, self.C, self.lo_off = buf, cb_off, C, lo_off

    @property
    def N(self):
        return self.buf.shape[0]

    @property
    def cb_total(self):
        return self.buf.shape[1]

    @property
    def dims(self):
        return tuple(self.buf.shape[2:5])

    @property
    def vox(self):
        d = self.dims
        return d[0] * d[1] * d[2]

    @property
    def hi_ptr(self):
        return self.buf.data_ptr()

    @property
    def lo_ptr(self):
        if self.lo_off is None:
            return None
        return self.buf.data_ptr() + (self.lo_off - self.cb_off) * self.vox * 16


class Raw:
    """fp32 c8 pre-normalisation tensor + its per-(n,c) {sum, sumsq} statistics."""
    __slots__ = ("t", "C", "stats")

    def __init__(self, t, C, stats):
        self.t, self.C, self.stats = t, C, stats

    @property
    def cb_total(self):
        return self.t.shape[1]


class Tokens:
    """[B, T, C] fp16 token matrix read in place by the transposed convolutions (proj_feat is a no-op)."""
    __slots__ = ("t", "grid")

    def __init__(self, t, grid):
        self.t, self.grid = t, grid


class LiveWeight:
    """a conv weight parameter that changes between replays (training): calling it gives the tensor the conv
    should see now (the parameter, or its transposed + flipped view for the data-gradient conv)."""

    def __init__(self, param, transpose_flip=False):
        self.param, self.transpose_flip = param, transpose_flip

    def __call__(self):
        w = self.param.detach()
        return w.flip(2, 3, 4).transpose(0, 1) if self.transpose_flip else w


class Plan:
    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("dose_prediction_b200 runs on CUDA devices only (no CPU fallback)")
        self.lib = _lib.lib()
        self.steps = []
        self.stats = torch.zeros(STATS_DOUBLES, dtype=torch.float64, device=self.device)
        self.stats_used = 0
        self.err = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.keep = []
        self.pool = {}
        self.bytes_alloc = 0
        self.graph = None
        self.launches = 0
        self.flops = {}          # algorithmic FLOPs (2*MAC of the reference op) per kernel family, per replay
        self.step_flops = []     # the same per recorded launch (parallel to self.steps)
        self.step_kernels = []   # device kernels per recorded C-ABI call (parallel to self.steps; 0 for host-side steps)
        self._pending_flops = 0.0
        self.bytes = {}          # algorithmic HBM bytes (inputs + outputs once) per HBM-bound kernel family, per replay
        self.training = False    # training plans re-derive packed weights from the live parameters every step
        self.refresh = []        # (packed tensor, function returning its new value)
        self.refresh_launches = []   # (C entry point name, args): device-side re-packing of live parameters
        self._na_producer = {}       # (buffer ptr, cb_off) -> the plain norm_act launch that wrote that activation
        self._poolable = []          # activation / pre-norm buffers compact() may overlay by liveness (fully written tensors only)

    # ------------------------------------------------------------------ memory
    def zeros(self, shape, dtype):
        t = torch.zeros(shape, dtype=dtype, device=self.device)
        self.bytes_alloc += t.numel() * t.element_size()
        self.keep.append(t)
        return t

    def new_buf(self, N, cb_total, dims, poolable=False):
        t = self.zeros((N, cb_total) + tuple(dims) + (8,), torch.float16)
        if poolable:
            self._poolable.append(t)
        return t

    def new_act(self, N, C, dims, lo=False):
        nb = blocks16(C)
        # channel counts that are not a multiple of 16 leave padding blocks nobody writes: they rely on the zero fill of
        # their own allocation and are never overlaid with other tensors
        buf = self.new_buf(N, nb * (2 if lo else 1), dims, poolable=(C % 16 == 0))
        return Act(buf, 0, C, nb if lo else None)

    def new_concat(self, N, Cs, dims, lo=False):
        """one buffer holding the channel concatenation of len(Cs) tensors (torch.cat made free)."""
        nbs = [blocks16(c) for c in Cs]
        tot = sum(nbs)
        buf = self.new_buf(N, tot * (2 if lo else 1), dims, poolable=all(c % 16 == 0 for c in Cs))
        acts, off = [], 0
        for c, nb in zip(Cs, nbs):
            acts.append(Act(buf, off, c, tot + off if lo else None))
            off += nb
        return acts

    def new_stats(self, N, C):
        n = N * C * 2
        if self.stats_used + n > STATS_DOUBLES:
            raise RuntimeError("statistics arena exhausted")
        s = self.stats[self.stats_used:self.stats_used + n]
        self.stats_used += n
        return s

    def get_raw(self, N, C, dims, with_stats=True):
        key = (N, blocks16(C)) + tuple(dims)
        free = self.pool.setdefault(key, [])
        if free:
            t = free.pop()
        else:
            t = self.zeros((N, blocks16(C)) + tuple(dims) + (8,), torch.float32)
            if C % 16 == 0:
                self._poolable.append(t)
        return Raw(t, C, self.new_stats(N, C) if with_stats else None)

    def release(self, raw):
        key = (raw.t.shape[0], raw.t.shape[1]) + tuple(raw.t.shape[2:5])
        self.pool[key].append(raw.t)

    def dev(self, t, dtype=torch.float32):
        t = t.detach().to(device=self.device, dtype=dtype).contiguous()
        self.keep.append(t)
        return t

    # ------------------------------------------------------------------ recording / replay
    def add(self, name, *args, kernels=1):
        """record one C-ABI call; `kernels` = device kernels that call launches (for the gpu_launches claim)."""
        self.step_kernels.append(kernels)
        self.steps.append((getattr(self.lib, name), args, name))
        self.step_flops.append(self._pending_flops)
        self._pending_flops = 0.0

    def add_zero(self, t):
        """re-zero a (split-K / atomic) accumulation target at this point of every replay."""
        self.steps.append((None, (t,), "zero"))
        self.step_flops.append(0.0)
        self.step_kernels.append(0)

    @property
    def kernels_per_step(self):
        return sum(self.step_kernels)

    def add_py(self, fn, capturable=True):
        """host-side plumbing at this point of every replay.  capturable: the function only enqueues device work on the
        current stream (torch ops), so it may be recorded into a CUDA graph; False for collectives / anything that must
        run on the host every step (such a step splits the launch list into separately captured graph segments)."""
        self.steps.append((None, (fn,), "py" if capturable else "py_host"))
        self.step_flops.append(0.0)
        self.step_kernels.append(0)

    def derived(self, fn):
        """a tensor computed from live parameters (packed / transposed / fp16 copy): kept, and in training
        plans recomputed in place by refresh_weights() after every optimizer step."""
        t = fn().contiguous()
        self.keep.append(t)
        if self.training:
            self.refresh.append((t, fn))
        return t

    def refresh_weights(self):
        with torch.cuda.device(self.device):
            self._refresh_weights()

    def _refresh_weights(self):
        s = torch.cuda.current_stream(self.device).cuda_stream
        for name, args in self.refresh_launches:
            rc = getattr(self.lib, name)(*args, s)
            if rc:
                _lib.check(rc, name)
        with torch.no_grad():
            for t, fn in self.refresh:
                t.copy_(fn())

    def derived_f16(self, param, transpose=False):
        """fp16 copy of a live fp32 parameter ([out,in], or its transpose), refreshed on the device every step."""
        if transpose or not self.training:
            return self.derived(lambda: (param.detach().to(self.device).t() if transpose else param.detach().to(self.device)).half())
        src = param.detach()
        t = src.half().contiguous()
        self.keep.append(t)
        self.refresh_launches.append(("dp_cast_f16", (src.data_ptr(), src.numel(), t.data_ptr())))
        return t

    def run(self):
        # kernels launch on the CURRENT device: make it this plan's device for the duration of the replay, so that one
        # process can drive several GPUs (one plan / handle per device) without the caller switching devices
        with torch.cuda.device(self.device):
            self._run()

    def run_range(self, start, stop, zero_stats):
        """replay steps[start:stop] only (the autograd shims run the forward and the backward halves of a training
        launch list separately); zero_stats: this range begins a new forward pass"""
        with torch.cuda.device(self.device):
            self._run(start, stop, zero_stats)

    def _run(self, start=0, stop=None, zero_stats=True):
        s = torch.cuda.current_stream(self.device).cuda_stream
        if self.stats_used and zero_stats:
            self.stats[:self.stats_used].zero_()
        n = 0
        for fn, args, name in (self.steps if (start == 0 and stop is None) else self.steps[start:stop]):
            if fn is None:
                if name in ("py", "py_host"):
                    args[0]()
                else:
                    args[0].zero_()
                continue
            rc = fn(*args, s)
            if rc:
                _lib.check(rc, name)
            n += 1
        self.launches = n

    def count_bytes(self, name, nbytes):
        self.bytes[name] = self.bytes.get(name, 0.0) + float(nbytes)

    def count_flops(self, name, flops):
        """call right before the add() of the launch that performs `flops` algorithmic FLOPs."""
        self.flops[name] = self.flops.get(name, 0.0) + float(flops)
        self._pending_flops = float(flops)

    def profile_launches(self):
        """per-launch device time (CUDA events): list of (family, label, ms) in schedule order."""
        stream = torch.cuda.current_stream(self.device)
        s = stream.cuda_stream
        if self.stats_used:
            self.stats[:self.stats_used].zero_()
        evs = []
        for fn, args, name in self.steps:
            if fn is None:
                args[0]() if name in ("py", "py_host") else args[0].zero_()
                continue
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            rc = fn(*args, s)
            e1.record(stream)
            if rc:
                _lib.check(rc, name)
            label = ""
            if name == "dp_conv3d_stack":
                label = f"N{args[5]} {args[6]}x{args[7]}x{args[8]} cout{args[9]} k{args[10]} chunks{args[3]}"
            elif name == "dp_conv3d_tc":
                label = f"N{args[5]} {args[6]}x{args[7]}x{args[8]} cout{args[9]} k{args[10]} dil{args[11]} chunks{args[3]}"
            elif name == "dp_gemm_tc":
                label = f"M{args[2]} N{args[3]} K{args[4]} batch{args[5]} split{args[12]}"
            elif name == "dp_attention":
                label = f"B{args[3]} heads{args[4]} T{args[5]} hd{args[7]}"
            elif name == "dp_conv3d_direct":
                label = f"cin{args[4]} N{args[5]} {args[6]}x{args[7]}x{args[8]} k{args[9]} s{args[10]} cout{args[16]}"
            elif name == "dp_deconv2x":
                label = f"cin{args[5]} cout{args[6]} N{args[7]} {args[8]}x{args[9]}x{args[10]}"
            evs.append((name, label, e0, e1))
        torch.cuda.synchronize(self.device)
        fl = [f for (fn, _, _), f in zip(self.steps, self.step_flops) if fn is not None]
        return [(n, l, e0.elapsed_time(e1), f) for (n, l, e0, e1), f in zip(evs, fl)]

    def profile_families(self, repeats=1):
        """device time per kernel family: eager replay with a CUDA-event pair around every launch on the
        launch stream; returns {family: {"ms": per-replay total, "launches": n}}."""
        stream = torch.cuda.current_stream(self.device)
        s = stream.cuda_stream
        out = {}
        for _ in range(repeats):
            if self.stats_used:
                self.stats[:self.stats_used].zero_()
            evs = []
            for fn, args, name in self.steps:
                if fn is None:
                    args[0]() if name in ("py", "py_host") else args[0].zero_()
                    continue
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                rc = fn(*args, s)
                e1.record(stream)
                if rc:
                    _lib.check(rc, name)
                evs.append((name, e0, e1))
            torch.cuda.synchronize(self.device)
            for name, e0, e1 in evs:
                d = out.setdefault(name, {"ms": 0.0, "launches": 0})
                d["ms"] += e0.elapsed_time(e1) / repeats
                d["launches"] += 1
        for d in out.values():
            d["launches"] //= repeats
        return out

    def check_device_errors(self):
        if int(self.err.item()) != 0:
            raise RuntimeError("libdose_b200: an in-kernel mbarrier wait timed out (pipeline protocol error)")

    def capture(self):
        """Record run() into a CUDA graph (launch-bound replay); falls back to nothing — errors raise."""
        with torch.cuda.device(self.device):
            st = torch.cuda.Stream(self.device)
            st.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(st):
                self.run()
            torch.cuda.current_stream(self.device).wait_stream(st)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.run()
            self.graph = g

    def replay(self):
        if self.graph is not None:
            with torch.cuda.device(self.device):
                self.graph.replay()
        else:
            self.run()

    # ------------------------------------------------------------------ liveness-packed activation arena
    def compact(self):
        """Overlay the plan's activation / pre-norm buffers in ONE arena by liveness.  Every launch is a
        (C function, args) tuple, so a buffer's lifetime is [first, last] launch whose arguments point into it; buffers whose
        lifetimes do not overlap share addresses (largest first, first fit).  Launch arguments (plain pointers and the
        pointer arrays of the multi-source entry points) are rewritten and the tensors re-bound to the arena with set_(), so
        Act / Raw objects held elsewhere stay valid.  Only fully written tensors take part (see new_act); plans with
        host-side steps (trainers, sliding-window accumulation) are left as they are.  Returns (bytes before, bytes after)."""
        if self.training or self.graph is not None or not self._poolable:
            return None
        if any(fn is None and name in ("py", "py_host") for fn, _, name in self.steps):
            return None
        bufs = []
        for t in {id(t): t for t in self._poolable}.values():
            nbytes = t.numel() * t.element_size()
            bufs.append({"t": t, "ptr": t.data_ptr(), "n": nbytes, "first": None, "last": None})
        bufs.sort(key=lambda b: b["ptr"])
        import bisect
        starts = [b["ptr"] for b in bufs]

        def find(v):
            if not isinstance(v, int) or v < (1 << 40):
                return None
            i = bisect.bisect_right(starts, v) - 1
            if i >= 0 and v < bufs[i]["ptr"] + bufs[i]["n"]:
                return bufs[i]
            return None

        def touch(b, i):
            b["first"] = i if b["first"] is None else b["first"]
            b["last"] = i

        arrays = {}
        for i, (fn, args, name) in enumerate(self.steps):
            if fn is None:                      # "zero" step: the tensor it re-zeroes is in use at this point
                b = find(args[0].data_ptr()) if isinstance(args[0], torch.Tensor) else None
                if b is not None:
                    touch(b, i)
                continue
            for a in args:
                if isinstance(a, ctypes.Array) and a._type_ is ctypes.c_void_p:
                    arrays[id(a)] = a
                    for v in a:
                        b = find(v)
                        if b is not None:
                            touch(b, i)
                else:
                    b = find(a)
                    if b is not None:
                        touch(b, i)
        live = [b for b in bufs if b["first"] is not None]
        if not live:
            return None
        placed = []
        for b in sorted(live, key=lambda b: -b["n"]):
            size = (b["n"] + 1023) // 1024 * 1024
            taken = sorted((q["off"], q["off"] + q["size"]) for q in placed
                           if not (q["last"] < b["first"] or b["last"] < q["first"]))
            off = 0
            for lo, hi in taken:
                if off + size <= lo:
                    break
                off = max(off, hi)
            b["off"], b["size"] = off, size
            placed.append(b)
        total = max(b["off"] + b["size"] for b in placed)
        before = sum(b["n"] for b in live)
        arena = torch.zeros(total, dtype=torch.uint8, device=self.device)
        base = arena.data_ptr()

        def remap(v):
            b = find(v)
            if b is None or "off" not in b:
                return v
            return base + b["off"] + (v - b["ptr"])
        for a in arrays.values():
            for j in range(len(a)):
                if a[j] is not None:
                    a[j] = remap(a[j])
        self.steps = [(fn, args if fn is None else tuple(remap(a) for a in args), name) for fn, args, name in self.steps]
        self._na_producer = {}
        pooled = {id(b["t"]) for b in placed}
        for b in placed:
            t = b["t"]
            t.set_(arena[b["off"]:b["off"] + b["n"]].view(t.dtype).view(t.shape))
        self.keep = [k for k in self.keep if not (isinstance(k, torch.Tensor) and id(k) in pooled)]
        self.keep.append(arena)
        self.pool = {}
        self.bytes_alloc += total - before
        self.arena_bytes = (before, total)
        return before, total

    # ------------------------------------------------------------------ weight packing
    def pack_conv_tc(self, w, parts, mode, stacked=False, _values_only=False, split_half=False, dpair=False):
        """w [Co,Ci,k,k,k] fp32 (or a function returning it) -> fp16 [kd][chunk][kh][kw][2][Co][8] + per-chunk
        input block table.  mode p1: x_hi.W_hi;  p2: + x_lo.W_hi;  p3: + x_hi.W_lo  (operand splitting by K expansion)."""
        w_src = w
        w = (w() if callable(w) else w).detach().to(self.device, torch.float32)
        Co, Ci, k = w.shape[0], w.shape[1], w.shape[2]
        assert Co % 16 == 0, "tensor-core conv needs C_out % 16 == 0"
        whi = w.half().float()
        wlo = (w - whi).half().float()
        if mode == "p3f":     # 3-term split folded into N: hi chunks see [W_hi | W_lo], lo chunks [W_hi | 0]
            terms = [("hi", torch.cat((whi, wlo), 0)), ("lo", torch.cat((whi, torch.zeros_like(whi)), 0))]
            Co = 2 * Co
        else:
            terms = {"p1": [("hi", whi)], "p2": [("hi", whi), ("lo", whi)],
                     "p3": [("hi", whi), ("lo", whi), ("hi", wlo)]}[mode]
        mats, chunks = [], []
        for which, wt in terms:
            base = 0
            for a in parts:
                first = a.cb_off if which == "hi" else a.lo_off
                assert first is not None, "operand split requested but the input has no lo part"
                for j in range(ceil_div(a.C, 16)):
                    c0 = base + 16 * j
                    c1 = min(base + a.C, c0 + 16)
                    m = torch.zeros((Co, 16, k, k, k), device=self.device)
                    m[:, :c1 - c0] = wt[:, c0:c1]
                    mats.append(m)
                    chunks.append(first + 2 * j)
                base += a.C
            assert base == Ci, f"input parts carry {base} channels, weight expects {Ci}"
        nch = len(mats)
        W = torch.stack(mats, dim=1).view(Co, nch, 2, 8, k, k, k)
        if stacked:
            # depth taps become MMA columns; G = k+1 pre-rotated copies, ring slot s of copy r holding depth tap
            # kd = k-1-j with j = (s - r) mod G, and zeros for j == k: [rot][chunk][kh][kw][khalf][s][co][e]
            G = k + 1
            Z = torch.cat((W.flip(4), torch.zeros_like(W[:, :, :, :, :1])), dim=4)            # dim 4 indexed by j
            rots = []
            for r in range(G):
                idx = self._rot_index(G, r)
                rots.append(Z.index_select(4, idx).permute(1, 5, 6, 2, 4, 0, 3))
            W = torch.stack(rots, dim=0).contiguous()
            if split_half:
                # conv3d_stack3h_kernel (k = 3, fold): a tap's 128 rows are [W_hi: slot 0..3 x 16 | W_lo: slot 0..3 x 16]
                # instead of [W_hi | W_lo] per slot; tail = per rotation tap (0,0) of chunk 0 with the rows of the newest
                # ring plane (slot r + 2) zeroed as well (the first tap of an input plane, see the kernel header)
                assert k == 3 and Co == 32 and mode == "p3f"
                W = W.view(G, nch, k, k, 2, G, 2, 16, 8).permute(0, 1, 2, 3, 4, 6, 5, 7, 8).contiguous()
                first = W[:, 0, 0, 0].clone()                          # [rot][khalf][half][slot][16][8]
                for r in range(G):
                    first[r, :, :, (r + 2) % G] = 0
                W = torch.cat((W.reshape(-1), first.reshape(-1)))
            W = W.half()
        else:         # [kd][chunk][kh][kw][khalf][co][e]
            W = W.permute(4, 1, 5, 6, 2, 0, 3).contiguous()
            if dpair:
                # depth-pair mode of dp_conv3d_tc: k + 1 virtual depth taps, rows [W[kd = v] | W[kd = v - 1]] (zeros at the ends)
                z = torch.zeros_like(W[:1])
                W = torch.cat((torch.cat((W, z), 0), torch.cat((z, W), 0)), dim=5).contiguous()
            W = W.half()
        if _values_only:
            return W
        self.keep.append(W)
        if self.training:
            if isinstance(w_src, LiveWeight) and mode == "p1" and w_src.param.is_cuda and w_src.param.dtype == torch.float32:
                ci0, nci, base = [], [], 0
                for a in parts:
                    for j in range(ceil_div(a.C, 16)):
                        ci0.append(base + 16 * j)
                        nci.append(min(16, a.C - 16 * j))
                    base += a.C
                arrs = (_lib.int_array(ci0), _lib.int_array(nci))
                self.keep.append(arrs)
                self.refresh_launches.append(("dp_pack_conv_weight", (w_src.param.data_ptr(), Co, Ci, k, int(w_src.transpose_flip),
                                                                      *arrs, nch, int(stacked), W.data_ptr())))
            else:
                self.refresh.append((W, lambda: self.pack_conv_tc(w_src, parts, mode, stacked, _values_only=True,
                                                                  split_half=split_half, dpair=dpair)))
        assert max(chunks) < 256
        arr = (ctypes.c_uint8 * nch)(*chunks)
        self.keep.append(arr)
        return W, arr, nch

    def _rot_index(self, G, r):
        """device index tensor of ring rotation r (cached: re-packing may run inside a CUDA-graph capture, where a fresh
        host -> device copy is not allowed)"""
        cache = self.__dict__.setdefault("_rot_cache", {})
        if (G, r) not in cache:
            cache[(G, r)] = torch.tensor([(sl - r) % G for sl in range(G)], device=self.device)
        return cache[(G, r)]

    def affine(self, Co, bias=None, bn=None):
        """epilogue y = acc*scale + shift from a conv bias and/or eval-mode BatchNorm3d (running stats)."""
        scale = torch.ones(Co, device=self.device)
        shift = torch.zeros(Co, device=self.device)
        if bias is not None:
            shift = bias.detach().to(self.device, torch.float32).clone()
        if bn is not None:
            g = bn.weight.detach().to(self.device, torch.float32)
            b = bn.bias.detach().to(self.device, torch.float32)
            rm = bn.running_mean.detach().to(self.device, torch.float32)
            rv = bn.running_var.detach().to(self.device, torch.float32)
            s = g / torch.sqrt(rv + bn.eps)
            shift = (shift - rm) * s + b
            scale = s
        scale, shift = scale.contiguous(), shift.contiguous()
        self.keep += [scale, shift]
        return scale, shift

    # ------------------------------------------------------------------ kernel emitters
    def pack_input(self, x, out):
        """x: static fp32 NCDHW input tensor [N,C,...] -> out Act (hi[/lo])."""
        N, C = x.shape[0], x.shape[1]
        self.add("dp_pack_ncdhw", x.data_ptr(), N, C, out.vox, out.hi_ptr, out.lo_ptr, out.cb_total, out.cb_off)

    def unpack(self, a, dst):
        self.add("dp_unpack_c8", a.hi_ptr, a.lo_ptr, a.cb_total, a.cb_off, a.N, a.C, a.vox, dst.data_ptr())

    def conv_tc(self, parts, weight, k, dil, mode, scale, shift, relu, out_raw=None, out_act=None, stats=None,
                tap_mask_fn=None):
        a0 = parts[0]
        D, H, W = a0.dims
        wshape = (weight() if callable(weight) else weight).shape
        Co = wshape[0]
        stacked = STACKED_CONV and dil == 1 and k in (3, 7) and Co in (16, 32) and tap_mask_fn is None
        fold = stacked and mode == "p3" and Co == 16 and FOLD_P3
        if fold and k == 3 and STACK_SPLIT_HALF:
            fold = 2                    # split-half column layout: x_lo chunks issue N = 64 (conv3d_stack3h_kernel)
        fold_tc = (not stacked) and mode == "p3" and Co <= FOLD_TC_MAX and FOLD_P3
        # 7^3 convs with C_out = 64 (plain kernel, N = 64 MMAs): two output planes per tile -> N = 128 (dp_conv3d_tc fold = 2)
        dpair = (DEPTH_PAIR and not stacked and not self.training and mode == "p1" and k == 7 and Co == 64 and dil == 1
                 and tap_mask_fn is None and D >= 2)
        wp, chunks, nch = self.pack_conv_tc(weight, parts, "p3f" if (fold or fold_tc) else mode, stacked=stacked,
                                            split_half=(fold == 2), dpair=dpair)
        if out_raw is not None:
            of32, ohi, olo, cbt, cbo = out_raw.t.data_ptr(), None, None, out_raw.cb_total, 0
            st = out_raw.stats if stats is None else stats
        else:
            of32, ohi, olo, cbt, cbo = None, out_act.hi_ptr, out_act.lo_ptr, out_act.cb_total, out_act.cb_off
            st = stats
        flops = 2.0 * a0.N * D * H * W * k ** 3 * wshape[1] * Co
        if stacked:
            self.count_flops("dp_conv3d_stack", flops)
            self.add("dp_conv3d_stack", a0.buf.data_ptr(), a0.cb_total, chunks, nch, wp.data_ptr(), a0.N, D, H, W, Co, k,
                     scale.data_ptr(), shift.data_ptr(), int(relu), of32, ohi, olo, cbt, cbo,
                     st.data_ptr() if st is not None else None, self.err.data_ptr(), 0, STACK_TILES, int(fold))
            return
        masks = None
        if tap_mask_fn is not None:
            masks = (ctypes.c_uint32 * nch)(*tap_mask_fn(nch))
            self.keep.append(masks)
        self.count_flops("dp_conv3d_tc", flops if tap_mask_fn is None else tap_mask_fn.flops)
        self.add("dp_conv3d_tc", a0.buf.data_ptr(), a0.cb_total, chunks, nch, wp.data_ptr(), a0.N, D, H, W, Co, k, dil,
                 scale.data_ptr(), shift.data_ptr(), int(relu), of32, ohi, olo, cbt, cbo,
                 st.data_ptr() if st is not None else None, self.err.data_ptr(), 0, masks, 2 if dpair else int(fold_tc))

    def conv_direct(self, a, weight, k, stride, dil, scale, shift, relu, out_raw=None, out_act=None, stats=None):
        """generic direct conv on one Act whose C is a multiple of 8 (stride-2 convs of net_A)."""
        D, H, W = a.dims
        Co, Ci = weight.shape[0], weight.shape[1]
        cin = ceil_div(a.C, 8) * 8
        w = torch.zeros((k * k * k, cin, Co), device=self.device)
        w[:, :Ci] = weight.detach().to(self.device, torch.float32).permute(2, 3, 4, 1, 0).reshape(k * k * k, Ci, Co)
        self.keep.append(w)
        if out_raw is not None:
            of32, ohi, olo, cbt, cbo, st = out_raw.t.data_ptr(), None, None, out_raw.cb_total, 0, out_raw.stats
        else:
            of32, ohi, olo, cbt, cbo, st = None, out_act.hi_ptr, out_act.lo_ptr, out_act.cb_total, out_act.cb_off, stats
        self.add("dp_conv3d_direct", a.hi_ptr, a.lo_ptr, a.cb_total, a.cb_off, cin, a.N, D, H, W, k, stride, dil,
                 w.data_ptr(), scale.data_ptr(), shift.data_ptr(), int(relu), Co, of32, ohi, olo, cbt, cbo,
                 st.data_ptr() if st is not None else None)

    def norm_act(self, src, out, stats=None, gamma=None, beta=None, act=None, res=None, res_stats=None,
                 act_after_res=None, stats_out=None, s2d=None, identity=False):
        """out = act_after(act(IN(src)*gamma+beta) [+ IN?(res)]); src: Raw or Act; res: Act or Raw.
        identity=True: no normalisation (plain fp32 -> fp16 conversion + activation)."""
        if isinstance(src, Raw):
            rf, rh, rl, icb, ioff, C = src.t.data_ptr(), None, None, src.cb_total, 0, src.C
            N, vox = src.t.shape[0], src.t.shape[2] * src.t.shape[3] * src.t.shape[4]
            if stats is None and not identity:
                stats = src.stats
        else:
            rf, rh, rl, icb, ioff, C = None, src.hi_ptr, src.lo_ptr, src.cb_total, src.cb_off, src.C
            N, vox = src.N, src.vox
        eh = el = er = es = None
        ecb = eoff = 0
        if res is not None:
            if isinstance(res, Raw):
                er, ecb, eoff = res.t.data_ptr(), res.cb_total, 0
                es = (res.stats if res_stats is None else res_stats).data_ptr()
            else:
                eh, el, ecb, eoff = res.hi_ptr, res.lo_ptr, res.cb_total, res.cb_off
        oh = ol = None
        ocb = ooff = 0
        if out is not None:
            oh, ol, ocb, ooff = out.hi_ptr, out.lo_ptr, out.cb_total, out.cb_off
        cpad = ceil_div(C, 8) * 8

        def _bytes(t):          # bytes per voxel-channel of a tensor as stored
            if t is None:
                return 0
            return 4 if isinstance(t, Raw) else (4 if t.lo_off is not None else 2)
        self.count_bytes("dp_norm_act", N * vox * cpad * (_bytes(src) + _bytes(res) + _bytes(out) + _bytes(s2d)))
        if isinstance(src, Raw) and res is None and s2d is None and stats_out is None and out is not None and not identity \
                and not self.training:
            # remembered so that a 1^3 head reading `out` can be folded into this launch (Plan.head)
            self._na_producer[(out.buf.data_ptr(), out.cb_off)] = (len(self.steps), src, stats, gamma, beta, act, out, N, C, vox)
        self.add("dp_norm_act", rf, rh, rl, icb, ioff, stats.data_ptr() if stats is not None else None,
                 gamma.data_ptr() if gamma is not None else None, beta.data_ptr() if beta is not None else None,
                 ACT_ID[act], eh, el, er, es, ecb, eoff, ACT_ID[act_after_res], oh, ol, ocb, ooff,
                 stats_out.data_ptr() if stats_out is not None else None, N, C, vox,
                 s2d.hi_ptr if s2d is not None else None, s2d.lo_ptr if s2d is not None else None,
                 s2d.cb_total if s2d is not None else 0, s2d.cb_off if s2d is not None else 0,
                 *(tuple(2 * d for d in s2d.dims) if s2d is not None else (0, 0, 0)))

    def conv3_c1(self, x_planar, weight, bias, out_raw):
        """3^3 conv of a one-channel planar fp32 volume [N,1,D,H,W] -> Raw (16 channels), exact fp32 (csrc/conv_small.cu).
        Returns the [N][2] fp64 {sum x, sum x^2} statistics of the input (re-zeroed every replay)."""
        N, _, D, H, W = x_planar.shape
        assert weight.shape[0] == 16 and weight.shape[1] == 1 and weight.shape[2] == 3
        wh = weight.detach().to("cpu", torch.float32).contiguous()
        wa = (ctypes.c_float * wh.numel())(*wh.flatten().tolist())
        ba = None
        if bias is not None:
            ba = (ctypes.c_float * 16)(*bias.detach().to("cpu", torch.float32).tolist())
        xstats = self.new_stats(N, 1)
        self.keep.append((wa, ba))
        self.count_flops("dp_conv3d_c1", 2.0 * N * D * H * W * 27 * 16)
        self.count_bytes("dp_conv3d_c1", N * D * H * W * (4 + 64))
        self.add("dp_conv3d_c1", x_planar.data_ptr(), ctypes.cast(wa, ctypes.c_void_p),
                 ctypes.cast(ba, ctypes.c_void_p) if ba is not None else None, N, D, H, W, out_raw.t.data_ptr(), out_raw.cb_total,
                 out_raw.stats.data_ptr(), xstats.data_ptr())
        return xstats

    def norm_act_resx(self, src, out, x_planar, w1, xstats, act_after_res):
        """out = act(IN(src) + norm3(conv3(x))) for a one-channel x: the residual branch in closed form (dp_norm_act_resx)."""
        N, vox = src.t.shape[0], src.t.shape[2] * src.t.shape[3] * src.t.shape[4]
        w = self.dev(w1.reshape(-1))
        cpad = ceil_div(src.C, 8) * 8
        self.count_bytes("dp_norm_act", N * vox * (cpad * (4 + (4 if out.lo_off is not None else 2)) + 4))
        self.add("dp_norm_act_resx", src.t.data_ptr(), src.cb_total, src.stats.data_ptr(), x_planar.data_ptr(), w.data_ptr(),
                 xstats.data_ptr(), ACT_ID[act_after_res], out.hi_ptr, out.lo_ptr, out.cb_total, out.cb_off, N, src.C, vox)

    def head(self, a, weight, bias, out_planar):
        """1x1x1 conv C -> head_co (+bias) of activation `a` into NCDHW fp32 (dose heads, conv_out_A, seg logits).
        Folded into the norm_act launch that produced `a` when there is one; else a pointwise launch."""
        Co = weight.shape[0]
        prod = self._na_producer.get((a.buf.data_ptr(), a.cb_off)) if FUSE_HEADS else None
        if prod is None or Co > 8 or a.C > 128 or prod[8] != a.C:
            return self.pointwise([(a, None, None)], weight, bias, out_planar=out_planar)
        idx, src, stats, gamma, beta, act, out, N, C, vox = prod
        assert self.steps[idx][2] == "dp_norm_act"
        w = self.dev(weight.reshape(Co, -1))
        b = self.dev(bias) if bias is not None else None
        self.count_bytes("dp_norm_act", N * vox * 4 * Co)
        args = (src.t.data_ptr(), src.cb_total, stats.data_ptr(), gamma.data_ptr() if gamma is not None else None,
                beta.data_ptr() if beta is not None else None, ACT_ID[act], out.hi_ptr, out.lo_ptr, out.cb_total, out.cb_off,
                w.data_ptr(), b.data_ptr() if b is not None else None, Co, out_planar.data_ptr(), N, C, vox)
        self.steps[idx] = (getattr(self.lib, "dp_norm_act_head"), args, "dp_norm_act_head")
        del self._na_producer[(a.buf.data_ptr(), a.cb_off)]

    def pointwise_tc_ok(self, Cs, Co):
        """whether pointwise() will take the tcgen05 kernel for these source channel counts (the only path that can apply a
        two-stage normalisation on load)"""
        return (not self.training and POINTWISE_TCK and Co in (16, 32, 64) and sum(ceil_div(c, 8) for c in Cs) <= 16)

    def pointwise(self, srcs, weight, bias, out_raw=None, out_act=None, out_planar=None, out_act_fn=None):
        """1x1x1 conv over cat(srcs); srcs: list of (Act|Raw, stats|None, act|None) or, with a first normalisation stage
        applied before that one, (Raw, stats, act, stats0, act0) (tcgen05 path only: check pointwise_tc_ok)."""
        hi, lo, raw, cbt, cbo, Cs, sts, acts = [], [], [], [], [], [], [], []
        sts0, acts0 = [], []
        for t, st, act, *pre in srcs:
            sts0.append(pre[0].data_ptr() if pre else None)
            acts0.append(ACT_ID[pre[1]] if pre else 0)
            if isinstance(t, Raw):
                hi.append(None); lo.append(None); raw.append(t.t.data_ptr()); cbt.append(t.cb_total); cbo.append(0)
                N, vox = t.t.shape[0], t.t.shape[2] * t.t.shape[3] * t.t.shape[4]
            else:
                hi.append(t.hi_ptr); lo.append(t.lo_ptr); raw.append(None); cbt.append(t.cb_total); cbo.append(t.cb_off)
                N, vox = t.N, t.vox
            Cs.append(t.C)
            sts.append(st.data_ptr() if st is not None else None)
            acts.append(ACT_ID[act])
        Co = weight.shape[0]
        w = self.dev(weight.reshape(Co, -1))
        assert w.shape[1] == sum(Cs), f"pointwise: weight expects {w.shape[1]} inputs, sources carry {sum(Cs)}"
        b = self.dev(bias) if bias is not None else None
        arrs = [_lib.ptr_array(hi), _lib.ptr_array(lo), _lib.ptr_array(raw), _lib.int_array(cbt), _lib.int_array(cbo),
                _lib.int_array(Cs), _lib.ptr_array(sts), _lib.int_array(acts)]
        self.keep.append(arrs)
        two_stage = any(x is not None for x in sts0)
        arrs0 = [_lib.ptr_array(sts0), _lib.int_array(acts0)] if two_stage else [None, None]
        self.keep.append(arrs0)
        of = oh = ol = op = st = None
        ocb = ooff = 0
        if out_raw is not None:
            of, ocb = out_raw.t.data_ptr(), out_raw.cb_total
            st = out_raw.stats.data_ptr() if out_raw.stats is not None else None
        if out_act is not None:
            oh, ol, ocb, ooff = out_act.hi_ptr, out_act.lo_ptr, out_act.cb_total, out_act.cb_off
        if out_planar is not None:
            op = out_planar.data_ptr()
        in_b = sum((ceil_div(t.C, 8) * 8) * (4 if isinstance(t, Raw) or t.lo_off is not None else 2) for t, *_ in srcs)
        out_b = 0
        if out_raw is not None:
            out_b += 4 * ceil_div(Co, 8) * 8
        if out_act is not None:
            out_b += (4 if out_act.lo_off is not None else 2) * ceil_div(Co, 8) * 8
        if out_planar is not None:
            out_b += 4 * Co
        ncb = sum(ceil_div(c, 8) for c in Cs)
        if (not self.training and POINTWISE_TCK and Co in (16, 32, 64) and ncb <= 16 and out_planar is None
                and out_act_fn is None and ((out_raw is not None) != (out_act is not None))):
            # tcgen05 contraction; the sources' pending IN + activation are applied by the threads staging the A operand
            Wp = torch.zeros((Co, (ncb + (ncb & 1)) * 8), device=self.device)
            wl = weight.detach().reshape(Co, -1).to(self.device, torch.float32)
            kp = lb = 0
            for c in Cs:
                Wp[:, kp:kp + c] = wl[:, lb:lb + c]
                kp += ceil_div(c, 8) * 8
                lb += c
            whi = Wp.half()
            wlo = (Wp - whi.float()).half()
            wpk = torch.cat((whi, wlo), 0).view(2 * Co, -1, 8).permute(1, 0, 2).contiguous()      # [K/8][2*Co][8]
            self.keep.append(wpk)
            self.count_flops("dp_pointwise_tc", 2.0 * N * vox * sum(Cs) * Co)
            self.count_bytes("dp_pointwise_tc", N * vox * (in_b + out_b))
            self.add("dp_pointwise_tc", len(srcs), *arrs, *arrs0, wpk.data_ptr(), b.data_ptr() if b is not None else None, Co, N,
                     vox, of, oh, ol, ocb, ooff, st, self.err.data_ptr())
            return
        assert not two_stage, "pointwise: a two-stage normalisation on load needs the tcgen05 path (pointwise_tc_ok)"
        if (not self.training and POINTWISE_TC and ncb >= 12 and Co % 16 == 0 and out_raw is not None and out_act is None
                and out_planar is None and out_act_fn is None and all(c % 16 == 0 for c in Cs)):
            # wide 1^3 convs of the coarse levels (128 / 256 input channels, few voxels): the SIMT kernel is latency
            # bound there (0.2 ms per launch), so materialise act(IN(src)) once (a few MB) and contract on the tensor cores
            wide = any(isinstance(t, Raw) or t.lo_off is not None for t, _, _ in srcs)
            dims = tuple(srcs[0][0].t.shape[2:5]) if isinstance(srcs[0][0], Raw) else srcs[0][0].dims
            slots = self.new_concat(N, Cs, dims, lo=wide)
            for (t, st_, act), z in zip(srcs, slots):
                self.norm_act(t, z, stats=st_, act=act, identity=(st_ is None and not isinstance(t, Raw)))
            scale, shift = self.affine(Co, bias=bias)
            self.conv_tc(list(slots), weight.detach().reshape(Co, sum(Cs), 1, 1, 1), 1, 1, "p3" if wide else "p1", scale, shift,
                         False, out_raw=out_raw)
            return
        use_cw = not self.training and ncb in (1, 2, 3, 4, 6, 8) and POINTWISE_CW
        self.count_bytes("dp_pointwise_conv_cw" if use_cw else "dp_pointwise_conv", N * vox * (in_b * ceil_div(Co, 16) + out_b))
        if use_cw:
            # static weights: hand them over as kernel parameters (constant bank) instead of device pointers
            wh = weight.detach().reshape(Co, -1).to("cpu", torch.float32).contiguous()
            bh = bias.detach().to("cpu", torch.float32).contiguous() if bias is not None else None
            wa = (ctypes.c_float * wh.numel())(*wh.flatten().tolist())
            ba = (ctypes.c_float * Co)(*bh.tolist()) if bh is not None else None
            self.keep.append((wa, ba))
            self.add("dp_pointwise_conv_cw", len(srcs), *arrs, ctypes.cast(wa, ctypes.c_void_p),
                     ctypes.cast(ba, ctypes.c_void_p) if ba is not None else None, Co, N, vox, of, oh, ol, ocb, ooff, op, st,
                     ACT_ID[out_act_fn], kernels=ceil_div(Co, 16))
            return
        self.add("dp_pointwise_conv", len(srcs), *arrs, w.data_ptr(), b.data_ptr() if b is not None else None, Co, N,
                 vox, of, oh, ol, ocb, ooff, op, st, ACT_ID[out_act_fn])

    def deconv2x(self, src, weight, out):
        """ConvTranspose3d k2 s2 (no bias): src Act or Tokens -> out Act (usually a slice of a concat buffer)."""
        Ci, Co = weight.shape[0], weight.shape[1]
        if isinstance(src, Tokens) and Co % 16 == 0:
            B, T, C = src.t.shape
            D, H, W = src.grid
            assert C == Ci and C % 8 == 0
            wnk = self.derived(lambda: weight.detach().to(self.device).permute(2, 3, 4, 1, 0).reshape(8 * Co, Ci).half())
            self.count_flops("dp_deconv2x_gemm", 2.0 * B * T * Ci * Co * 8)
            self.add("dp_deconv2x_gemm", src.t.data_ptr(), wnk.data_ptr(), B, D, H, W, Ci, Co, out.hi_ptr, out.lo_ptr,
                     out.cb_total, out.cb_off, self.err.data_ptr())
            return
        if (not self.training) and isinstance(src, Act) and DECONV_TC and Ci % 16 == 0 and Co % 16 == 0 and Co <= 128 \
                and src.C == Ci:
            # 1^3 implicit GEMM with 8*Co columns (<= 256 per launch) and a scatter epilogue; operand precision follows
            # the input: hi/lo activations take the 3-term split like the 3^3 convs of the same net
            D, H, W = src.dims
            mode = "p3" if src.lo_off is not None else "p1"
            nq = min(8, 256 // Co)
            in_bpc = 4 if src.lo_off is not None else 2
            self.count_bytes("dp_deconv2x_tc", src.N * src.vox * (Ci * in_bpc + 8 * Co * (4 if out.lo_off is not None else 2)))
            for q0 in range(0, 8, nq):
                def wq(q0=q0):
                    wn = weight.detach().to(self.device, torch.float32).permute(2, 3, 4, 1, 0).reshape(8 * Co, Ci)
                    return wn[q0 * Co:(q0 + nq) * Co].reshape(nq * Co, Ci, 1, 1, 1)
                wp, chunks, nch = self.pack_conv_tc(wq, [src], mode)
                scale, shift = self.affine(nq * Co)
                self.count_flops("dp_deconv2x_tc", 2.0 * src.N * src.vox * Ci * Co * nq)
                self.add("dp_deconv2x_tc", src.buf.data_ptr(), src.cb_total, chunks, nch, wp.data_ptr(), src.N, D, H, W, Co,
                         q0, nq, scale.data_ptr(), shift.data_ptr(), out.hi_ptr, out.lo_ptr, out.cb_total, out.cb_off,
                         self.err.data_ptr())
            return
        w = self.derived(lambda: weight.detach().to(self.device, torch.float32).permute(2, 3, 4, 0, 1).reshape(8, Ci, Co))
        n_in = (src.t.shape[0] * src.t.shape[1]) if isinstance(src, Tokens) else src.N * src.vox
        in_bpc = 2 if isinstance(src, Tokens) or src.lo_off is None else 4
        nb_ = n_in * (Ci * in_bpc + 8 * blocks16(Co) * 8 * (4 if out.lo_off is not None else 2))
        cw_ = (not self.training) and (not isinstance(src, Tokens)) and Ci in (32, 64) and POINTWISE_CW
        self.count_bytes("dp_deconv2x_cw" if cw_ else "dp_deconv2x", nb_)
        if isinstance(src, Tokens):
            B, T, C = src.t.shape
            D, H, W = src.grid
            assert C == Ci and C % 8 == 0
            self.add("dp_deconv2x", src.t.data_ptr(), None, T * C, C, 8, Ci, Co, B, D, H, W, w.data_ptr(),
                     out.hi_ptr, out.lo_ptr, out.cb_total, out.cb_off)
        else:
            D, H, W = src.dims
            assert src.C == Ci and Ci % 8 == 0
            vox = src.vox
            base = src.buf.data_ptr() + src.cb_off * vox * 16
            lo = None if src.lo_off is None else src.buf.data_ptr() + src.lo_off * vox * 16
            if not self.training and Ci in (32, 64) and POINTWISE_CW:
                wh = weight.detach().to("cpu", torch.float32).permute(2, 3, 4, 0, 1).reshape(-1).contiguous()
                wa = (ctypes.c_float * wh.numel())(*wh.tolist())
                self.keep.append(wa)
                self.add("dp_deconv2x_cw", base, lo, src.cb_total * vox * 8, vox * 8, Ci, Co, src.N, D, H, W,
                         ctypes.cast(wa, ctypes.c_void_p), out.hi_ptr, out.lo_ptr, out.cb_total, out.cb_off,
                         kernels=ceil_div(Co, 16) * (1 if Ci == 32 else 2))
                return
            self.add("dp_deconv2x", base, lo, src.cb_total * vox * 8, 8, vox * 8, Ci, Co, src.N, D, H, W, w.data_ptr(),
                     out.hi_ptr, out.lo_ptr, out.cb_total, out.cb_off)

    def upsample2x(self, src, out):
        D, H, W = src.dims
        bpc = 4 if src.lo_off is not None else 2
        self.count_bytes("dp_upsample2x", src.N * src.vox * ceil_div(src.C, 8) * 8 * bpc * (1 + 8))
        self.add("dp_upsample2x", src.hi_ptr, src.lo_ptr, src.cb_total, src.cb_off, ceil_div(src.C, 8), src.N, D, H, W,
                 out.hi_ptr, out.lo_ptr, out.cb_total, out.cb_off)

    def gemm(self, A, B, M, N, K, *, batch=1, a_batch_rows=0, b_batch_rows=0, c_batch_stride=0, c_batch_period=0,
             c_batch_stride2=0, ldc=None, split_k=1, bias=None, rowvec=None, row_period=0, resid=None, alpha=1.0,
             act=None, out_f32=None, atomic=False, out_f16=None, qkv=None):
        self.count_flops("dp_gemm_tc", 2.0 * M * N * K * batch)
        ldc = N if ldc is None else ldc
        mode_qkv, heads, hd, T, vt_ld, q, k, vt, qs = 0, 0, 0, 0, 0, None, None, None, 1.0
        if qkv is not None:
            mode_qkv = 1
            heads, hd, T, q, k, vt, qs = qkv
            vt_ld = vt.shape[-1]
            q, k, vt = q.data_ptr(), k.data_ptr(), vt.data_ptr()
        p = lambda t: t.data_ptr() if t is not None else None
        self.add("dp_gemm_tc", p(A), p(B), M, N, K, batch, a_batch_rows, b_batch_rows, c_batch_stride, c_batch_period,
                 c_batch_stride2, ldc, split_k, p(bias), p(rowvec), row_period, p(resid), float(alpha), ACT_ID[act],
                 p(out_f32), int(atomic), p(out_f16), mode_qkv, heads, hd, T, vt_ld, q, k, vt, float(qs), self.err.data_ptr())

    def gemm_splitk(self, A, B, M, N, K, split_k, out_f32, bias=None, rowvec=None, row_period=0):
        """deterministic split-K: fp32 partials to a workspace, then one reduction pass (+bias, +rowvec)."""
        if split_k <= 1:
            return self.gemm(A, B, M, N, K, bias=bias, rowvec=rowvec, row_period=row_period, out_f32=out_f32)
        ws = self.zeros((split_k, M, N), torch.float32)
        self.gemm(A, B, M, N, K, split_k=split_k, out_f32=ws)
        self.add("dp_splitk_reduce", ws.data_ptr(), split_k, M, N, bias.data_ptr() if bias is not None else None,
                 rowvec.data_ptr() if rowvec is not None else None, row_period, out_f32.data_ptr())

    def layernorm(self, x, gamma, beta, rows, cols, out_f16=None, out_f32=None):
        self.add("dp_layernorm", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rows, cols,
                 out_f16.data_ptr() if out_f16 is not None else None, out_f32.data_ptr() if out_f32 is not None else None)

    def attention(self, q, k, vt, batch, heads, T, hd, out):
        """fused softmax(q k^T) v -> merged heads (monai SABlock.forward); q is pre-scaled by the qkv GEMM epilogue"""
        self.count_flops("dp_attention", 4.0 * batch * heads * T * T * hd)
        self.add("dp_attention", q.data_ptr(), k.data_ptr(), vt.data_ptr(), batch, heads, T, vt.shape[-1], hd,
                 out.data_ptr(), out.shape[-1], self.err.data_ptr())

    def softmax(self, s, rows, cols, p):
        self.add("dp_softmax", s.data_ptr(), rows, cols, s.shape[-1], p.data_ptr(), p.shape[-1])

    def patchify(self, a, ncb, out):
        D, H, W = a.dims
        self.add("dp_patchify", a.buf.data_ptr(), a.cb_total, a.cb_off, ncb, a.N, D, H, W, out.data_ptr())

    def crop_pack(self, src, R, windows, out):
        """windows: list of (b, x0, y0, z0); window i becomes batch entry i of `out` (an Act of R^3 volumes)."""
        C, S0, S1, S2 = src.shape[1], src.shape[2], src.shape[3], src.shape[4]
        arrs = [_lib.int_array([w[k] for w in windows]) for k in range(4)]
        self.keep.append(arrs)
        self.add("dp_crop_pack", src.data_ptr(), C, S0, S1, S2, R, len(windows), *arrs, out.hi_ptr, out.lo_ptr,
                 out.cb_total, out.cb_off)

    def window_add(self, win_logits, R, windows, out_sum):
        ncls = win_logits.shape[1]
        arrs = [_lib.int_array([w[k] for w in windows]) for k in range(4)] + [_lib.int_array(list(range(len(windows))))]
        self.keep.append(arrs)
        self.add("dp_window_add", win_logits.data_ptr(), ncls, R, len(windows), *arrs, out_sum.data_ptr(),
                 out_sum.shape[2], out_sum.shape[3], out_sum.shape[4])

    def div_count(self, data, count):
        vol = count.numel()
        self.add("dp_div_count", data.data_ptr(), count.data_ptr(), vol, data.numel() // vol)

    def handoff(self, logits, ptv, ct, out, structures=None):
        N, ncls, S = logits.shape[0], logits.shape[1], logits.shape[2]
        self.count_bytes("dp_handoff", N * S ** 3 * (4 * (ncls + 2) + 16 * (4 if out.lo_off is not None else 2)
                                                     + (36 if structures is not None else 0)))
        self.add("dp_handoff", logits.data_ptr(), ncls, ptv.data_ptr(), ct.data_ptr(), N, S, out.hi_ptr, out.lo_ptr,
                 out.cb_total, out.cb_off, structures.data_ptr() if structures is not None else None)
