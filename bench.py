#!/usr/bin/env python
"""bench.py - CRW walk fwd+bwd clips/s on the Kinetics-shaped batch (BASELINE.json configs[1]) and DAVIS-shaped
label-propagation frames/s (configs[3]), with the reference beside them.

    python bench.py --gpus N --steps K --warmup W                    # our arm (sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the UNMODIFIED reference on the host cores

One "step" (M1) = one pass of the hot path over one batch per GPU: patch mean-pool (a1) -> Linear head -> fused walk
forward+backward (a4-a6) -> head backward -> pool backward, on precomputed encoder maps (B=20 clips/GPU, N=49, T=4, C_e=512,
8x8 maps, tau 0.07, edge dropout 0.1).  Prints ONE JSON line.  Extra keys (one GPU): `label_prop` (M2, with
its own roofline / cpu_baseline / torch_gpu_baseline / e2e / clocks), `torch_gpu_baseline` (the reference's own PyTorch code
on the same B200), `superpixel` (configs[2]), `sweep` (configs[4] points), `e2e_module` (CRW(args)(x) with the ResNet-18).
Every leg is timed for >= 50 ms with CUDA events and reports a per-iteration median.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(B=20, N=49, T=4, Ce=512, D=128, H=8, W=8, tau=0.07, p=0.1)
LP = dict(C=256, h=60, w=107, n_ctx=20, n_tgt=37, k=10, radius=12, tau=0.07, L=4)   # 37 targets x 56 query tiles = 14 x 148 CTAs
SP = dict(B=8, T=8, C=512, SP=196, size=256)                                        # BASELINE configs[2]
SWEEP = [dict(T=4, N=49, B=2048), dict(T=8, N=256, B=256), dict(T=16, N=1024, B=16)]   # BASELINE configs[4]: B sized for >= 50 ms of work
RESNET18_GRAD_FLOATS = 11_176_512 + 512 * 128          # encoder + head parameters (44.97 MB)
MIN_MS = 50.0


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d["bf16_tflops"]), "measured"
    except Exception:
        return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm)}


def timed_median(fn, min_ms=MIN_MS, warm=3, max_iters=5000, min_iters=5):
    """CUDA-event time of fn(), one event pair per iteration, repeated until the timed iterations add up to >= min_ms.
    -> (median ms, iterations, total ms)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts, total = [], 0.0
    while (total < min_ms or len(ts) < min_iters) and len(ts) < max_iters:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        total += ts[-1]
    ts.sort()
    return ts[len(ts) // 2], len(ts), total


def count_kernels(fn):
    """Kernels one call of fn() launches, counted by the CUDA profiler (kineto), not assumed: -> (ours, others, names)."""
    try:
        from torch.profiler import ProfilerActivity, profile
        fn()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        names = [e.name for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA
                 and not e.name.lower().startswith(("memcpy", "memset"))]
        ours = [n for n in names if "crw::" in n or n.startswith("crw")]
        return len(ours), len(names) - len(ours), sorted(set(n.split("(")[0][:60] for n in names))
    except Exception as e:                                       # pragma: no cover - profiler unavailable
        return None, None, ["profiler unavailable: %r" % (e,)]


# --------------------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference (oracle/_ref, via oracle/ref_harness.py) or, if it is not in place, the oracle port
# --------------------------------------------------------------------------------------------------------------
def reference_walk_step(B, device):
    """-> (step(i) -> loss, kind): forward + backward of the reference's own CRW on the configs[1] maps on `device`."""
    c = CFG
    g = torch.Generator().manual_seed(1000)
    maps = torch.randn(B * c["N"], c["Ce"], c["T"], c["H"], c["W"], generator=g).to(device).requires_grad_(True)
    torch.manual_seed(0)
    head = torch.nn.Linear(c["Ce"], c["D"], bias=False).to(device)
    from oracle import ref_harness as RH
    if RH.available():
        return RH.walk_step(maps, head.weight.detach(), B, c["N"], c["T"], c["tau"], c["p"], device), "reference"
    from oracle import crw_oracle as O

    def step(i):
        torch.manual_seed(123 + i)
        q = O.patch_nodes(maps, head.weight, B)
        u12, u21p = O.draw_uniforms(B, c["N"], c["T"], device=device)
        loss, *_ = O.walk_loss(q, c["tau"], c["p"], u12, u21p)
        maps.grad = None
        head.weight.grad = None
        loss.mean().backward()
        return loss

    return step, "port"


def time_cpu_walk(B, steps, warmup):
    step, kind = reference_walk_step(B, "cpu")
    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        step(warmup + i)
    dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps * 1e3, kind


def host_threads():
    """All host cores, also under torchrun (which exports OMP_NUM_THREADS=1)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(n)
    return torch.get_num_threads()


def run_reference(args, rank, world):
    if rank != 0:
        return
    c = CFG
    cores = host_threads()
    steps, warmup = args.steps, args.warmup
    note = None
    if steps > 400:                                   # bounded: ~50 ms per 20-clip step on 16 cores
        note = "steps capped at 400 (asked %d)" % steps
        steps = 400
    val, ms, kind = time_cpu_walk(c["B"], steps, warmup)
    what = ("the unmodified reference CRW.forward + backward (code/model.py:334-415, encoder swapped for the precomputed maps)"
            if kind == "reference" else "oracle/crw_oracle.py port (oracle/_ref not built)")
    out = {"metric": "crw_walk_fwd_bwd_clips_per_s", "value": val, "unit": "clips/s", "n_gpus": args.gpus, "steps": steps,
           "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "impl": "reference",
           "config": workload_config(1),
           "cpu_baseline": {"value": val, "unit": "clips/s", "cores": cores, "kind": kind,
                            "sample": "%d steps of the full %d-clip batch: %s" % (steps, c["B"], what)},
           "e2e": {"value": val, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if note:
        out["note"] = note
    print(json.dumps(out), flush=True)


def workload_config(n):
    c = CFG
    return {"workload": "BASELINE configs[1]: Kinetics-shaped CRW hot path on precomputed encoder maps "
                        "(%d,%d,%d,%d,%d) fp32 per GPU = B %d clips x N %d patches, clip_len %d, temp %.2f, edge dropout %.1f; "
                        "pool -> head -> walk fwd+bwd -> head bwd -> pool bwd" % (c["B"] * c["N"], c["Ce"], c["T"], c["H"], c["W"],
                                                                                 c["B"], c["N"], c["T"], c["tau"], c["p"]),
            "global_batch": c["B"] * n, "parallelism": "dp%d" % n,
            "l2": "inputs (514 MB/GPU) exceed the 126 MB L2, no flush needed"}


# --------------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------------
class HotPath:
    """The step of the headline metric on one GPU, eager or as a captured CUDA graph.

    parts == 1: the whole 20-clip batch as one chain of launches (pool -> head -> walk -> head bwd -> pool bwd).
    parts >= 2: pipeline.PatchWalkPipeline - the clips split into micro-batches on staggered streams so that the HBM-bound pooling
    of one micro-batch runs beside the latency-bound walk of another (pooling confined to `pool_sms` SMs)."""

    def __init__(self, dev, rank, use_graph=True, parts=1, pool_sms=0, sizes=None, head_splits=1, walk_flags=0, pool_sms_bwd=None):
        from sapienza_video_contrastive_b200 import ops
        from sapienza_video_contrastive_b200.pipeline import PatchWalkPipeline
        self.ops, self.dev = ops, dev
        c = CFG
        g = torch.Generator(device="cpu").manual_seed(1000 + rank)
        # encoder-native physical layout (BN, T, C, H, W) (resnet.From3D), logically (BN, C, T, H, W)
        self.maps = torch.randn(c["B"] * c["N"], c["T"], c["Ce"], c["H"], c["W"], generator=g).to(dev)
        torch.manual_seed(0)
        self.head = torch.nn.Linear(c["Ce"], c["D"], bias=False).to(dev)
        self.rng_state = torch.tensor([123, 0], dtype=torch.int64, device=dev)
        self.ones = torch.ones(1, device=dev)
        self.sizes = [int(x) for x in sizes] if sizes else None
        self.parts = len(self.sizes) if self.sizes else int(parts)
        if self.sizes is None and self.parts > 1:
            self.sizes = [c["B"] // self.parts] * self.parts
        self.autograd_chain = self.sizes is None          # one chain of torch.autograd operators; else the direct-call pipeline
        self.graph = None
        self.use_graph = use_graph
        self.loss = torch.zeros(1, device=dev)
        if self.autograd_chain:
            self.maps_in = self.maps.clone().requires_grad_(True)
            self.pipe = None
        else:
            self.maps_parts = self._split(self.maps, clone=True)
            self.pipe = PatchWalkPipeline(self.head.weight, c["B"], c["N"], c["T"], c["tau"], c["p"], pool_sms=pool_sms, seed=123,
                                          device=dev, sizes=self.sizes, head_splits=head_splits, walk_flags=walk_flags, pool_sms_bwd=pool_sms_bwd)
            self.gmaps, self.ghead = None, None

    def _split(self, maps, clone=False):
        out, r0 = [], 0
        for b in self.sizes:
            r1 = r0 + b * CFG["N"]
            out.append(maps[r0:r1].clone() if clone else maps[r0:r1].detach())
            r0 = r1
        return out

    def set_inputs(self, maps):
        """Point the step at another (BN, T, C, H, W) device tensor (the e2e loop's freshly uploaded batch)."""
        if self.autograd_chain:
            self.maps_in = maps
        else:
            self.maps_parts = self._split(maps)

    def _step_eager(self):
        c = CFG
        ops = self.ops
        if self.pipe is not None:
            self.loss, self.gmaps, self.ghead = self.pipe.step(self.maps_parts)       # per-micro-batch losses (weights b_i / B)
            return self.loss
        self.maps_in.grad = None
        self.head.weight.grad = None
        pooled = ops.pool_patch(self.maps_in)                                        # (BN, T, Ce)
        f = ops.head_linear(pooled, self.head.weight).view(c["B"], c["N"], c["T"], c["D"])
        q, loss, xent, acc = ops.walk(f, c["tau"], c["p"], rng="device", rng_state=self.rng_state)
        loss.backward(self.ones)
        ops.join_side_streams()                       # the head's weight gradient ran beside the pooling backward
        self.loss = loss
        return loss

    def head_grad(self):
        return self.ghead if self.pipe is not None else self.head.weight.grad

    def loss_value(self):
        """float: the step's loss (mean over all clips)."""
        if self.pipe is not None:
            return float(self.pipe.loss())
        return float(self.loss.detach())

    def prepare(self):
        if not self.use_graph:
            for _ in range(3):
                self._step_eager()
            torch.cuda.synchronize()
            return
        # whole-step capture (forward + backward): warm up on a side stream FIRST so that the autograd accumulators
        # of the leaves are bound to a capturable stream, then capture
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                self._step_eager()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        # Capture, then check the instantiated graph: with several concurrent branches the same capture occasionally comes up in a
        # slow mode (measured: ~0.35 instead of 0.25 ms per replay for one instantiation in five with 5 micro-batches on 108 SMs,
        # profiles/r02_split_sweep.jsonl) - the branch-to-hardware-queue assignment differs between instantiations.  Keep the best
        # of up to four instantiations (all outside the timed region).
        best, seen = None, []
        for attempt in range(4 if self.pipe is not None else 1):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss = self._step_eager()
            torch.cuda.synchronize()
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            seen.append(ms)
            if best is None or ms < best[0]:
                best = (ms, g, loss)
            two = sorted(seen)[:2]
            if len(two) == 2 and two[1] <= 1.05 * two[0]:
                break                                   # two instantiations agree on the fast mode
        self.capture_trials = seen
        self.graph, self._static_loss = best[1], best[2]
        self.capture_ms = best[0]
        torch.cuda.synchronize()

    def step(self):
        if self.graph is not None:
            self.graph.replay()
            return self._static_loss
        return self._step_eager()


def time_gpu_steps(fn, steps, warmup, world, after=None, finish=None):
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
        if after:
            after()
    if finish:
        finish()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
        if after:
            after()
    if finish:
        finish()                     # joins the last exchange: every all-reduce is inside the timed region
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    return ms


def kernel_roofline(dev, clips=None, sms=0, sms_bwd=0):
    """The dominant kernels of the step are the two HBM streams over the 514 MB of maps, launched per micro-batch of `clips`
    clips on `sms` SMs (0 = all) exactly as the step launches them.  Timed live, alone, with CUDA events on the launching stream
    (>= 50 ms each), cycling over as many distinct micro-batch buffers as the step has (together > L2, so every launch finds its
    data in HBM, as in the step); achieved = algorithmic bytes of ONE launch / mean launch time."""
    from sapienza_video_contrastive_b200 import _lib
    c = CFG
    L = _lib.lib()
    clips = clips or c["B"]
    rows, hw = clips * c["N"] * c["T"] * c["Ce"], c["H"] * c["W"]
    nbuf = max(2, -(-c["B"] // clips))
    maps = [torch.randn(rows, hw, device=dev) for _ in range(nbuf)]
    gms = [torch.empty(rows, hw, device=dev) for _ in range(nbuf)]
    pooled = torch.randn(rows, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    i = [0]

    def fwd():
        i[0] = (i[0] + 1) % nbuf
        L.crw_pool_patch_fwd_sm(maps[i[0]].data_ptr(), pooled.data_ptr(), rows, hw, sms, st)

    def bwd():
        i[0] = (i[0] + 1) % nbuf
        L.crw_pool_patch_bwd_sm(pooled.data_ptr(), gms[i[0]].data_ptr(), rows, hw, sms_bwd, st)

    for name, fn in (("pool_patch_fwd", fwd), ("pool_patch_bwd", bwd)):
        for _ in range(2 * nbuf):
            fn()
        torch.cuda.synchronize()
        # the launches are replayed from a CUDA graph, as in the step (no host launch gaps between 20-us kernels)
        per_graph = 4 * nbuf
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            st_side = side.cuda_stream
            if name == "pool_patch_fwd":
                launch = lambda j: L.crw_pool_patch_fwd_sm(maps[j % nbuf].data_ptr(), pooled.data_ptr(), rows, hw, sms, st_side)
            else:
                launch = lambda j: L.crw_pool_patch_bwd_sm(pooled.data_ptr(), gms[j % nbuf].data_ptr(), rows, hw, sms_bwd, st_side)
            launch(0)
            torch.cuda.synchronize()
            with torch.cuda.graph(graph, stream=side):
                st_side = torch.cuda.current_stream().cuda_stream
                for j in range(per_graph):
                    launch(j)
        torch.cuda.synchronize()
        graph.replay()
        torch.cuda.synchronize()
        n, total = 0, 0.0
        while total < MIN_MS:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            total += e0.elapsed_time(e1)
            n += per_graph
        ms = total / n
        bytes_alg = rows * hw * 4 + rows * 4
        res[name] = {"us": ms * 1e3, "launches_timed": n, "clips_per_launch": clips,
                     "sms": (sms if name == "pool_patch_fwd" else sms_bwd) or 148, "buffers": nbuf,
                     "algorithmic_bytes": bytes_alg, "gbs": bytes_alg / ms / 1e6}
    return res


def ncu_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu capture of THIS round (profiles/ncu_summary.json; offline - a
    number taken under the profiler is never a bench value, so it is only quoted next to the live measurement)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            d = json.load(f)
        ent = d.get(kernel, {})
        return ent.get("dram_bytes_per_launch"), d.get("_source", "profiles/ncu_summary.json")
    except Exception:
        return None, None


# ---- M2: label propagation ----------------------------------------------------------------------------------------------
def lp_inputs(replicate_first, n_tgt=None, seed=0):
    c = LP
    n_tgt = n_tgt or c["n_tgt"]
    g = torch.Generator().manual_seed(seed)
    feats = torch.nn.functional.normalize(torch.randn(1, c["C"], c["n_ctx"] + n_tgt, c["h"], c["w"], generator=g), dim=1)
    if replicate_first:
        feats[:, :, : c["n_ctx"] + 1] = feats[:, :, :1]                     # vos.py:148-149 replicates frame 0 videoLen times
    lbls = torch.zeros(c["n_ctx"] + n_tgt, c["h"], c["w"], c["L"])
    lbls[: c["n_ctx"] + 1] = torch.nn.functional.one_hot(torch.randint(0, c["L"], (c["h"], c["w"]), generator=g), c["L"]).float()
    return feats, lbls


def lp_issued_flops(stats):
    """Tensor-core work the label-propagation kernels ISSUE per call (to set beside the algorithmic 2*C*pairs)."""
    c = LP
    R = 11                                                                    # radius 12: offsets up to 11
    per_tile = 0
    for ty in range((c["h"] + 15) // 16):
        rows = min(ty * 16 + 15 + R, c["h"] - 1) - max(ty * 16 - R, 0) + 1
        per_tile_keys = c["n_ctx"] * rows * 32 + ((c["h"] * c["w"] + 31) // 32) * 32
        per_tile += ((c["w"] + 7) // 8) * 128 * per_tile_keys * c["C"] * 2
    pre = per_tile * c["n_tgt"]
    tiles = max(stats.get("tiles", 1), 1)
    redo = stats.get("listed_tiles", 0) if stats.get("uncertified_queries", 0) > 1024 else 0      # (else the open queries are settled on the fp32 pipe)
    return pre + 3 * pre * redo / tiles


def label_prop_bench(dev, cpu=True, gpu_baseline=True):
    """BASELINE configs[3]: DAVIS-480p-shaped label propagation (C=256, 60x107, 20 context frames + long memory, radius 12,
    top-k 10).  Device-resident value, host-buffer e2e, roofline, the reference on the host cores and on this GPU."""
    from sapienza_video_contrastive_b200 import LabelPropagator
    c = LP
    hbm, tf, src = peaks()
    useful_flops = 2 * c["C"] * 90_214_480           # in-radius + long-memory score pairs per target frame
    lp = LabelPropagator(c["n_ctx"], [0], c["radius"], c["k"], c["tau"], normalize=True)
    out = {"metric": "label_prop_frames_per_s", "unit": "frames/s",
           "config": "BASELINE configs[3]: C=%d %dx%d, %d context + long-mem [0], radius %d, top-k %d, %d target frames per call "
                     "(layout + fp16 split + tcgen05 pre-ranking + exact fp32 re-ranking + certification + gathers)"
                     % (c["C"], c["h"], c["w"], c["n_ctx"], c["radius"], c["k"], c["n_tgt"])}
    variants = {}
    with ClockSampler(dev.index or 0) as clk:
        for name, rep in (("distinct_frames", False), ("replicated_first_frame", True)):
            feats, lbls = lp_inputs(rep)
            fd, ld = feats.to(dev), lbls.to(dev)
            med, n, total = timed_median(lambda: lp(fd, ld))
            st = dict(lp.stats)
            variants[name] = {"frames_per_s": c["n_tgt"] / med * 1e3, "ms_per_call": med, "calls_timed": n, "ms_timed": total,
                              "certification": st}
            if not rep:
                main_med, main_stats = med, st
                # ---- e2e: host feats / labels in (pinned), label maps out (pinned), copies inside the timed region ----
                hf, hl = feats.pin_memory(), lbls.pin_memory()
                hout = torch.empty(c["n_tgt"], c["h"], c["w"], c["L"]).pin_memory()

                def e2e_call():
                    preds, _ = lp(hf.to(dev, non_blocking=True), hl.to(dev, non_blocking=True))
                    hout.copy_(preds, non_blocking=True)

                emed, en, etot = timed_median(e2e_call)
                out["e2e"] = {"value": c["n_tgt"] / emed * 1e3, "unit": "frames/s", "ms_per_call": emed, "calls_timed": en,
                              "h2d_bytes_per_call": hf.numel() * 4 + hl.numel() * 4, "d2h_bytes_per_call": hout.numel() * 4,
                              "note": "pinned host encoder features + labels -> device -> LabelPropagator -> soft label maps back to host"}
                # the step right after it in test.py: full-resolution hard label images of every target frame
                preds, _ = lp(fd, ld)
                pal = torch.randint(0, 256, (c["L"], 3))
                pmed, pn, _ = timed_median(lambda: lp.label_images(preds, pal, (c["h"] * 8, c["w"] * 8)))
                out["postprocess"] = {"what": "upsample x8 (cv2 bilinear rule) + arg-max + colour table for the %d frames, one launch" % c["n_tgt"],
                                      "ms": pmed, "frames_per_s": c["n_tgt"] / pmed * 1e3}
    fps = c["n_tgt"] / main_med * 1e3
    out.update(value=fps, ms_per_frame=main_med / c["n_tgt"], clocks=clk.summary(), variants=variants)
    issued = lp_issued_flops(main_stats)
    out["roofline"] = {"bound": "tensor", "achieved": useful_flops * fps / 1e12, "peak": tf, "unit": "TFLOP/s",
                       "frac": useful_flops * fps / 1e12 / tf, "peak_source": src,
                       "issued_over_algorithmic": issued / (useful_flops * c["n_tgt"]),
                       "hbm_bound_frames_per_s": hbm * 1e9 / 144.6e6,
                       "note": "algorithmic flops = 2*C*(in-radius + long-memory pairs) per frame (SURVEY 8d: 46.2 GFLOP); the "
                               "kernel issues whole 16x8 query windows once on the fp16 hi planes, plus 3 MMAs per step on the listed tiles"}
    if cpu:
        from oracle import ref_harness as RH
        cores = host_threads()
        n_cpu = 2
        feats, lbls = lp_inputs(False, n_tgt=n_cpu)
        feats = torch.nn.functional.normalize(feats, dim=1)
        if RH.available():
            _, _, _, dt = RH.lp_video(feats, lbls, c["n_ctx"], [0], c["radius"], c["k"], c["tau"], "cpu")
            kind = "reference"
        else:
            from oracle import crw_oracle as O
            t0 = time.perf_counter()
            ki = O.context_index_bank(c["n_ctx"], [0], n_cpu)
            Wo, Io = O.lp_topk(feats[0].flatten(-2), ki, c["n_ctx"], 1, c["h"], c["w"], c["radius"], c["tau"], c["k"])
            O.lp_propagate(lbls, ki, Wo, Io, c["n_ctx"])
            dt, kind = time.perf_counter() - t0, "port"
        out["cpu_baseline"] = {"value": n_cpu / dt, "unit": "frames/s", "cores": cores, "kind": kind,
                               "sample": "%d target frames of the same shape through code/test.py's loop (key bank, radius mask, "
                                         "mem_efficient_batched_affinity, label gather), %.1f s" % (n_cpu, dt)}
    if gpu_baseline:
        try:
            from oracle import ref_harness as RH
            n_g = 8
            feats, lbls = lp_inputs(False, n_tgt=n_g)
            feats = torch.nn.functional.normalize(feats, dim=1)
            if RH.available():
                RH.lp_video(feats[:, :, : c["n_ctx"] + 2], lbls[: c["n_ctx"] + 2], c["n_ctx"], [0], c["radius"], c["k"], c["tau"], dev)   # warm-up
                _, _, _, dt = RH.lp_video(feats, lbls, c["n_ctx"], [0], c["radius"], c["k"], c["tau"], dev)
                out["torch_gpu_baseline"] = {"value": n_g / dt, "unit": "frames/s", "kind": "reference",
                                             "sample": "the unmodified code/test.py loop with args.device = this B200 (features start on the host as in "
                                                       "the reference, chunks are moved per 2 frames), %d target frames, %.2f s wall" % (n_g, dt)}
        except Exception as e:
            out["torch_gpu_baseline"] = {"error": repr(e)[:300]}
    return out


# ---- configs[2]: superpixel graph ------------------------------------------------------------------------------------------
def superpixel_bench(dev):
    """BASELINE configs[2] shape (8 clips x 8 frames, 196 superpixels on 256x256, 512-channel 32x32 maps): the whole step
    (segment-mean pooling -> head -> walk fwd+bwd -> head bwd -> pooling bwd) and the pooling kernels alone, plain and with the
    reference's default --dilate-superpixels element (51x51 'L1')."""
    from sapienza_video_contrastive_b200 import ops
    B, T, C, SPn, size = SP["B"], SP["T"], SP["C"], SP["SP"], SP["size"]
    hbm, _, _ = peaks()
    g = torch.Generator(device=dev).manual_seed(0)
    pts = torch.rand(B, T, SPn, 2, generator=g, device=dev) * size
    # label ids in the scan order of their segments, as the reference's SLIC labels are (enforce_connectivity numbers the segments in
    # scan order): sort the Voronoi seeds by (row band, x)
    band = (pts[..., 0] / (size / SPn ** 0.5)).floor()
    order = (band * size + pts[..., 1]).argsort(-1)
    pts = torch.gather(pts, 2, order[..., None].expand(-1, -1, -1, 2))
    yx = torch.stack(torch.meshgrid(torch.arange(size, device=dev), torch.arange(size, device=dev), indexing="ij"), -1).float()
    lab = torch.stack([torch.cdist(yx.reshape(1, -1, 2).expand(T, -1, -1), pts[b]).argmin(-1) for b in range(B)]).reshape(B, T, size, size)
    maps = torch.randn(B, C, T, 32, 32, generator=g, device=dev, requires_grad=True)
    gout = torch.randn(B, SPn, T, C, generator=g, device=dev)
    torch.manual_seed(0)
    w = torch.nn.Linear(C, 128, bias=False).to(dev).weight
    ones = torch.ones(1, device=dev)

    def pool_step(fn):
        fn().backward(gout)
        maps.grad = None

    def full_step():
        maps.grad = None
        w.grad = None
        f = ops.head_linear(ops.segment_mean(maps, lab, SPn), w)
        q, loss, xent, acc = ops.walk(f, 0.07, 0.1, rng="philox")
        loss.backward(ones)

    def graphed(fn):
        """GPU time of fn() replayed from a CUDA graph (no host launch gaps between its kernels); None if it cannot be captured."""
        try:
            fn()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                keep = fn()
            ms, _, _ = timed_median(gr.replay)
            del keep
            return ms
        except Exception:
            torch.cuda.synchronize()
            return None

    step_ms, n, _ = timed_median(full_step)
    fwd_ms, _, _ = timed_median(lambda: ops.segment_mean(maps.detach(), lab, SPn))
    fwd_graph = graphed(lambda: ops.segment_mean(maps.detach(), lab, SPn))
    dil_graph = graphed(lambda: ops.segment_mean_dilated(maps.detach(), lab, SPn, 51, "L1"))
    plain, _, _ = timed_median(lambda: pool_step(lambda: ops.segment_mean(maps, lab, SPn)))
    dil, _, _ = timed_median(lambda: pool_step(lambda: ops.segment_mean_dilated(maps, lab, SPn, 51, "L1")))
    dil_fwd, _, _ = timed_median(lambda: ops.segment_mean_dilated(maps.detach(), lab, SPn, 51, "L1"))
    bytes_alg = maps.numel() * 4 + lab.numel() * 8
    best = fwd_graph if fwd_graph is not None else fwd_ms
    return {"config": "BASELINE configs[2]: B=%d T=%d C=%d SP=%d labels %dx%d (ids in scan order, as SLIC's) maps 32x32" % (B, T, C, SPn, size, size),
            "step_ms": step_ms, "step_clips_per_s": B / step_ms * 1e3, "steps_timed": n,
            "pool_fwd_ms": best, "pool_fwd_eager_ms": fwd_ms, "pool_fwd_gbs": bytes_alg / best / 1e6, "pool_fwd_hbm_frac": bytes_alg / best / 1e6 / hbm,
            "pool_fwd_timing": "CUDA-graph replay of the op (label lists + tensor-core pooling)" if fwd_graph is not None else "eager call",
            "dilated_L1_51_fwd_graph_ms": dil_graph,
            "pool_fwd_bwd_ms": plain, "pool_fwd_bwd_clips_per_s": B / plain * 1e3,
            "dilated_L1_51_fwd_ms": dil_fwd, "dilated_L1_51_fwd_bwd_ms": dil, "dilated_clips_per_s": B / dil * 1e3}


# ---- the data-loader producers (patch grid, superpixel label maps) ---------------------------------------
def producers_bench(dev, cpu=True):
    """The inputs of the two walk variants for one training batch of CFG's shape (B clips x T frames of 256 x 256): the patch grid
    (augs.py:59-82: 49 crops + PIL bilinear resizes per frame) and the SLIC label maps (data/superpixels.py:9-63, --num-sp 30
    --compactness 200 defaults), produced on the GPU; beside them the reference's per-frame CPU procedure on one host core (a
    DataLoader worker), timed on a few frames."""
    from sapienza_video_contrastive_b200 import augs, superpixels
    hbm, _, _ = peaks()
    F, size = CFG["B"] * CFG["T"], 256
    g = torch.Generator(device="cpu").manual_seed(0)
    frames = torch.randint(0, 256, (F, size, size, 3), dtype=torch.uint8, generator=g)
    torch.manual_seed(0)
    P = ((size - 64) // 32 + 1) ** 2
    boxes = augs.draw_patch_boxes(F, P, 64).to(dev)
    fr_d = frames.to(dev)
    pg_ms, n_pg, _ = timed_median(lambda: augs.patch_grid_frames(fr_d, boxes, 64, 32, 64))
    pg_bytes = F * P * 3 * 64 * 64 * 4 + frames.numel()
    vid = torch.nn.functional.avg_pool2d(torch.randn(F, 3, size, size, generator=g), 9, 1, 4).to(dev)
    sl_ms, n_sl, _ = timed_median(lambda: superpixels.slic_frames(vid, 30, 200.0), min_iters=3)
    sl_nc_ms, _, _ = timed_median(lambda: superpixels.slic_frames(vid, 30, 200.0, enforce_connectivity=False), min_iters=3)
    out = {"config": "%d frames of %dx%d (B=%d clips x T=%d)" % (F, size, size, CFG["B"], CFG["T"]),
           "patch_grid": {"ms": pg_ms, "frames_per_s": F / pg_ms * 1e3, "gbs": pg_bytes / pg_ms / 1e6, "hbm_frac": pg_bytes / pg_ms / 1e6 / hbm,
                          "windows_per_frame": P, "parity": "bit-exact with Pillow (tests)"},
           "slic": {"ms": sl_ms, "frames_per_s": F / sl_ms * 1e3, "clustering_only_ms": sl_nc_ms, "n_segments": 30, "compactness": 200.0,
                    "parity": "unpinned (scikit-image absent): bit-exact with oracle/slic_oracle.py's restatement"}}
    if cpu:
        from PIL import Image
        import numpy as np
        nf = 4
        bx = boxes[:nf].cpu().numpy()
        mean, std = np.array(augs.IMG_MEAN, dtype=np.float32), np.array(augs.IMG_STD, dtype=np.float32)
        t0 = time.perf_counter()
        for f in range(nf):                                       # augs.py:70-80, window by window
            x = frames[f].numpy()
            for p in range(P):
                wy, wx = (p // 7) * 32, (p % 7) * 32
                i, j, h, w = (int(v) for v in bx[f, p])
                im = Image.fromarray(x[wy:wy + 64, wx:wx + 64]).crop((j, i, j + w, i + h)).resize((64, 64), Image.BILINEAR)
                _ = ((np.asarray(im, dtype=np.float32) / 255.0 - mean) / std).transpose(2, 0, 1)
        cpu_pg = nf / (time.perf_counter() - t0)
        out["patch_grid"]["cpu_baseline"] = {"value": cpu_pg, "unit": "frames/s", "cores": 1, "kind": "port",
                                             "sample": "%d frames, the reference's per-window PIL crop + resize + normalise loop" % nf}
        from oracle import slic_oracle
        t0 = time.perf_counter()
        slic_oracle.slic_labels(np.moveaxis(vid[0].cpu().numpy(), 0, -1), 30, 200.0)
        out["slic"]["cpu_baseline"] = {"value": 1.0 / (time.perf_counter() - t0), "unit": "frames/s", "cores": 1, "kind": "port",
                                       "sample": "1 frame through oracle/slic_oracle.py (numpy + a Python flood fill; scikit-image's Cython "
                                                 "is not in this image and would be faster)"}
    return out


# ---- configs[4]: sweep points ------------------------------------------------------------------------------------------------
def sweep_bench(dev):
    from sapienza_video_contrastive_b200 import ops
    _, tf, _ = peaks()
    res = []
    for pt in SWEEP:
        B, N, T, D = pt["B"], pt["N"], pt["T"], 128
        try:
            f = torch.randn(B, N, T, D, device=dev, requires_grad=True)
            ones = torch.ones(1, device=dev)

            def step():
                f.grad = None
                q, loss, xent, acc = ops.walk(f, 0.07, 0.1, rng="philox")
                loss.backward(ones)

            ms, n, total = timed_median(step, warm=2, min_iters=3)
            flops = 3 * (2 * (T - 1) * N * N * D + 2 * N ** 3 * 3 * (T - 2)) * B          # fwd + 2x bwd, minimal chain count
            res.append({"B": B, "N": N, "T": T, "ms": ms, "iters_timed": n, "clips_per_s": B / ms * 1e3,
                        "algorithmic_tflops": flops / ms / 1e9, "tensor_frac": flops / ms / 1e9 / tf})
            del f
            torch.cuda.empty_cache()
        except Exception as e:
            res.append({"B": B, "N": N, "T": T, "error": repr(e)[:200]})
    return res


# ---- the reference's own PyTorch code on this GPU ("the bar is the stock PyTorch/cuBLAS sequence on the same B200") ----
def torch_gpu_walk_baseline(dev):
    c = CFG
    step, kind = reference_walk_step(c["B"], dev)
    i = [0]

    def fn():
        step(i[0])
        i[0] += 1

    med, n, total = timed_median(fn)
    ours, others, _ = count_kernels(fn)
    return {"value": c["B"] / med * 1e3, "unit": "clips/s", "ms_per_step": med, "steps_timed": n, "kind": kind,
            "kernel_launches_per_step": (ours or 0) + (others or 0),
            "what": "the reference's CRW.forward + backward (eager PyTorch: ATen / cuBLAS kernels) on the same maps, device-resident"}


# ---- module-level end to end: CRW(args)(x) with the stock ResNet-18 -------------------------------------------------------------
def module_e2e(dev, world, steps=10):
    """SURVEY 8d config 2 'end-to-end variant' (code/train.py:58-81): x (20,4,147,64,64) per GPU from pinned host memory ->
    CRW(args)(x) -> loss.backward(), stock ResNet-18 on cuDNN, DistributedDataParallel when world > 1; the reference module (or the
    port) on the same GPU beside it (one GPU only)."""
    import argparse as ap
    from sapienza_video_contrastive_b200 import CRW
    c = CFG
    ns = ap.Namespace(device=str(dev), dropout=c["p"], featdrop=0.0, temp=c["tau"], head_depth=0, model_type="scratch",
                      remove_layers=[], dilate_superpixels=False, flip=False, sk_targets=False)
    torch.manual_seed(0)
    crw = CRW(ns).to(dev)
    model = crw
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(crw, device_ids=[dev.index])
    hx = [torch.randn(c["B"], c["T"], c["N"] * 3, 64, 64).pin_memory() for _ in range(2)]
    loss_host = torch.zeros(1).pin_memory()

    def run(m, i):
        x = hx[i & 1].to(dev, non_blocking=True)
        for p_ in m.parameters():
            p_.grad = None
        q, loss, diags = m(x, None, None)
        loss.mean().backward()
        loss_host.copy_(loss.detach(), non_blocking=True)

    k = [0]

    def ours():
        run(model, k[0])
        k[0] += 1

    med, n, _ = timed_median(ours, min_ms=200.0, warm=2, min_iters=3)
    out = {"value": c["B"] * world / med * 1e3, "unit": "clips/s", "ms_per_step": med, "steps_timed": n,
           "h2d_bytes_per_step": hx[0].numel() * 4, "d2h_bytes_per_step": 4,
           "what": "CRW(args)(x) + backward, x (20,4,147,64,64) fp32 per GPU from pinned host memory, stock ResNet-18 (cuDNN) + the "
                   "sm_100a hot path" + (", DistributedDataParallel over %d GPUs" % world if world > 1 else "")}
    if world == 1:
        try:
            from oracle import ref_import
            if ref_import.available():
                import contextlib
                import io
                ref_model, _, _ = ref_import.load()
                with contextlib.redirect_stdout(io.StringIO()):
                    torch.manual_seed(0)
                    ref = ref_model.CRW(ref_import.namespace(device=str(dev), dropout=c["p"], temp=c["tau"])).to(dev)
                ref.load_state_dict(crw.state_dict())
                j = [0]

                def theirs():
                    run(ref, j[0])
                    j[0] += 1

                rmed, rn, _ = timed_median(theirs, min_ms=200.0, warm=2, min_iters=3)
                out["reference_module_same_gpu"] = {"value": c["B"] / rmed * 1e3, "unit": "clips/s", "ms_per_step": rmed, "steps_timed": rn,
                                                    "what": "the unmodified reference CRW (code/model.py) with the same weights on the same B200"}
        except Exception as e:
            out["reference_module_same_gpu"] = {"error": repr(e)[:300]}
    return out


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    from sapienza_video_contrastive_b200 import _lib, ops
    _lib.lib()
    ops.check_device(dev)
    c = CFG
    ops.set_async_wgrad(True)
    sizes = [int(x) for x in args.part_sizes.split(",")] if args.part_sizes else None
    if sizes is None and args.parts == 1:
        args.pool_sms = args.pool_sms_bwd = 0
    hp = HotPath(dev, rank, use_graph=not args.eager, parts=args.parts, pool_sms=args.pool_sms, sizes=sizes, head_splits=args.head_splits,
                 pool_sms_bwd=args.pool_sms_bwd)
    hp.prepare()
    grads = torch.zeros(RESNET18_GRAD_FLOATS, device=dev) if world > 1 else None
    pending = []

    def finish():
        while pending:
            pending.pop().wait()                 # the compute stream waits for the exchange in flight (no host block)

    def after():
        # The data-parallel exchange of a ResNet-18-sized gradient buffer; the head's real gradient rides in
        # it.  In training this all-reduce hides behind the encoder backward, which is not part of this path: here it runs on
        # NCCL's stream beside the NEXT step's pooling / walk kernels and is joined before the following exchange is issued
        # (and before the timed region closes), so every exchange is paid for inside the measurement.
        finish()
        grads[: c["D"] * c["Ce"]].copy_(hp.head_grad().view(-1))
        pending.append(dist.all_reduce(grads, async_op=True))

    with ClockSampler(local_rank) as clk:
        ms = time_gpu_steps(hp.step, args.steps, args.warmup, world, after if world > 1 else None, finish if world > 1 else None)
        # the same step, one event pair per step, for >= 50 ms: the per-step median next to the contract's K-step mean
        med_ms, med_n, _ = timed_median(hp.step, warm=0) if world == 1 else (None, None, None)
    clocks = clk.summary()
    value = c["B"] * world * args.steps / ms * 1e3
    hp.step()
    loss_now = hp.loss_value()
    if not (1.0 < loss_now < 12.0):                                  # random init: near log(49) = 3.9
        raise SystemExit("bench.py: the timed step produced loss %r - not a valid run" % loss_now)

    # ---- end to end through the public operator API with HOST buffers (pinned), copies inside the timed region ----
    host = [torch.randn(hp.maps.shape, pin_memory=True) for _ in range(2)]
    dev_in = [torch.empty_like(hp.maps).requires_grad_(True) for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    loss_host = torch.zeros(8, pin_memory=True)
    ghead_host = torch.zeros(c["D"], c["Ce"], pin_memory=True)
    h2d = host[0].numel() * 4
    d2h = 4 + ghead_host.numel() * 4

    e2e = HotPath(dev, rank, use_graph=False, parts=args.parts, pool_sms=args.pool_sms, sizes=sizes, head_splits=args.head_splits,
                  pool_sms_bwd=args.pool_sms_bwd)   # same operators, eager, on uploaded inputs
    e2e.head = hp.head

    def e2e_loop(n):
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        done = [torch.cuda.Event(), torch.cuda.Event()]
        with torch.cuda.stream(copy_stream):
            dev_in[0].data.copy_(host[0], non_blocking=True)
            ready[0].record()
        for i in range(n):
            cur, nxt = i & 1, (i + 1) & 1
            if i + 1 < n:
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(done[nxt]) if i >= 1 else None
                    dev_in[nxt].data.copy_(host[nxt], non_blocking=True)
                    ready[nxt].record()
            torch.cuda.current_stream().wait_event(ready[cur])
            e2e.set_inputs(dev_in[cur])
            loss = e2e._step_eager()
            loss_host[: loss.numel()].copy_(loss.detach(), non_blocking=True)
            ghead_host.copy_(e2e.head_grad(), non_blocking=True)
            done[cur].record()

    e2e_loop(2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_e2e = max(6, min(args.steps, 10))                              # ~9 ms per step: >= 50 ms timed
    e0.record()
    e2e_loop(n_e2e)
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t)
    e2e_val = c["B"] * world * n_e2e / ms_e2e * 1e3

    mod = None
    if args.e2e_module:
        mod = module_e2e(dev, world)

    if rank != 0:
        return
    hbm, tf, src = peaks()
    kr = kernel_roofline(dev, clips=max(hp.sizes) if hp.pipe is not None else None, sms=args.pool_sms if hp.pipe is not None else 0,
                         sms_bwd=args.pool_sms_bwd if hp.pipe is not None else 0)
    dom = max(kr, key=lambda k: kr[k]["us"])
    traffic, traffic_src = ncu_traffic(dom)
    n_ours, n_other, knames = count_kernels(hp._step_eager)
    step_desc = ("pipeline.PatchWalkPipeline: micro-batches of %s clips on staggered streams, pooling confined to %d SMs" % (hp.sizes, args.pool_sms)
                 if hp.pipe is not None else "one chain of autograd operators")
    out = {"metric": "crw_walk_fwd_bwd_clips_per_s", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world),
           "launch": "eager" if args.eager else "cuda_graph", "parts": hp.sizes or [c["B"]], "pool_sms": args.pool_sms, "pool_sms_bwd": args.pool_sms_bwd, "schedule": step_desc, "graph_instantiations_ms": getattr(hp, "capture_trials", None),
           "clocks": clocks,
           "e2e": {"value": e2e_val, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps_timed": n_e2e,
                   "ms_timed": ms_e2e,
                   "note": "pinned host maps -> device (double-buffered copy stream) -> hot path -> loss + head grad back to host"},
           "gpu_launches": (n_ours or 0) * args.steps,
           "gpu_launches_note": "counted with the CUDA profiler on one step: %s kernels of libcrw_b200.so + %s others (torch elementwise) per step; "
                                "kernels: %s" % (n_ours, n_other, ", ".join(knames)),
           "roofline": {"bound": "hbm", "kernel": dom, "achieved": kr[dom]["gbs"], "peak": hbm, "unit": "GB/s",
                        "frac": kr[dom]["gbs"] / hbm, "traffic": traffic, "traffic_source": traffic_src, "peak_source": src,
                        "l2": "launches cycle over the step's micro-batch buffers (together 514 MB > L2)",
                        "step_hbm_frac": 2 * (c["B"] * c["N"] * c["T"] * c["Ce"] * (c["H"] * c["W"] + 1) * 4) / (ms / args.steps * 1e-3) / 1e9 / hbm},
           "kernels": kr, "loss": loss_now}
    if med_ms is not None:
        out["ms_per_step_median"] = med_ms
        out["value_median"] = c["B"] / med_ms * 1e3
        out["median_steps_timed"] = med_n
    if mod is not None:
        out["e2e_module"] = mod
    if world == 1 and not args.no_cpu:
        cores = host_threads()
        val, cms, kind = time_cpu_walk(c["B"], 20, 2)
        out["cpu_baseline"] = {"value": val, "unit": "clips/s", "cores": cores, "kind": kind,
                               "sample": "20 steps of the full %d-clip batch (%s), %.0f ms/step"
                                         % (c["B"], "unmodified reference CRW.forward + backward on the precomputed maps" if kind == "reference"
                                            else "oracle/crw_oracle.py port", cms)}
    if world == 1 and not args.no_extras:
        for key, fn in (("torch_gpu_baseline", lambda: torch_gpu_walk_baseline(dev)),
                        ("label_prop", lambda: label_prop_bench(dev, cpu=not args.no_cpu)),
                        ("superpixel", lambda: superpixel_bench(dev)),
                        ("sweep", lambda: sweep_bench(dev)),
                        ("producers", lambda: producers_bench(dev, cpu=not args.no_cpu)),
                        ("e2e_module", (lambda: module_e2e(dev, 1)) if mod is None else None)):
            if fn is None or (key == "label_prop" and args.no_lp):
                continue
            try:
                out[key] = fn()
            except Exception as e:                                   # the headline line must still print
                out[key] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=250)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--eager", action="store_true", help="launch the step op by op instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu", action="store_true", help="skip the host-core baselines")
    ap.add_argument("--no-lp", action="store_true", help="skip the label-propagation leg")
    ap.add_argument("--no-extras", action="store_true", help="headline line only (no label_prop / superpixel / sweep / baselines on the GPU)")
    ap.add_argument("--parts", type=int, default=1, help="with --part-sizes '': equal micro-batches per step (1 = one chain of autograd operators)")
    ap.add_argument("--head-splits", type=int, default=4, help="split-K slices of the head forward GEMM of a micro-batch")
    ap.add_argument("--part-sizes", default="4,4,4,4,4", help="clips per micro-batch on staggered streams (pipeline.PatchWalkPipeline), must add up to "
                                                           "20; '' = use --parts.  Default: the best of the measured sweep (profiles/r02_split_sweep.jsonl)")
    ap.add_argument("--pool-sms-bwd", type=int, default=0, help="SMs of the pooling BACKWARD kernels (store-bound: 0 = all SMs measured best)")
    ap.add_argument("--pool-sms", type=int, default=100, help="SMs the pooling kernels are confined to while micro-batches overlap (0 = all)")
    ap.add_argument("--nccl-ctas", type=int, default=16, help="NCCL_MAX_CTAS for the gradient all-reduce (several GPUs)")
    ap.add_argument("--e2e-module", action="store_true", help="also time CRW(args)(x) with the ResNet-18 (DDP when several GPUs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU baseline)")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # The gradient exchange of step i runs beside step i+1, whose pooling kernels are HBM-bound and whose walk kernels need
        # whole SMs: cap NCCL's CTAs (2 GPUs, ms/step with 4 / 8 / 16 / 32 CTAs: 0.58 / 0.35 / 0.291 / 0.294 -
        # profiles/r02_multi_gpu.txt; below 16 the exchange itself outlasts the step).  An explicit NCCL_MAX_CTAS in the environment wins.
        os.environ.setdefault("NCCL_MAX_CTAS", str(args.nccl_ctas))
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
