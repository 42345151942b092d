#!/usr/bin/env python
"""bench.py - CRW walk fwd+bwd clips/s on the Kinetics-shaped batch (BASELINE.json configs[1]) and, as extra keys,
DAVIS-shaped label-propagation frames/s (configs[3]).

    python bench.py --gpus N --steps K --warmup W            # our arm (sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port) on host cores

One "step" = one pass of the hot path over one batch per GPU: patch mean-pool (a1) -> Linear head (stock cuBLAS) ->
fused walk forward+backward (a4-a6) -> head backward -> pool backward, on precomputed encoder maps
(B=20 clips/GPU, N=49, T=4, C_e=512, 8x8 maps, tau 0.07, edge dropout 0.1).  Prints ONE JSON line.
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
RESNET18_GRAD_FLOATS = 11_176_512 + 512 * 128          # encoder + head parameters (44.97 MB)


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


# --------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, timed on the host cores
# --------------------------------------------------------------------------------------------------------------
def cpu_step_fn(B):
    from oracle import crw_oracle as O
    c = CFG
    g = torch.Generator().manual_seed(1000)
    maps = torch.randn(B * c["N"], c["Ce"], c["T"], c["H"], c["W"], generator=g).requires_grad_(True)
    torch.manual_seed(0)
    head = torch.nn.Linear(c["Ce"], c["D"], bias=False)

    def step(i):
        torch.manual_seed(123 + i)
        q = O.patch_nodes(maps, head.weight, B)
        u12, u21p = O.draw_uniforms(B, c["N"], c["T"])
        loss, *_ = O.walk_loss(q, c["tau"], c["p"], u12, u21p)
        maps.grad = None
        head.weight.grad = None
        loss.mean().backward()
        return float(loss)

    return step


def time_cpu(B, steps, warmup):
    step = cpu_step_fn(B)
    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        step(warmup + i)
    dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps * 1e3


def run_reference(args, rank, world):
    if rank != 0:
        return
    c = CFG
    cores = torch.get_num_threads()
    steps, warmup = min(args.steps, 40), min(max(args.warmup, 1), 3)     # bounded: ~0.25 s per 20-clip step on 8 cores
    val, ms = time_cpu(c["B"], steps, warmup)
    out = {"metric": "crw_walk_fwd_bwd_clips_per_s", "value": val, "unit": "clips/s", "n_gpus": args.gpus, "steps": steps,
           "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "impl": "reference",
           "config": workload_config(1),
           "cpu_baseline": {"value": val, "unit": "clips/s", "cores": cores, "kind": "port",
                            "sample": "%d steps of the full %d-clip batch (oracle/crw_oracle.py: pool+head+walk fwd+bwd)" % (steps, c["B"])},
           "e2e": {"value": val, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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
    """The step of the headline metric on one GPU, eager or as a captured CUDA graph."""

    def __init__(self, dev, rank, use_graph=True):
        from sapienza_video_contrastive_b200 import ops
        self.ops, self.dev = ops, dev
        c = CFG
        g = torch.Generator(device="cpu").manual_seed(1000 + rank)
        # encoder-native physical layout (BN, T, C, H, W) (resnet.From3D), logically (BN, C, T, H, W)
        self.maps = torch.randn(c["B"] * c["N"], c["T"], c["Ce"], c["H"], c["W"], generator=g).to(dev)
        torch.manual_seed(0)
        self.head = torch.nn.Linear(c["Ce"], c["D"], bias=False).to(dev)
        self.rng_state = torch.tensor([123, 0], dtype=torch.int64, device=dev)
        self.ones = torch.ones(1, device=dev)
        self.maps_in = self.maps.clone().requires_grad_(True)
        self.graph = None
        self.use_graph = use_graph
        self.loss = torch.zeros(1, device=dev)
        self.gmaps = None

    def _step_eager(self):
        c = CFG
        ops = self.ops
        self.maps_in.grad = None
        self.head.weight.grad = None
        pooled = ops.pool_patch(self.maps_in)                                        # (BN, T, Ce)
        f = ops.head_linear(pooled, self.head.weight).view(c["B"], c["N"], c["T"], c["D"])
        q, loss, xent, acc = ops.walk(f, c["tau"], c["p"], rng="device", rng_state=self.rng_state)
        loss.backward(self.ones)
        ops.join_side_streams()                       # the head's weight gradient ran beside the pooling backward
        self.loss = loss
        return loss

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
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._static_loss = self._step_eager()
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


def kernel_roofline(dev):
    """The dominant kernels of the step are the two HBM streams over the 514 MB of maps.  Timed live with CUDA events
    on the launching stream; achieved = algorithmic bytes / time."""
    from sapienza_video_contrastive_b200 import _lib
    c = CFG
    L = _lib.lib()
    rows, hw = c["B"] * c["N"] * c["T"] * c["Ce"], c["H"] * c["W"]
    maps = torch.randn(rows, hw, device=dev)
    pooled = torch.empty(rows, device=dev)
    gm = torch.empty(rows, hw, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    for name, fn in (("pool_patch_fwd", lambda: L.crw_pool_patch_fwd(maps.data_ptr(), pooled.data_ptr(), rows, hw, st)),
                     ("pool_patch_bwd", lambda: L.crw_pool_patch_bwd(pooled.data_ptr(), gm.data_ptr(), rows, hw, st))):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 30
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        bytes_alg = rows * hw * 4 + rows * 4
        res[name] = {"us": us, "algorithmic_bytes": bytes_alg, "gbs": bytes_alg / us / 1e3}
    return res


def label_prop_bench(dev):
    """BASELINE configs[3]: DAVIS-480p-shaped label propagation, features resident, top-k for n_tgt target frames."""
    from sapienza_video_contrastive_b200 import LabelPropagator
    c = LP
    g = torch.Generator().manual_seed(0)
    feats = torch.randn(1, c["C"], c["n_ctx"] + c["n_tgt"], c["h"], c["w"], generator=g).to(dev)
    lbls = torch.zeros(c["n_ctx"] + c["n_tgt"], c["h"], c["w"], c["L"])
    lbls[: c["n_ctx"] + 1, :, :, 0] = 1
    lp = LabelPropagator(c["n_ctx"], [0], c["radius"], c["k"], c["tau"], normalize=True)
    lp(feats, lbls)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        lp(feats, lbls)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fps = c["n_tgt"] / ms * 1e3
    # the step right after it in test.py: full-resolution hard label images of every target frame
    preds, _ = lp(feats, lbls)
    pal = torch.randint(0, 256, (c["L"], 3), generator=g)
    lp.label_images(preds, pal, (c["h"] * 8, c["w"] * 8))
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(10):
        lp.label_images(preds, pal, (c["h"] * 8, c["w"] * 8))
    p1.record()
    torch.cuda.synchronize()
    post_ms = p0.elapsed_time(p1) / 10
    useful_flops = 2 * c["C"] * 90_214_480           # in-radius + long-memory score pairs per target frame
    _, tf, _ = peaks()
    return {"metric": "label_prop_frames_per_s", "value": fps, "unit": "frames/s", "ms_per_frame": ms / c["n_tgt"],
            "postprocess": {"what": "upsample x8 (cv2 bilinear rule) + arg-max + colour table for the %d frames, one launch" % c["n_tgt"],
                            "ms": post_ms, "frames_per_s": c["n_tgt"] / post_ms * 1e3},
            "config": "C=%d %dx%d, %d context + long-mem [0], radius %d, top-k %d, %d target frames per call (layout + hi/lo split + tcgen05 top-k + gathers)"
                      % (c["C"], c["h"], c["w"], c["n_ctx"], c["radius"], c["k"], c["n_tgt"]),
            "roofline": {"bound": "tensor", "achieved": useful_flops * fps / 1e12, "peak": tf, "unit": "TFLOP/s",
                         "frac": useful_flops * fps / 1e12 / tf,
                         "note": "algorithmic flops = 2*C*(in-radius + long-memory pairs) per frame; the fp32-faithful fp16 hi/lo "
                                 "split issues 3 MMAs per product and whole 16x8 query windows, so the tensor pipe does ~8x this"}}


def superpixel_pool_bench(dev):
    """BASELINE configs[2] shape (8 clips x 8 frames, 196 superpixels on 256x256, 512-channel 32x32 maps): superpixel pooling
    forward + backward, plain and with the reference's default --dilate-superpixels element (51x51 'L1')."""
    from sapienza_video_contrastive_b200 import ops
    B, T, C, SP, size = 8, 8, 512, 196, 256
    g = torch.Generator(device=dev).manual_seed(0)
    pts = torch.rand(B, T, SP, 2, generator=g, device=dev) * size
    yx = torch.stack(torch.meshgrid(torch.arange(size, device=dev), torch.arange(size, device=dev), indexing="ij"), -1).float()
    lab = torch.stack([torch.cdist(yx.reshape(1, -1, 2).expand(T, -1, -1), pts[b]).argmin(-1) for b in range(B)]).reshape(B, T, size, size)
    maps = torch.randn(B, C, T, 32, 32, generator=g, device=dev, requires_grad=True)
    gout = torch.randn(B, SP, T, C, generator=g, device=dev)

    def timed(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    def step(fn):
        fn().backward(gout)
        maps.grad = None

    plain = timed(lambda: step(lambda: ops.segment_mean(maps, lab, SP)))
    dil = timed(lambda: step(lambda: ops.segment_mean_dilated(maps, lab, SP, 51, "L1")))
    return {"config": "B=%d T=%d C=%d SP=%d labels %dx%d maps 32x32, forward + backward" % (B, T, C, SP, size, size),
            "plain_ms": plain, "plain_clips_per_s": B / plain * 1e3,
            "dilated_L1_51_ms": dil, "dilated_clips_per_s": B / dil * 1e3}


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    from sapienza_video_contrastive_b200 import _lib, ops
    _lib.lib()
    ops.check_device(dev)
    c = CFG
    ops.set_async_wgrad(True)
    hp = HotPath(dev, rank, use_graph=not args.eager)
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
        grads[: c["D"] * c["Ce"]].copy_(hp.head.weight.grad.view(-1))
        pending.append(dist.all_reduce(grads, async_op=True))

    with ClockSampler(local_rank) as clk:
        ms = time_gpu_steps(hp.step, args.steps, args.warmup, world, after if world > 1 else None, finish if world > 1 else None)
    clocks = clk.summary()
    value = c["B"] * world * args.steps / ms * 1e3

    # ---- end to end through the public operator API with HOST buffers (pinned), copies inside the timed region ----
    host = [torch.randn(hp.maps.shape, pin_memory=True) for _ in range(2)]
    dev_in = [torch.empty_like(hp.maps).requires_grad_(True) for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    loss_host = torch.zeros(1, pin_memory=True)
    ghead_host = torch.zeros(c["D"], c["Ce"], pin_memory=True)
    h2d = host[0].numel() * 4
    d2h = 4 + ghead_host.numel() * 4

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
            x = dev_in[cur]
            x.grad = None
            hp.head.weight.grad = None
            pooled = ops.pool_patch(x)
            f = ops.head_linear(pooled, hp.head.weight).view(c["B"], c["N"], c["T"], c["D"])
            q, loss, xent, acc = ops.walk(f, c["tau"], c["p"], rng="device", rng_state=hp.rng_state)
            loss.backward(hp.ones)
            ops.join_side_streams()
            loss_host.copy_(loss.detach(), non_blocking=True)
            ghead_host.copy_(hp.head.weight.grad, non_blocking=True)
            done[cur].record()

    e2e_loop(2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_e2e = max(3, min(args.steps, 10))
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

    if rank != 0:
        return
    hbm, tf, src = peaks()
    kr = kernel_roofline(dev)
    dom = max(kr, key=lambda k: kr[k]["us"])
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            traffic = json.load(f).get(dom, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    out = {"metric": "crw_walk_fwd_bwd_clips_per_s", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world),
           "launch": "eager" if args.eager else "cuda_graph",
           "clocks": clocks,
           "e2e": {"value": e2e_val, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "note": "pinned host maps -> device (double-buffered copy stream) -> hot path -> loss + head grad back to host"},
           "gpu_launches": 9 * args.steps,
           "gpu_launches_note": "per step: crw pool_fwd, gemm_tf32 (head fwd), walk_pairs_fwd, walk_chain_cluster, walk_pairs_bwd, "
                                "gemm_tf32 (head dgrad), gemm_tf32 (split-K head wgrad), splitk_reduce, pool_bwd; "
                                "plus one torch elementwise scale of the walk gradient",
           "roofline": {"bound": "hbm", "kernel": dom, "achieved": kr[dom]["gbs"], "peak": hbm, "unit": "GB/s",
                        "frac": kr[dom]["gbs"] / hbm, "traffic": traffic, "peak_source": src,
                        "step_hbm_frac": (2 * kr[dom]["algorithmic_bytes"]) / (ms / args.steps * 1e-3) / 1e9 / hbm},
           "kernels": kr}
    if world == 1 and not args.no_cpu:
        cores = torch.get_num_threads()
        val, cms = time_cpu(c["B"], 20, 1)
        out["cpu_baseline"] = {"value": val, "unit": "clips/s", "cores": cores, "kind": "port",
                               "sample": "20 steps of the full %d-clip batch (oracle/crw_oracle.py), %.0f ms/step" % (c["B"], cms)}
    if world == 1 and not args.no_lp:
        try:
            out["label_prop"] = label_prop_bench(dev)
        except Exception as e:                                   # the headline line must still print
            out["label_prop"] = {"error": repr(e)}
        try:
            out["superpixel_pooling"] = superpixel_pool_bench(dev)
        except Exception as e:
            out["superpixel_pooling"] = {"error": repr(e)}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--eager", action="store_true", help="launch the step op by op instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-lp", action="store_true")
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
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
