#!/usr/bin/env python
"""bench.py -- headline benchmark of the correlation + flow_warp hot path (BASELINE.json).

Workload (config.workload): BASELINE.json configs[1], "PWC-Lite kitti_base training step,
synthetic 256x832 pairs, B=8, 1 B200" restricted to the path this repo owns: every Correlation
and flow_warp call of one training iteration, forward and backward (unsamflow_b200/hotpath.py:
10 corr fwd + 10 corr bwd + 16 warp fwd + 16 warp bwd), on seeded synthetic tensors.

  value      pairs/s with all inputs resident in HBM; a step = one CUDA graph (both flow directions on two
             streams inside it).  L2 is flushed between steps (outside the per-step event pair).  Device
             time from CUDA events, max over ranks.
  e2e        same step through the public tensor API with HOST inputs: H2D of every input from pinned
             memory and D2H of the step's results are inside the timed region.
  roofline   the dominant op's kernel: algorithmic bytes per launch / mean launch duration (external
             timing events recorded inside the graph around that kernel, on its stream, in every timed
             step) against MEASURED_PEAKS.json.
  cpu_baseline / --impl reference
             the reference's CPU path (oracle/torch_port.py: the same ATen calls the reference
             makes) on the host cores, on a bounded sample (1 pair per step).

Launch: python bench.py [--gpus N --steps K --warmup W]; for N > 1 under torchrun (one rank per GPU,
weak scaling: every rank runs its own B=8 batch, no data-path collective).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "corr+warp fwd+bwd hot path of a PWC-Lite kitti_base training step: pairs/s (GB/s vs HBM peak in roofline)"
UNIT = "pairs/s"
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # CUDA-core FFMA peak at max SM clock


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--height", type=int, default=256)
    ap.add_argument("--width", type=int, default=832)
    ap.add_argument("--flow-sigma", type=float, default=3.0)
    ap.add_argument("--e2e-steps", type=int, default=40)
    ap.add_argument("--cpu-reps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--op-table", default="", help="write the per-op timing table (JSON) here")
    ap.add_argument("--train-steps", type=int, default=10,
                    help="timed steps of the full PWC-Lite training step (0 = skip); reported as train_step")
    ap.add_argument("--batch-directions", action="store_true",
                    help="evaluate the two flow directions as one batch of 2B pairs (half the launches)")
    ap.add_argument("--streams", type=int, default=2, choices=[1, 2],
                    help="2: the two flow directions of the step run on two CUDA streams inside the graphs")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# reference CPU path (the ONLY place bench.py touches oracle/): cpu_baseline and --impl reference
# ------------------------------------------------------------------------------------------------
def cpu_reference_step(B, H, W, flow_sigma, seed=42):
    """Build closures that run the same op schedule as HotPathStep through the torch port on CPU."""
    from oracle import torch_port
    from unsamflow_b200 import hotpath
    gen = torch.Generator().manual_seed(seed)
    levels, scales = hotpath.pyramid_shapes(H, W), hotpath.loss_shapes(H, W)
    data = []
    for lvl, (C, h, w) in enumerate(levels):
        f1, f2 = torch.randn(B, C, h, w, generator=gen), torch.randn(B, C, h, w, generator=gen)
        for _ in range(2):  # two directions
            fl = hotpath.synthetic_flow(B, h, w, flow_sigma, gen)
            go = torch.randn(B, 81, h, w, generator=gen)
            data.append(("model", lvl, f1, f2, fl, go))
            f1, f2 = f2, f1
    for (C, h, w) in scales:
        for _ in range(2):
            im = torch.rand(B, C, h, w, generator=gen)
            fl = hotpath.synthetic_flow(B, h, w, flow_sigma, gen)
            go = torch.randn(B, C, h, w, generator=gen)
            data.append(("loss", 0, im, None, fl, go))

    def step():
        for kind, lvl, a, b, fl, go in data:
            if kind == "model":
                x1 = a.clone().requires_grad_(True)
                x2 = b.clone().requires_grad_(True)
                f = fl.clone().requires_grad_(True)
                x2w = torch_port.flow_warp(x2, f) if lvl > 0 else x2
                torch_port.cost_volume(x1, x2w).backward(go)
            else:
                f = fl.clone().requires_grad_(True)
                torch_port.flow_warp(a, f, pad="border").backward(go)
    return step


def time_cpu_reference(args, steps, warmup, sample_pairs=1, budget_s=120.0):
    """Times `steps` CPU steps (stops early once budget_s of timed work is spent; the JSON says so)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = cpu_reference_step(sample_pairs, args.height, args.width, args.flow_sigma)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    done = 0
    while done < steps and (done == 0 or time.perf_counter() - t0 < budget_s):
        step()
        done += 1
    steps = done
    dt = (time.perf_counter() - t0) / steps
    return {"value": sample_pairs / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample_pairs} pair(s) per step of the same op schedule ({steps} timed steps, "
                      f"{warmup} warm-up), oracle/torch_port.py on torch {torch.__version__} CPU",
            "ms_per_step": dt * 1e3, "steps": steps}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warmup = max(1, min(args.warmup, 2))
    cb = time_cpu_reference(args, max(1, args.steps), warmup)
    steps = cb["steps"]
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, extra={"note": "CPU reference path; step bounded to 1 pair; "
                                                        f"{steps} of --steps {args.steps} run inside a 120 s budget"}),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, extra=None):
    cfg = {"workload": f"PWC-Lite kitti_base training-step hot path (corr+flow_warp fwd+bwd), synthetic "
                       f"{args.height}x{args.width} pairs, B={args.batch} per GPU",
           "batch_per_gpu": args.batch, "height": args.height, "width": args.width, "max_displacement": 4,
           "flow_sigma_px": args.flow_sigma,
           "flow": "smooth synthetic flow: N(0,sigma^2) on a 24-px lattice, bilinear, + N(0,0.05^2) residual",
           "parallelism": f"dp{args.gpus} (batch replicas, no data-path collective)"}
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.period, self.samples, self.stop_flag, self.h, self.ok = period, [], False, None, False
        self.marks = []
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def run(self):
        while self.ok and not self.stop_flag:
            try:
                mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                try:
                    rs = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    rs = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, rs))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def summary(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvml unavailable: " + self.err}
        win = [s for s in self.samples if t0 <= s[0] <= t1]
        note = "sampled inside the timed region"
        if len(win) < 3:
            win, note = self.samples, "timed region too short for 3 samples: includes warm-up"
        mask = 0
        for s in win:
            mask |= s[2]
        reasons = sorted({name for bit, name in self.REASONS.items() if mask & bit})
        return {"sm_mhz": statistics.median([s[1] for s in win]) if win else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(win), "note": note}


# ------------------------------------------------------------------------------------------------
# ours
# ------------------------------------------------------------------------------------------------
def capture(fn):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


def ncu_traffic(op):
    """DRAM bytes per launch of this op's kernel from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t.get(f"{op.kind}." + "x".join(str(v) for v in op.shape))
    except Exception:  # noqa: BLE001
        return None


def per_launch_bytes(op):
    B, C, H, W = op.shape
    if op.kind == "corr_bwd" and op.launches == 2:  # cp.async fallback: one launch per gradient, each reads
        return (81 + 2 * C) * 4 * B * H * W         # gO + one feature map and writes one gradient
    return op.bytes // max(1, op.launches)            # fused TMA launch: the op's algorithmic bytes


def time_train_step(args, local, world, barrier):
    """Full synthetic PWC-Lite training step (unsamflow_b200/train_step.py): forward with_bk, unflow loss,
    backward, clip, Adam; DDP over NCCL when world > 1 (per-GPU batch fixed: weak scaling)."""
    from unsamflow_b200.train_step import TrainStep
    torch.backends.cudnn.benchmark = True  # train.py:19
    graphed = world == 1  # DDP's bucket hooks are not captured: the multi-GPU step stays eager
    ts = TrainStep(args.batch, args.height, args.width, f"cuda:{local}", seed=42, batch_directions=True,
                   ddp=world > 1, cuda_graph=graphed, fuse_cost_concat=True)
    for _ in range(3):
        ts.step()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.train_steps):
        loss = ts.step()
    b.record()
    barrier()
    info = {"what": "PWC-Lite kitti_base training step (fwd with_bk + unflow loss + bwd + clip + Adam), synthetic "
                    f"{args.height}x{args.width} pairs, B={args.batch} per GPU, PyTorch / cuDNN around the usf kernels, "
                    f"{'captured in one CUDA graph' if graphed else 'eager (DDP)'}",
            "batch_directions": True, "fuse_cost_concat": True, "cuda_graph": graphed, "loss": float(loss), "ddp": world > 1,
            "params": ts.model.num_parameters()}
    return a.elapsed_time(b), info


def run_ours(args):
    from unsamflow_b200 import ops
    from unsamflow_b200.hotpath import HotPathStep
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; unsamflow_b200 has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    step = HotPathStep(args.batch, args.height, args.width, device=f"cuda:{local}", flow_sigma=args.flow_sigma,
                       batch_directions=args.batch_directions)
    if args.batch_directions:
        args.streams = 1
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=step.device)  # > 126 MB L2
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    # ---- per-op table (also picks the dominant op) ------------------------------------------------
    step.measure_launches()
    torch.cuda.synchronize()
    table = []
    for op in step.ops:
        ts = []
        for _ in range(5):
            flush.zero_()
            a, b = ev(), ev()
            a.record(); op.fn(); b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        us = statistics.median(ts)
        table.append({"op": op.name, "kind": op.kind, "shape": list(op.shape), "us_cold_l2": round(us, 2),
                      "launches": op.launches, "alg_MB": round(op.bytes / 1e6, 3),
                      "GBps": round(op.bytes / us / 1e3, 1), "TFLOPs": round(op.flops / us / 1e6, 2)})
    # dominant KERNEL: the op with the longest average launch (a multi-kernel op such as the three-launch warp
    # backward is not one kernel, and launching it eagerly would put host gaps into the step)
    dom_i = max(range(len(step.ops)), key=lambda i: table[i]["us_cold_l2"] / max(1, table[i]["launches"]))
    dom = step.ops[dom_i]
    if args.op_table and rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.op_table)), exist_ok=True)
        with open(args.op_table, "w") as f:
            json.dump(table, f, indent=1)

    # ---- the step: ONE CUDA graph, the two flow directions on two streams inside it (the reference calls its two
    # decoders back to back although they are independent, pwclite.py:430-432; --streams 1 serialises them).  The
    # dominant op's kernel is timed where it runs: two external timing events recorded inside the graph on the
    # kernel's own stream, around its launch (CUDA allows event-record nodes with cudaEventRecordExternal in a
    # capture), read after every replay.  If this torch cannot record external events during capture, the step
    # falls back to round 1's form: graph A -> the dominant op launched eagerly between two events -> graph B.
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step.run()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    if args.streams == 2:
        step.run_concurrent(step.chains)  # warm the side streams up
        torch.cuda.synchronize()
    run_full = (lambda: step.run_concurrent(step.chains)) if args.streams == 2 else step.run
    gMarked = None
    step_form = ("one CUDA graph per step (value, ms_per_step); dominant kernel timed in place by external events "
                 "inside that graph, and alone -- graph A -> kernel between two events -> graph B -- in a second loop")
    k0 = k1 = None
    try:
        k0 = torch.cuda.Event(enable_timing=True, external=True)
        k1 = torch.cuda.Event(enable_timing=True, external=True)
        step.marks = {dom.name: (k0, k1)}
        if args.streams != 2:  # single stream: step.run() has no hook, wrap the op itself
            base_fn = dom.fn

            def marked_fn():
                k0.record()
                base_fn()
                k1.record()
            dom.fn = marked_fn
        gMarked = capture(run_full)
        gMarked.replay()
        torch.cuda.synchronize()
        if not (0.0 < k0.elapsed_time(k1) < 1e3):
            raise RuntimeError("external events inside the graph returned no usable time")
    except Exception as exc:  # noqa: BLE001
        step_form = f"one CUDA graph per step; external events unavailable in capture ({exc!r}): kernel timed alone only"
        gMarked = None
    finally:
        step.marks = {}
        if args.streams != 2 and "base_fn" in locals():
            dom.fn = base_fn
    gFull = gMarked if gMarked is not None else capture(run_full)
    # the split form: everything that precedes the dominant op, the op alone, everything after it
    if args.streams == 2:
        pre, post = step.split_around(dom)
        step.run_concurrent(pre); dom.fn(); step.run_concurrent(post)
        torch.cuda.synchronize()
        gA = capture(lambda: step.run_concurrent(pre))
        gB = capture(lambda: step.run_concurrent(post))
    else:
        before, after = step.ops[:dom_i], step.ops[dom_i + 1:]
        gA = capture(lambda: step.run(before)) if before else None
        gB = capture(lambda: step.run(after)) if after else None

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        flush.zero_()
        gFull.replay()
    barrier()
    ops.reset_launch_count()
    t_wall0 = time.perf_counter()
    total_ms, inplace_ms = 0.0, 0.0
    for _ in range(args.steps):
        flush.zero_()
        s0, s1 = ev(), ev()
        s0.record()
        gFull.replay()
        s1.record()
        s1.synchronize()  # between steps, outside the event pair: the in-graph events are re-recorded by the next replay
        total_ms += s0.elapsed_time(s1)
        if gMarked is not None:
            inplace_ms += k0.elapsed_time(k1)
    barrier()
    t_wall1 = time.perf_counter()
    # the dominant kernel alone inside the step (round 1's form of the step; comparable across rounds): nothing else
    # runs between the two events, its launch is already queued behind graph A
    alone_steps = max(5, min(args.steps, 50))
    alone_ms = 0.0
    for i in range(3 + alone_steps):
        flush.zero_()
        e0, e1 = ev(), ev()
        if gA is not None:
            gA.replay()
        e0.record()
        dom.fn()
        e1.record()
        if gB is not None:
            gB.replay()
        e1.synchronize()
        if i >= 3:
            alone_ms += e0.elapsed_time(e1)
    barrier()
    eager_launches = ops.launch_count()  # launches issued through the ABI outside graph replays
    dom_ms = alone_ms / alone_steps
    inplace_us = inplace_ms / args.steps * 1e3 / max(1, dom.launches) if gMarked is not None else None
    clocks = sampler.summary(t_wall0, t_wall1)

    # ---- e2e: host inputs -> H2D -> step -> D2H results, all inside the timed region ---------------
    # Every step copies ALL of its inputs from pinned host memory and reads its results back; the copies
    # of step i+1 / i-1 overlap the kernels of step i (two device buffer sets, copy-in / compute / copy-out
    # streams chained by events), so the sustained rate is max(PCIe time, kernel time) rather than their sum.
    step2 = HotPathStep(args.batch, args.height, args.width, device=f"cuda:{local}", flow_sigma=args.flow_sigma,
                        batch_directions=args.batch_directions)
    torch.cuda.synchronize()
    g2 = capture(lambda: step2.run_concurrent(step2.chains)) if args.streams == 2 else capture(step2.run)
    sets = [(step, gFull), (step2, g2)]
    host_in = step.input_arena.cpu().pin_memory()  # every input of the step, one pinned block (see HotPathStep._pack)
    host_out = [torch.empty_like(st.result_arena, device="cpu").pin_memory() for st, _ in sets]
    s_in, s_run, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    in_ready = [torch.cuda.Event() for _ in sets]
    run_done = [torch.cuda.Event() for _ in sets]
    out_done = [torch.cuda.Event() for _ in sets]

    def e2e_step(i):
        st, g = sets[i % 2]
        with torch.cuda.stream(s_in):
            s_in.wait_event(run_done[i % 2])          # this buffer set's previous kernels are finished
            st.input_arena.copy_(host_in, non_blocking=True)
            in_ready[i % 2].record(s_in)
        with torch.cuda.stream(s_run):
            s_run.wait_event(in_ready[i % 2])
            s_run.wait_event(out_done[i % 2])         # previous results of this set have left the device
            g.replay()
            run_done[i % 2].record(s_run)
        with torch.cuda.stream(s_out):
            s_out.wait_event(run_done[i % 2])
            host_out[i % 2].copy_(st.result_arena, non_blocking=True)
            out_done[i % 2].record(s_out)

    for i in range(4):
        e2e_step(i)
    barrier()
    e2e_a, e2e_b = ev(), ev()
    e2e_a.record(s_in)
    for i in range(args.e2e_steps):
        e2e_step(i)
    s_out.wait_stream(s_in)
    s_out.wait_stream(s_run)
    e2e_b.record(s_out)
    barrier()
    e2e_ms = e2e_a.elapsed_time(e2e_b)
    sampler.stop_flag = True

    # ---- the whole PWC-Lite training step around the path (cuDNN convs + loss + Adam are PyTorch) --------
    train_ms, train_info = 0.0, None
    if args.train_steps > 0:
        del gA, gB, gFull, gMarked
        train_ms, train_info = time_train_step(args, local, world, barrier)

    from unsamflow_b200 import dist_util
    total_ms, e2e_ms, train_ms = dist_util.max_over_ranks([total_ms, e2e_ms, train_ms])  # NCCL: on this rank's GPU

    if rank == 0:
        peak, peak_src = peaks()
        launches = dom.launches
        alg = per_launch_bytes(dom)
        dom_us = dom_ms * 1e3 / launches
        achieved = alg / dom_us / 1e3  # GB/s
        line = {
            "metric": METRIC, "value": world * args.batch * args.steps / (total_ms / 1e3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, extra={
                "l2": "256 MB L2 flush between timed steps (outside the per-step CUDA-event pair); "
                      "step working set ~1 GB > 126 MB L2",
                "step": step_form,
                "streams": args.streams, "batch_directions": args.batch_directions}),
            "hot_path_gbs": world * step.bytes_per_step * args.steps / (total_ms / 1e3) / 1e9,
            "hot_path_tflops": world * step.flops_per_step * args.steps / (total_ms / 1e3) / 1e12,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(dom), "kernel": dom.name, "kind": dom.kind, "shape": list(dom.shape),
                         "launches_per_op": launches, "alg_bytes_per_launch": alg, "us_per_launch": dom_us,
                         "us_per_launch_in_place": inplace_us,
                         "timing": f"us_per_launch: kernel alone between two CUDA events inside the step (graph A -> kernel -> "
                                   f"graph B, {alone_steps} steps with the same L2 flush, run right after the timed steps); "
                                   "us_per_launch_in_place: external events around the kernel inside every timed step's "
                                   "graph, the other direction's stream running beside it",
                         "peak_source": peak_src,
                         "fp32_tflops": (dom.flops / launches) / dom_us / 1e6 if dom.flops else None,
                         "fp32_peak_tflops": FP32_PEAK_TFLOPS},
            "e2e": {"value": world * args.batch * args.e2e_steps / (e2e_ms / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": step.input_bytes(),
                    "d2h_bytes_per_step": step.result_bytes(),
                    "steps": args.e2e_steps, "ms_per_step": e2e_ms / args.e2e_steps,
                    "pipeline": "double-buffered: H2D of step i+1 and D2H of step i-1 overlap the kernels of step i; one pinned "
                                "block per direction (inputs / results live in one device arena each)"},
            "gpu_launches": step.launches_per_step * args.steps,
            "gpu_launches_note": f"{step.launches_per_step} usf kernels per step, replayed from one CUDA graph in the timed "
                                 f"region ({eager_launches} more issued eagerly through the C ABI by the kernel-alone loop)",
            "clocks": clocks,
        }
        if train_info is not None:
            train_info.update({"value": world * args.batch * args.train_steps / (train_ms / 1e3), "unit": UNIT,
                               "ms_per_step": train_ms / args.train_steps, "steps": args.train_steps})
            line["train_step"] = train_info
        if world == 1 and not args.no_cpu_baseline:
            cb = time_cpu_reference(args, steps=args.cpu_reps, warmup=1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
