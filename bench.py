"""bench.py - headline benchmark of the kNN voice-library matching path.

    python bench.py --gpus N --steps K --warmup W [--workload cfg4] [--impl reference]

Metric (BASELINE.json): query frames/sec matched (k=4) vs library size N at 1/2/4/8 B200.
Default workload = BASELINE configs[3] ("cfg4"): T=10,000 query frames against a
10,000,000-frame library (D=768, k=4), the library row-sharded over the N GPUs of one box
(strong scaling: total work fixed).  It fits one B200 (15.4 GB bf16 + 30.7 GB fp32), so the
same workload is used at N=1.  Other workloads: cfg1, cfg2 (latency), cfg3, cfg5.

A "step" = one pass of the hot path over one batch of T synthetic query frames:
  value : inputs resident in HBM, CUDA-event timed, barrier + synchronize on both sides,
          max over ranks, value = T / time
  e2e   : the same call through the public API with HOST buffers: pinned-host queries ->
          H2D -> match -> D2H of the matched features, every step
  roofline      : the fused tcgen05 similarity+top-list kernel (alive_knn_search), its own
                  CUDA-event duration inside the timed region vs MEASURED_PEAKS.json
  roofline_gather : the standalone gather-mean kernel (K4) on the step's own indices vs the HBM peak, with
                  torch's index_select of the same rows beside it (the practical ceiling of random 3 KB rows)
  roofline_pack : the library pack (K1) alone on a fresh 250k-frame chunk, channel-major (the reference's
                  layout) and row-major, vs the HBM peak; torch's transposing copy of the chunk beside it
  cpu_baseline  : the oracle's torch port of the reference (oracle/knn_oracle.py
                  match_features_torch = module/common.py:96-109 on CPU) on the box's host
                  cores, bounded sample, rank 0 at N=1 only
`--impl reference` times that CPU arm alone on the same config and prints the same line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

D = 768
K = 4

WORKLOADS = {
    # name: (B, T, N, description)
    "cfg1": (1, 1000, 100_000, "cfg1: 10 s utterance T=1000 frames vs 100k-frame library"),
    "cfg2": (1, 32, 200_000, "cfg2: streaming chunk T=32 frames vs 200k-frame library (latency)"),
    "cfg3": (1, 100_000, 1_000_000, "cfg3: offline batch T=100k frames vs 1M-frame library"),
    "cfg4": (1, 10_000, 10_000_000, "cfg4: T=10k frames vs 10M-frame library row-sharded over the GPUs"),
    "cfg5": (64, 1000, 500_000, "cfg5: 64 utterances x 1000 frames vs 64 per-speaker 500k-frame libraries"),
    # the reference's own default operating points: --chunk 48000 -> T=450 per call,
    # realtime chunk 960 x buffer 8 -> T=24 per call, library = one target utterance + 512 tokens
    "real_offline": (1, 450, 3512, "inference.py defaults: T=450 frames per chunk vs N=3512-frame library"),
    "real_realtime": (1, 24, 3512, "realtime_inference.py defaults: T=24 frames per chunk vs N=3512-frame library"),
}


# cfg2 is a latency config: SURVEY §8(d) asks for >= 1000 timed calls
DEFAULT_STEPS = {"cfg1": 50, "cfg2": 1000, "cfg3": 5, "cfg4": 5, "cfg5": 3, "real_offline": 1000, "real_realtime": 1000}


def load_traffic(workload: str, world: int):
    """DRAM bytes per launch of the dominant kernel from a committed ncu --set full capture
    (profiles/traffic_r01.json), or None when this (workload, n_gpus) has not been captured."""
    p = os.path.join(ROOT, "profiles", "traffic_r01.json")
    try:
        with open(p) as f:
            ent = json.load(f).get(f"{workload}@{world}")
        return ent["dram_bytes_per_launch"] if ent else None
    except Exception:
        return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"bf16_burst": d.get("bf16_tflops"), "bf16_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw = [], [], []
        reasons = set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# -------------------------------------------------------------------------------------------
# CPU arm: the oracle's torch port of the reference (kind "port"; the reference itself is
# Python under /root/reference and cannot travel to the GPU box)
# -------------------------------------------------------------------------------------------
def cpu_sample_shape(workload: str):
    """Bounded sample (seconds, not minutes, of CPU work) of the named workload and the factor
    that converts its rate to the full workload's query-frames/s."""
    B, T, N, _ = WORKLOADS[workload]
    if workload == "cfg1":
        return 1000, 100_000, 1.0, "full cfg1 (T=1000, N=100k)"
    if workload == "cfg2":
        return 32, 200_000, 1.0, "full cfg2 (T=32, N=200k), library re-normalised every call as the reference does"
    if workload == "cfg3":
        return 256, 1_000_000, 1.0, "T=256-frame tile of cfg3 vs the full 1M-frame library (per-frame rate)"
    if workload == "cfg4":
        return 64, 1_000_000, 0.1, ("T=64-frame tile vs a 1M-frame slice (1/10 of the 10M library); "
                                    "rate scaled x0.1 linearly in N (extrapolated)")
    if workload in ("real_offline", "real_realtime"):
        return T, N, 1.0, f"full {workload} (T={T}, N={N})"
    return 256, 500_000, 1.0, "T=256-frame tile of one speaker vs its 500k-frame library (per-frame rate)"


def run_cpu_arm(workload: str, steps: int, warmup: int):
    import torch
    from oracle.knn_oracle import match_features_torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    t_s, n_s, scale, what = cpu_sample_shape(workload)
    g = torch.Generator().manual_seed(1234)
    src = torch.randn(1, D, t_s, generator=g)
    ref = torch.randn(1, D, n_s, generator=g)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        match_features_torch(src, ref, K, 0.0)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    qps = t_s / mean * scale
    return {"value": qps, "unit": "query_frames/s", "cores": cores, "kind": "port",
            "sample": what, "sample_seconds_per_step": mean, "steps": len(times)}


def run_torch_eager_gpu(workload: str, dev):
    """Secondary line: the reference's own torch ops (norm, div, bmm, topk, gather,
    mean = oracle port of common.py:96-109) run eagerly on the same B200 - i.e. what `-d cuda` gives
    a user of the reference today.  Same bounded samples as the CPU arm (the [T,N] fp32 score
    matrix of the full cfg3/cfg4/cfg5 shapes does not fit)."""
    import torch
    from oracle.knn_oracle import match_features_torch

    t_s, n_s, scale, what = cpu_sample_shape(workload)
    g = torch.Generator(device=dev).manual_seed(1234)
    src = torch.randn(1, D, t_s, device=dev, generator=g)
    ref = torch.randn(1, D, n_s, device=dev, generator=g)
    for _ in range(3):
        match_features_torch(src, ref, K, 0.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        match_features_torch(src, ref, K, 0.0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    del src, ref
    torch.cuda.empty_cache()
    return {"value": t_s / (ms * 1e-3) * scale, "unit": "query_frames/s", "ms_per_call": ms, "sample": what,
            "note": "torch eager fp32 (cuBLAS sgemm, TF32 off) incl. per-call library normalisation"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, T, N, desc = WORKLOADS[args.workload]
    steps = max(1, min(args.steps, 100))      # each step is a bounded sample (about 1 s of CPU work)
    warmup = max(1, min(args.warmup, 10))
    cb = run_cpu_arm(args.workload, steps, warmup)
    line = {
        "impl": "reference", "metric": "query_frames_per_sec_matched_k4", "value": cb["value"],
        "unit": "query_frames/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": cb["sample_seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "B": B, "T": T, "N": N, "D": D, "k": K},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "query_frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------
def build_library(n_lo, n_hi, seed, device, chunk=250_000):
    """Rows [n_lo, n_hi) of the synthetic random-normal library, generated and packed shard by
    shard on the device (the fp32 [D, n] source chunk is freed after packing)."""
    import torch
    from alive_vc_b200 import matching as M

    n = n_hi - n_lo
    lib = M.alloc_packed(n, D, device)
    lib.row_base = n_lo
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        g = torch.Generator(device=device).manual_seed(seed * 1_000_003 + (n_lo + c0) // chunk)
        x = torch.randn(D, c1 - c0, device=device, generator=g)
        M.pack_into(lib, c0, x)
        del x
    return lib


def gpu_arm(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry
    from alive_vc_b200 import matching as M
    from alive_vc_b200.sharded import CudaShardBackend, ShardedLibrary, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1 (one process per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        entry.build()          # no-op when the in-tree .so is current (it travels with the snapshot)
    if world > 1:
        dist.barrier()         # nobody loads the library before rank 0 has (re)built it
    B, T, N, desc = WORKLOADS[args.workload]
    peaks = load_peaks()
    variant = args.variant

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident data: the (sharded) packed library, device queries, pinned host queries ----
    g = torch.Generator(device=dev).manual_seed(args.seed)
    if args.workload == "cfg5":
        if world > 1:
            per = B // world
            my_items = list(range(rank * per, (rank + 1) * per))
        else:
            my_items = list(range(B))
        # all of this rank's per-speaker libraries packed back to back: ONE pipeline launch per step
        lib = M.alloc_packed(len(my_items) * N, D, dev)
        lib.items = len(my_items)
        for i, b in enumerate(my_items):
            for c0 in range(0, N, 250_000):
                c1 = min(N, c0 + 250_000)
                gg = torch.Generator(device=dev).manual_seed((args.seed + 17 * (b + 1)) * 1_000_003 + c0 // 250_000)
                x = torch.randn(D, c1 - c0, device=dev, generator=gg)
                M.pack_into(lib, i * N + c0, x)
                del x
        libs = [lib]
        src_dev = torch.randn(len(my_items), D, T, device=dev, generator=g)

        def step(src):
            o, _, _ = M.match_packed(src, lib, K, 0.0, "screen", variant)
            return o.transpose(1, 2)
        units_per_step = B * T
        n_local = N
        scaling = "weak" if world == 1 else "strong"
        parallelism = f"{len(my_items)} speakers per GPU, no collective"
    else:
        lo, hi = shard_bounds(N, world, rank)
        lib = build_library(lo, hi, args.seed, dev)
        sharded = ShardedLibrary(CudaShardBackend(lib, "screen", variant), lib.n, lo, N, None,
                                 peer_memory=(args.exchange == "peer"))
        src_dev = torch.randn(B, D, T, device=dev, generator=g)
        if world > 1:
            dist.broadcast(src_dev, 0)

        def step(src):
            return sharded.match(src, K, 0.0)
        if args.workload in ("cfg1", "cfg2", "real_offline", "real_realtime") and world == 1 and not args.no_graph:
            # fixed-shape chunks (inference.py / realtime_inference.py call the match once per chunk with
            # the same T): pre-allocated buffers, the whole pipeline replayed as one CUDA graph
            streamer = M.StreamingMatcher(lib, T, K, 0.0, batch=B, mode="screen", variant=variant)

            def step(src):                                             # noqa: F811
                return streamer(src)
        units_per_step = B * T
        n_local = hi - lo
        scaling = "strong"
        parallelism = f"library rows sharded x{world}" + (
            "" if world == 1 else
            ", all-gather top-k + one peer-memory (NVLink) gather kernel" if args.exchange == "peer" else
            ", all-gather top-k + reduce-scatter rows + all-gather result")
    src_host = torch.empty(src_dev.shape, dtype=torch.float32).pin_memory()
    src_host.copy_(src_dev)
    # the result is a transposed view of a contiguous [B,T,D] block (like the reference's); the pinned
    # host buffer has the same strides, so the device->host read is one plain memcpy
    _b, _d, _t = src_dev.shape
    out_host = torch.empty((_b, _t, _d), dtype=torch.float32).pin_memory().transpose(1, 2)
    barrier()

    # ---- device-resident timing ----
    for _ in range(args.warmup):
        step(src_dev)
    barrier()
    M.search_events = []
    launches0 = M.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lat = []
    barrier()
    e0.record()
    latency_workload = args.workload in ("cfg2", "real_offline", "real_realtime")
    if latency_workload:
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        evs[0].record()
        for i in range(args.steps):
            step(src_dev)
            evs[i + 1].record()
    else:
        for _ in range(args.steps):
            step(src_dev)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    launches = M.launch_count - launches0
    search_ms = [a.elapsed_time(b) for a, b in M.search_events]
    if not search_ms:
        # graph-replayed path: the kernel cannot be bracketed inside the graph, so time the same
        # launches eagerly right after the timed region (same buffers, same clocks)
        for _ in range(min(args.steps, 50)):
            sharded.match(src_dev, K, 0.0)
        torch.cuda.synchronize()
        search_ms = [a.elapsed_time(b) for a, b in M.search_events]
    M.search_events = None
    if latency_workload:
        lat = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps))
    fallback = M.last_info.fallback_queries() if M.last_info is not None else 0
    if world > 1:
        tmax = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_total = float(tmax.item())
    ms_per_step = ms_total / args.steps
    value = units_per_step / (ms_per_step * 1e-3)

    # ---- end-to-end: host buffers through the public API ----
    streaming = args.workload in ("cfg1", "cfg2", "real_offline", "real_realtime") and world == 1 and not args.no_graph

    def e2e_step():
        # the streaming matcher copies the pinned host chunk straight into its static input buffer
        s = src_host if streaming else src_host.to(dev, non_blocking=True)
        o = step(s)
        out_host.copy_(o, non_blocking=True)
    for _ in range(min(args.warmup, 3)):
        e2e_step()
    barrier()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        tmax = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_ms = float(tmax.item())
    e2e_value = units_per_step / (e2e_ms / args.steps * 1e-3)
    io_bytes = src_host.numel() * 4

    # latency workloads: the synchronous host-to-host chunk time of the realtime loop
    # (realtime_inference.py:158-176) - pinned buffers, ONE graph (H2D + pipeline + D2H), one event wait
    host_lat = []
    if latency_workload and world == 1 and not args.no_graph:
        from alive_vc_b200.lifecycle import HostStreamingMatcher
        hm = HostStreamingMatcher(lib, T, K, 0.0, batch=B, mode="screen", variant=variant)
        chunk = src_host.clone()
        for _ in range(20):
            hm(chunk)
        for _ in range(min(args.steps, 2000)):
            t0 = time.perf_counter()
            hm(chunk)
            host_lat.append((time.perf_counter() - t0) * 1e3)
        host_lat.sort()

    if rank == 0:
        # roofline of the dominant kernel: algorithmic flops of ONE alive_knn_search launch
        # (2 * T * N_local * D) over its average CUDA-event duration
        if args.workload == "cfg5":
            flops_per_launch = 2.0 * lib.items * T * N * D
        else:
            flops_per_launch = 2.0 * B * T * n_local * D
        avg_search_ms = sum(search_ms) / max(1, len(search_ms))
        if args.workload == "cfg5":
            search_kernel = "knn_search_kernel"
        else:
            _plan = M.make_plan(B * T, n_local, D, dev, variant)
            search_kernel = "knn_search_skinny_kernel" if _plan.kernel == 1 else "knn_search_kernel"
        if latency_workload:
            bytes_per_launch = float(n_local) * D * 2
            achieved = bytes_per_launch / (avg_search_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "traffic": load_traffic(args.workload, world),
                    "kernel": search_kernel, "avg_kernel_ms": avg_search_ms, "peak_source": peaks["source"]}
        else:
            achieved = flops_per_launch / (avg_search_ms * 1e-3) / 1e12
            # conservative denominator: the burst cuBLAS figure, even though the kernel runs
            # back-to-back under the power cap (frac_of_sustained is reported beside it)
            sustained = False
            peak = peaks["bf16_burst"]
            roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": load_traffic(args.workload, world), "kernel": search_kernel,
                    "avg_kernel_ms": avg_search_ms,
                    "kernel_share_of_step": avg_search_ms / ms_per_step,
                    "peak_kind": "sustained" if sustained else "burst", "peak_source": peaks["source"],
                    "frac_of_burst": achieved / peaks["bf16_burst"],
                    "frac_of_sustained": achieved / peaks["bf16_sustained"]}
        # K4 alone (north_star: "fraction of HBM bandwidth for the gather"): the standalone gather-mean
        # kernel on this step's own neighbour indices; in the pipeline the same arithmetic is fused
        # into finish_kernel.  Algorithmic bytes per query frame: k*D*4 read + D*4 write (+ D*4 query).
        gather_roof = None
        if args.workload != "cfg5" and world == 1:   # (cfg5: batched libraries, skipped)
            info = M.last_info
            ws = getattr(info, "_workspace", None)
            if ws is not None:
                rows = B * T
                q_view = M.PackedFrames(n=rows, d=D, raw=ws[: rows * D * 4].view(torch.float32).view(rows, D),
                                        norms=None, packed=None, err=None, stats=None)
                _, g_idx, _ = M.run_match(src_dev, lib, K, 0.0, "screen", variant, want_out=False)
                g_out = torch.empty((rows, D), dtype=torch.float32, device=dev)
                for _ in range(3):
                    M.gather_mean(lib, g_idx.view(rows, K), q_view, 0.0, g_out)
                torch.cuda.synchronize()
                ge0, ge1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 20
                ge0.record()
                for _ in range(reps):
                    M.gather_mean(lib, g_idx.view(rows, K), q_view, 0.0, g_out)
                ge1.record()
                torch.cuda.synchronize()
                g_ms = ge0.elapsed_time(ge1) / reps
                g_bytes = rows * (K * D * 4 + D * 4 + D * 4)
                # what the memory system gives ANY kernel for this access pattern: torch's own row gather of
                # the same T*k random 3 KB rows (read + write T*k rows) - the practical ceiling beside the
                # streaming-copy peak
                flat_idx = g_idx.reshape(-1)
                sel_out = torch.empty((flat_idx.numel(), D), dtype=torch.float32, device=dev)
                for _ in range(3):
                    torch.index_select(lib.raw, 0, flat_idx, out=sel_out)
                torch.cuda.synchronize()
                ge0.record()
                for _ in range(reps):
                    torch.index_select(lib.raw, 0, flat_idx, out=sel_out)
                ge1.record()
                torch.cuda.synchronize()
                sel_ms = ge0.elapsed_time(ge1) / reps
                sel_gbs = 2.0 * flat_idx.numel() * D * 4 / (sel_ms * 1e-3) / 1e9
                del sel_out
                gather_roof = {"bound": "hbm", "kernel": "gather_mean_kernel", "achieved": g_bytes / (g_ms * 1e-3) / 1e9,
                               "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": g_bytes / (g_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                               "avg_kernel_ms": g_ms, "bytes_per_query_frame": g_bytes // rows,
                               "torch_index_select_gbs": sel_gbs,
                               "note": "standalone K4 on this step's indices; random 3 KB rows of the raw library; "
                                       "torch_index_select_gbs = torch's row gather of the same rows (read + write), "
                                       "the practical ceiling of this access pattern"}
        # K1 alone (once per library, generate_voice_library.py / load time): the pack kernel on a fresh
        # channel-major [D, n] chunk far larger than L2.  Algorithmic bytes per frame: D*(4 read + 4 raw
        # + 2 packed) + 8 (norm, err) = 7,688 B.
        pack_roof = None
        if world == 1:
            pn = 250_000
            px = torch.randn(D, pn, device=dev)
            pdst = M.alloc_packed(pn, D, dev)
            for _ in range(2):
                M.pack_into(pdst, 0, px)
            torch.cuda.synchronize()
            pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            pe0.record()
            for _ in range(reps):
                M.pack_into(pdst, 0, px)
            pe1.record()
            torch.cuda.synchronize()
            p_ms = pe0.elapsed_time(pe1) / reps
            p_bytes = pn * (D * 10 + 8)
            # the same frames as a row-major producer would hand them over (pack_rows): pack_rm_kernel
            px_rows = px.t().contiguous()
            for _ in range(2):
                M.pack_into(pdst, 0, px_rows.t())
            pe0.record()
            for _ in range(reps):
                M.pack_into(pdst, 0, px_rows.t())
            pe1.record()
            torch.cuda.synchronize()
            p_ms_rows = pe0.elapsed_time(pe1) / reps
            # torch's own transposing copy of the same chunk ([768, n] -> [n, 768], read + write 4 B per element):
            # what a library kernel gets out of the channel-major access pattern
            t_out = torch.empty((pn, D), dtype=torch.float32, device=dev)
            for _ in range(2):
                t_out.copy_(px.t())
            pe0.record()
            for _ in range(reps):
                t_out.copy_(px.t())
            pe1.record()
            torch.cuda.synchronize()
            t_gbs = 2.0 * pn * D * 4 / (pe0.elapsed_time(pe1) / reps * 1e-3) / 1e9
            del t_out
            pack_roof = {"bound": "hbm", "kernel": "pack_cm_kernel", "achieved": p_bytes / (p_ms * 1e-3) / 1e9,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": p_bytes / (p_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                         "avg_kernel_ms": p_ms, "bytes_per_frame": D * 10 + 8, "frames": pn,
                         "row_major": {"kernel": "pack_rm_kernel", "achieved": p_bytes / (p_ms_rows * 1e-3) / 1e9,
                                       "frac": p_bytes / (p_ms_rows * 1e-3) / 1e9 / peaks["hbm_gbs"], "avg_kernel_ms": p_ms_rows},
                         "torch_transpose_copy_gbs": t_gbs,
                         "note": "standalone K1 on a channel-major [768, 250k] fp32 chunk (1.9 GB per launch, no L2 reuse); "
                                 "row_major = the same frames as [250k, 768] rows"}
            del px, pdst, px_rows
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu = run_cpu_arm(args.workload, 3, 1)
        eager = None
        if world == 1 and args.torch_eager:
            eager = run_torch_eager_gpu(args.workload, dev)
        line = {
            "metric": "query_frames_per_sec_matched_k4", "value": value, "unit": "query_frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": desc, "B": B, "T": T, "N": N, "D": D, "k": K, "parallelism": parallelism,
                       "l2": "library (bf16 %.1f GB per GPU) is far larger than L2, no flush needed"
                             % (n_local * D * 2 / 1e9) if n_local * D * 2 > 256e6 else
                             "library smaller than 2x L2: numbers are warm-L2 steady state of a resident library",
                       "variant": variant, "fallback_queries_last_step": fallback,
                       "api": "StreamingMatcher (one CUDA graph per chunk)" if streaming else
                              ("match_packed on pack_libraries (one launch for all speakers)" if args.workload == "cfg5" else "ShardedLibrary.match")},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "query_frames/s", "h2d_bytes_per_step": io_bytes,
                    "d2h_bytes_per_step": io_bytes, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        if eager is not None:
            line["torch_eager_gpu"] = eager
        if gather_roof is not None:
            line["roofline_gather"] = gather_roof
        if pack_roof is not None:
            line["roofline_pack"] = pack_roof
        if lat:
            line["latency_ms"] = {"p50": lat[len(lat) // 2], "p99": lat[min(len(lat) - 1, int(len(lat) * 0.99))],
                                  "min": lat[0]}
        if host_lat:
            line["host_chunk_latency_ms"] = {"p50": host_lat[len(host_lat) // 2],
                                             "p99": host_lat[min(len(host_lat) - 1, int(len(host_lat) * 0.99))],
                                             "min": host_lat[0], "timer": "perf_counter around one blocking call"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: per workload, see DEFAULT_STEPS)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--variant", type=int, default=0, help="0 default, 1 = cta_group::1, 2 = CTA pair")
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="cfg1/cfg2: do not use the CUDA-graph streaming matcher")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "peer"],
                    help="multi-GPU row exchange: NCCL reduce-scatter/all-gather, or one gather kernel over CUDA-IPC peer memory")
    ap.add_argument("--torch-eager", action="store_true",
                    help="also time the reference's torch ops on the GPU (secondary line, where it fits)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = DEFAULT_STEPS[args.workload]
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
