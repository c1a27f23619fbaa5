"""bench.py - headline benchmark of the kNN voice-library matching path.

    python bench.py --gpus N --steps K --warmup W [--workload cfg4] [--impl reference] [--no-extras]

Metric (BASELINE.json): query frames/sec matched (k=4) vs library size N at 1/2/4/8 B200; p50 chunk latency.
Headline workload = BASELINE configs[3] ("cfg4"): T=10,000 query frames against a 10,000,000-frame library
(D=768, k=4), the library row-sharded over the N GPUs of one box (strong scaling: total work fixed).  It fits one
B200 (15.4 GB bf16 + 30.7 GB fp32), so the same workload is used at N=1.

A "step" = one pass of the hot path over one batch of T synthetic query frames:
  value : inputs resident in HBM, CUDA-event timed, barrier + synchronize on both sides, max over ranks,
          value = T / time
  e2e   : the same call through the public API with HOST buffers: pinned-host queries -> H2D -> match -> D2H of
          the matched features, every step.  Throughput workloads go through lifecycle.HostPipeline (double-buffered
          device staging: the copies of neighbouring steps overlap the match; the timed region ends when the last
          result has reached the host), latency workloads through HostStreamingMatcher (one blocking chunk at a
          time).  At N > 1 every query byte crosses PCIe ONCE: rank r copies its T/N slice of the frames, the slices
          are all-gathered over NVLink, and rank r reads back its T/N slice of the result
          (ShardedLibrary.match(scattered=True) inside the same pipeline); the byte counts are the whole job's.
  roofline        : the fused tcgen05 similarity+top-list kernel (alive_knn_search), its own CUDA-event duration
                    inside the timed region vs MEASURED_PEAKS.json
  roofline_gather : the standalone gather-mean kernel (K4) on the step's own indices vs the HBM peak (algorithmic
                    bytes k*D*4 read + D*4 written per query frame), torch's index_select of the same rows beside it
  roofline_pack   : the library pack (K1) alone on a fresh 250k-frame chunk, channel-major (the reference's
                    layout) and row-major, vs the HBM peak
  parity          : after the timed loops, 256 query frames through the SAME sharded call vs the exhaustive fp64
                    scan of every shard merged across ranks (indices) and vs the sequential mean of the raw frames
                    (features); any mismatch beyond a 1e-6 similarity tie makes the run exit 1
  cpu_baseline    : the oracle's torch port of the reference (oracle/knn_oracle.py match_features_torch =
                    module/common.py:96-109 on CPU) on the box's host cores, bounded sample, rank 0 at N=1 only
  workloads       : (N=1, default run) the other BASELINE configs measured the same way in the same process -
                    cfg2 (p50/p99 chunk latency, host-to-host chunk latency), cfg1, cfg3, cfg5 - and `clustered`:
                    cfg1's shape on a library of 100 tight clusters (nothing a bf16 screen can certify)
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
LATENCY_WORKLOADS = ("cfg2", "real_offline", "real_realtime")
GRAPH_WORKLOADS = ("cfg1", "cfg2", "real_offline", "real_realtime")

# cfg2 is a latency config: SURVEY §8(d) asks for >= 1000 timed calls
DEFAULT_STEPS = {"cfg1": 50, "cfg2": 1000, "cfg3": 5, "cfg4": 5, "cfg5": 3, "real_offline": 1000, "real_realtime": 1000}
EXTRA_WORKLOADS = ("cfg2", "cfg1", "cfg3", "cfg5")       # measured after the headline in the default N=1 run
PARITY_ROWS = 256


def load_traffic(workload: str, world: int):
    """DRAM bytes per launch of the dominant kernel from a committed ncu --set full capture
    (profiles/traffic_r02.json, else traffic_r01.json), or None when this (workload, n_gpus) has not been captured."""
    for name in ("traffic_r02.json", "traffic_r01.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                ent = json.load(f).get(f"{workload}@{world}")
            if ent:
                return ent["dram_bytes_per_launch"]
        except Exception:
            pass
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


def build_clustered(T, N, nclus, noise, seed, device):
    """cfg1's shape on a library of `nclus` tight clusters (centre + noise * N(0,1)): near-duplicate frames, what
    silence and sustained vowels look like to a content encoder - nothing a bf16 screen can certify."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    cent = torch.randn(D, nclus, device=device, generator=g)
    ref = cent[:, torch.randint(0, nclus, (N,), device=device, generator=g)] + noise * torch.randn(D, N, device=device, generator=g)
    src = cent[:, torch.randint(0, nclus, (T,), device=device, generator=g)] + noise * torch.randn(D, T, device=device, generator=g)
    return src[None].contiguous(), ref[None].contiguous()


class Env:
    pass


def parity_block(env, sharded, lib, src_dev, batched: bool):
    """256 query frames through the product call vs (a) the exhaustive fp64 scan of every shard (mode="exact")
    merged across ranks with alive_knn_merge_records, (b) the sequential float32 mean of the winning raw frames
    assembled with torch ops (each rank contributes the rows it owns; the sum over ranks adds zeros).  Returns
    the dict printed as `parity`; `ok` False makes the run exit 1."""
    import torch
    import torch.distributed as dist
    from alive_vc_b200 import _cabi, matching as M
    from alive_vc_b200.sharded import record_bytes

    dev, world = env.dev, env.world
    if batched:
        # cfg5: per-speaker libraries, no exchange step - screen vs exhaustive scan per item, 32 frames per speaker
        rows_per = 32
        q = src_dev[:, :, :rows_per].contiguous()
        out_s, idx_s, sc_s = M.match_packed(q, lib, K, 0.0, "screen", env.variant)
        out_e, idx_e, sc_e = M.match_packed(q, lib, K, 0.0, "exact", env.variant)
        rows = q.shape[0] * rows_per
        i_s, i_e = idx_s.reshape(rows, K), idx_e.reshape(rows, K)
        same = (i_s == i_e).all(dim=1)
        tie = (~same) & ((sc_s.reshape(rows, K) - sc_e.reshape(rows, K)).abs() <= 1e-6).all(dim=1)
        feat_ok = bool(torch.equal(out_s.reshape(rows, D)[same], out_e.reshape(rows, D)[same]))
        n_same, n_tie = int(same.sum()), int(tie.sum())
        return {"rows": rows, "index_exact": n_same, "tie_excused": n_tie, "features_bit_exact_rows": n_same if feat_ok else 0,
                "checker": "exhaustive fp64 scan per speaker (alive_knn_exact)", "ok": n_same + n_tie == rows and feat_ok}
    t_par = min(PARITY_ROWS, src_dev.shape[2])
    q = src_dev[:, :, :t_par].contiguous()
    out, idx = sharded.match(q, K, 0.0, return_indices=True)
    rows = q.shape[0] * t_par
    out_rows = out.transpose(1, 2).reshape(rows, D)
    idx = idx.reshape(rows, K)
    # (a) exhaustive scan of the local shard, records all-gathered, merged
    stride = record_bytes(rows, K)
    rec = torch.empty((stride,), dtype=torch.uint8, device=dev)
    rec_i = rec[: rows * K * 8].view(torch.int64)
    rec_s = rec[rows * K * 8: rows * K * 12].view(torch.float32)
    M.run_match(q, lib, K, 0.0, mode="exact", want_out=False,
                top_idx=rec_i.view(q.shape[0], t_par, K), top_score=rec_s.view(q.shape[0], t_par, K))
    if world > 1:
        gathered = torch.empty((world, stride), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(gathered.view(-1), rec)
    else:
        gathered = rec.view(1, stride)
    want_s = torch.empty((rows, K), dtype=torch.float32, device=dev)
    want_i = torch.empty((rows, K), dtype=torch.int64, device=dev)
    _cabi.check(_cabi.load().alive_knn_merge_records(gathered.data_ptr(), stride, world, rows, K, want_s.data_ptr(),
                                                     want_i.data_ptr(), torch.cuda.current_stream().cuda_stream), "merge_records")
    same = (idx == want_i).all(dim=1)
    # a differing row is excused only by a similarity tie within 1e-6: compare the exact scores of OUR frames
    # (computed by the exact scan as well: every frame's exact score is unique per (query, frame)) position-wise.
    # Ours are not returned with scores here, so recompute them from the raw frames in float64.
    n_tie = 0
    if not bool(same.all()):
        bad = (~same).nonzero().reshape(-1)
        qn = torch.nn.functional.normalize(q.transpose(1, 2).reshape(rows, D)[bad].double(), dim=1)
        mine = torch.zeros((bad.numel(), K), dtype=torch.float64, device=dev)
        lo, hi = lib.row_base, lib.row_base + lib.n
        own = (idx[bad] >= lo) & (idx[bad] < hi)
        loc = (idx[bad] - lo).clamp(0, max(lib.n - 1, 0))
        fr = torch.nn.functional.normalize(lib.raw[loc].double(), dim=2)                 # [bad, K, D]
        mine = torch.where(own, (fr * qn[:, None, :]).sum(dim=2), mine)
        if world > 1:
            dist.all_reduce(mine)
        n_tie = int(((mine.float() - want_s[bad]).abs() <= 1e-6).all(dim=1).sum())
    # (b) features: ((r0 + r1) + r2) + r3 in descending-score order, / k  (alpha = 0)
    lo, hi = lib.row_base, lib.row_base + lib.n
    own = (idx >= lo) & (idx < hi)
    loc = (idx - lo).clamp(0, max(lib.n - 1, 0))
    picked = torch.where(own[:, :, None], lib.raw[loc] if lib.n > 0 else torch.zeros((rows, K, D), device=dev),
                         torch.zeros((), device=dev))
    if world > 1:
        dist.all_reduce(picked)
    acc = picked[:, 0].clone()
    for j in range(1, K):
        acc = acc + picked[:, j]
    feat_rows = int((out_rows == acc / K).all(dim=1).sum())
    n_same = int(same.sum())
    return {"rows": rows, "index_exact": n_same, "tie_excused": n_tie, "features_bit_exact_rows": feat_rows,
            "checker": "exhaustive fp64 scan of every shard (alive_knn_exact) merged over ranks; sequential float32 mean of the raw frames",
            "ok": n_same + n_tie == rows and feat_rows == rows}


def measure(env, workload: str, steps: int, warmup: int, headline: bool, clustered=None):
    """One workload on env.world GPUs -> the dict of its numbers (rank 0; None elsewhere)."""
    import ctypes
    import torch
    import torch.distributed as dist
    from alive_vc_b200 import _cabi, matching as M
    from alive_vc_b200.sharded import CudaShardBackend, ShardedLibrary, shard_bounds

    world, rank, dev, args, peaks, variant = env.world, env.rank, env.dev, env.args, env.peaks, env.variant
    B, T, N, desc = WORKLOADS[workload]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident data: the (sharded) packed library, device queries, pinned host queries ----
    g = torch.Generator(device=dev).manual_seed(args.seed)
    sharded = None
    batched = workload == "cfg5"
    if batched:
        if world > 1:
            per = B // world
            my_items = list(range(rank * per, (rank + 1) * per))
        else:
            my_items = list(range(B))
        # all of this rank's per-speaker libraries packed back to back: ONE pipeline launch per step
        lib = M.alloc_packed(len(my_items) * N, D, dev, items=len(my_items))     # (no second plane: 147 GB fill the GPU)
        for i, b in enumerate(my_items):
            for c0 in range(0, N, 250_000):
                c1 = min(N, c0 + 250_000)
                gg = torch.Generator(device=dev).manual_seed((args.seed + 17 * (b + 1)) * 1_000_003 + c0 // 250_000)
                x = torch.randn(D, c1 - c0, device=dev, generator=gg)
                M.pack_into(lib, i * N + c0, x)
                del x
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
        if clustered is not None:
            src_dev, ref = build_clustered(T, N, clustered["clusters"], clustered["noise"], args.seed, dev)
            lib = M.pack_library(ref, fmt=clustered.get("fmt", "auto"))       # "auto": the pack-time probe chooses
            del ref
        else:
            lib = build_library(lo, hi, args.seed, dev)
            src_dev = torch.randn(B, D, T, device=dev, generator=g)
        sharded = ShardedLibrary(CudaShardBackend(lib, "screen", variant), lib.n, lo, N, None,
                                 peer_memory=(args.exchange == "peer"))
        if world > 1:
            dist.broadcast(src_dev, 0)

        def step(src):
            return sharded.match(src, K, 0.0)
        units_per_step = B * T
        n_local = hi - lo
        scaling = "strong"
        exchange = "one fused merge + peer-memory (NVLink, CUDA IPC) gather kernel" if sharded.peers is not None else \
            "merge + reduce-scatter of the owned rows + all-gather of the result (NCCL)"
        parallelism = f"library rows sharded x{world}" + ("" if world == 1 else f", one all-gather of the top-k records, {exchange}")
    lib_two_planes = lib.lo is not None
    streaming = workload in GRAPH_WORKLOADS and world == 1 and not args.no_graph and not batched
    if streaming:
        # fixed-shape chunks (inference.py / realtime_inference.py call the match once per chunk with
        # the same T): pre-allocated buffers, the whole pipeline replayed as one CUDA graph
        streamer = M.StreamingMatcher(lib, T, K, 0.0, batch=B, mode="screen", variant=variant)

        def step(src):                                             # noqa: F811
            return streamer(src)
    latency_workload = workload in LATENCY_WORKLOADS
    barrier()

    # ---- device-resident timing ----
    for _ in range(warmup):
        step(src_dev)
    barrier()
    M.search_events = []
    launches0 = M.launch_count
    sampler = ClockSampler(env.local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lat = []
    barrier()
    e0.record()
    if latency_workload:
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        evs[0].record()
        for i in range(steps):
            step(src_dev)
            evs[i + 1].record()
    else:
        for _ in range(steps):
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
        for _ in range(min(steps, 50)):
            sharded.match(src_dev, K, 0.0)
        torch.cuda.synchronize()
        search_ms = [a.elapsed_time(b) for a, b in M.search_events]
    M.search_events = None
    if latency_workload:
        lat = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(steps))
    info = M.last_info
    fallback = info.fallback_queries() if info is not None else 0
    exact_scan = info.exact_scan_queries() if info is not None else 0
    if world > 1:
        tmax = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_total = float(tmax.item())
    ms_per_step = ms_total / steps
    value = units_per_step / (ms_per_step * 1e-3)

    # ---- end-to-end: host buffers through the public API ----
    e2e_api = None

    def e2e_finish():
        pass
    scattered = world > 1 and not batched and B == 1
    if scattered:
        # every query byte crosses PCIe once: this rank's T/world slice in, its slice of the result out
        q_lo, q_hi = shard_bounds(T, world, rank)
        src_host = torch.empty((1, D, q_hi - q_lo), dtype=torch.float32).pin_memory()
        src_host.copy_(src_dev[:, :, q_lo:q_hi])
        out_rows_host = torch.empty((1, q_hi - q_lo, D), dtype=torch.float32).pin_memory()
        from alive_vc_b200.lifecycle import HostPipeline
        # the same double-buffered host feeding as on one GPU; the collectives of the scattered match run on the
        # main stream between the copies
        pipeline = HostPipeline(lib, 1, q_hi - q_lo, K, 0.0, match_fn=lambda s: sharded.match(
            s, K, 0.0, scattered=True, t_total=T).transpose(1, 2))

        def e2e_step():
            pipeline.step(src_host, out_rows_host)
        e2e_finish = pipeline.drain
        io_bytes = T * D * 4                       # whole job, all ranks together
    else:
        src_host = torch.empty(src_dev.shape, dtype=torch.float32).pin_memory()
        src_host.copy_(src_dev)
        # the result is a transposed view of a contiguous [B,T,D] block (like the reference's); the pinned
        # host buffer has the same strides, so the device->host read is one plain memcpy
        _b, _d, _t = src_dev.shape
        out_rows_host = torch.empty((_b, _t, _d), dtype=torch.float32).pin_memory()
        out_host = out_rows_host.transpose(1, 2)
        if world == 1 and not latency_workload:
            # throughput workloads: lifecycle.HostPipeline - the same three operations per step, double-buffered so
            # that the copies of neighbouring steps overlap the match (every step still moves its own bytes both ways)
            from alive_vc_b200.lifecycle import HostPipeline
            pipeline = HostPipeline(lib, _b, _t, K, 0.0, "screen", variant)
            e2e_api = "lifecycle.HostPipeline.step"

            def e2e_step():
                pipeline.step(src_host, out_rows_host)
            e2e_finish = pipeline.drain
        else:
            def e2e_step():
                # the streaming matcher copies the pinned host chunk straight into its static input buffer
                s = src_host if streaming else src_host.to(dev, non_blocking=True)
                o = step(s)
                out_host.copy_(o, non_blocking=True)
        io_bytes = src_host.numel() * 4 * (world if batched else 1)
    for _ in range(min(warmup, 3)):
        e2e_step()
    e2e_finish()
    barrier()
    e0.record()
    for _ in range(steps):
        e2e_step()
    e2e_finish()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        tmax = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_ms = float(tmax.item())
    e2e_value = units_per_step / (e2e_ms / steps * 1e-3)

    # latency workloads: the synchronous host-to-host chunk time of the realtime loop
    # (realtime_inference.py:158-176) - pinned buffers, ONE graph (H2D + pipeline + D2H), one event wait
    host_lat = []
    if latency_workload and world == 1 and not args.no_graph:
        from alive_vc_b200.lifecycle import HostStreamingMatcher
        hm = HostStreamingMatcher(lib, T, K, 0.0, batch=B, mode="screen", variant=variant)
        chunk = src_host.clone()
        for _ in range(20):
            hm(chunk)
        for _ in range(min(steps, 2000)):
            t0 = time.perf_counter()
            hm(chunk)
            host_lat.append((time.perf_counter() - t0) * 1e3)
        host_lat.sort()
        del hm

    # ---- parity of what was just timed (every rank takes part; rank 0 reports) ----
    parity = parity_block(env, sharded, lib, src_dev, batched)
    if world > 1:
        flag = torch.tensor([1 if parity["ok"] else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        parity["ok"] = bool(flag.item())

    res = None
    if rank == 0:
        # roofline of the dominant kernel: algorithmic flops of ONE alive_knn_search launch
        # (2 * T * N_local * D) over its average CUDA-event duration
        flops_per_launch = 2.0 * (lib.items if batched else B) * T * n_local * D
        avg_search_ms = sum(search_ms) / max(1, len(search_ms))
        if batched:
            _plan = _cabi.Plan()
            _cabi.check(_cabi.load().alive_knn_plan_batched(lib.items, T, n_local, D, M._num_sms(dev), variant,
                                                            ctypes.byref(_plan)), "alive_knn_plan_batched")
            search_kernel = "knn_search_resident_kernel" if resident_kernel(_plan) else "knn_search_kernel"
        else:
            _plan = M.make_plan(B * T, n_local, D, dev, variant)
            search_kernel = ("knn_search_skinny_kernel" if _plan.kernel == 1 else
                             "knn_search_resident_kernel" if resident_kernel(_plan) else "knn_search_kernel")
        if latency_workload:
            bytes_per_launch = float(n_local) * D * 2
            achieved = bytes_per_launch / (avg_search_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "traffic": load_traffic(workload, world),
                    "kernel": search_kernel, "avg_kernel_ms": avg_search_ms, "peak_source": peaks["source"]}
        else:
            achieved = flops_per_launch / (avg_search_ms * 1e-3) / 1e12
            # conservative denominator: the burst cuBLAS figure, even though the kernel runs
            # back-to-back under the power cap (frac_of_sustained is reported beside it)
            peak = peaks["bf16_burst"]
            roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": load_traffic(workload, world) if clustered is None else None,
                    "kernel": search_kernel, "avg_kernel_ms": avg_search_ms,
                    "kernel_share_of_step": avg_search_ms / ms_per_step,
                    "peak_kind": "burst", "peak_source": peaks["source"],
                    "frac_of_burst": achieved / peaks["bf16_burst"],
                    "frac_of_sustained": achieved / peaks["bf16_sustained"]}
        res = {
            "value": value, "unit": "query_frames/s", "ms_per_step": ms_per_step, "steps": steps, "warmup": warmup,
            "scaling": scaling,
            "config": {"workload": desc if clustered is None else
                       f"clustered: cfg1 shape (T={T}, N={N}), {clustered['clusters']} clusters, noise {clustered['noise']}",
                       "B": B, "T": T, "N": N, "D": D, "k": K, "parallelism": parallelism,
                       "l2": "library (bf16 %.1f GB per GPU) is far larger than L2, no flush needed"
                             % (n_local * D * 2 / 1e9) if n_local * D * 2 > 256e6 else
                             "library smaller than 2x L2: numbers are warm-L2 steady state of a resident library",
                       "variant": variant,
                       "planes": f"{'fp16' if lib.format == 1 else 'bf16'} ({'two planes' if lib_two_planes else 'one plane'})"
                                 + (f", requested {clustered.get('fmt', 'auto')}" if clustered is not None else ""),
                       "fallback_queries_last_step": fallback, "exact_scan_queries_last_step": exact_scan,
                       "api": "StreamingMatcher (one CUDA graph per chunk)" if streaming else
                              ("match_packed on pack_libraries (one launch for all speakers)" if batched else "ShardedLibrary.match")},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "query_frames/s", "h2d_bytes_per_step": io_bytes,
                    "d2h_bytes_per_step": io_bytes, "ms_per_step": e2e_ms / steps,
                    "path": "per rank: H2D of its T/N query slice -> all-gather over NVLink -> match -> D2H of its T/N result slice"
                            if scattered else ("H2D of the queries -> match -> D2H of the result, every step; double-buffered "
                                               "(lifecycle.HostPipeline: the copies of neighbouring steps overlap the match)"
                                               if e2e_api else "H2D of the queries -> match -> D2H of the result")},
            "gpu_launches": launches,
            "roofline": roof,
            "parity": parity,
        }
        if lat:
            res["latency_ms"] = {"p50": lat[len(lat) // 2], "p99": lat[min(len(lat) - 1, int(len(lat) * 0.99))],
                                 "min": lat[0]}
        if host_lat:
            res["host_chunk_latency_ms"] = {"p50": host_lat[len(host_lat) // 2],
                                            "p99": host_lat[min(len(host_lat) - 1, int(len(host_lat) * 0.99))],
                                            "min": host_lat[0], "timer": "perf_counter around one blocking call"}
        if headline:
            g_roof = gather_roofline(env, lib, src_dev, B, T) if (not batched and world == 1) else None
            if g_roof is not None:
                res["roofline_gather"] = g_roof
    # ---- release everything this workload held (the next one may need all of HBM) ----
    if sharded is not None:
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sharded.close()
    del lib, src_dev, sharded, info
    if streaming:
        del streamer
    M.clear_pack_cache()
    M.last_info = None
    torch.cuda.empty_cache()
    return res


def resident_kernel(plan) -> bool:
    """Which K2 instance alive_knn_search launches for this plan (mirrors launch_search in search_sm100.cu): CTA pairs at
    d = 768 that visit at least two library tiles per unit keep half of their query block resident."""
    return (plan.kernel == 0 and plan.ctas_per_unit == 2 and plan.d == 768 and plan.tiles_per_segment >= 2
            and os.environ.get("ALIVE_KNN_RESIDENT", "1") != "0")


def _settle(seconds: float = 2.0):
    """The standalone K1/K4 measurements ("a kernel timed alone", burst HBM peak) follow a loop that holds the GPU at its
    power cap; the SM clock stays low for a while after it (1.3 instead of 1.9 GHz) and these latency-sensitive
    kernels then measure 20-25 % lower.  Let the clocks come back before timing them."""
    import torch
    torch.cuda.synchronize()
    time.sleep(seconds)


def gather_roofline(env, lib, src_dev, B, T):
    """K4 alone (north_star: "fraction of HBM bandwidth for the gather"): the standalone gather-mean kernel on this
    step's own neighbour indices; in the pipeline the same arithmetic is fused into finish_kernel.  Algorithmic
    bytes per query frame: k*D*4 read + D*4 written = 15,360 B (alpha = 0: the query row is not read)."""
    import torch
    from alive_vc_b200 import matching as M

    dev, peaks = env.dev, env.peaks
    rows = B * T
    _, g_idx, _ = M.run_match(src_dev, lib, K, 0.0, "screen", env.variant, want_out=False)
    ws, off = M.last_info._workspace, M.last_info._offsets
    q_view = M.PackedFrames(n=rows, d=D, raw=ws[off[0]: off[0] + rows * D * 4].view(torch.float32).view(rows, D),
                            norms=ws[off[1]: off[1] + rows * 4].view(torch.float32), packed=None, err=None, stats=None)
    g_out = torch.empty((rows, D), dtype=torch.float32, device=dev)
    _settle()
    for _ in range(10):
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
    g_bytes = rows * (K * D * 4 + D * 4)
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
    return {"bound": "hbm", "kernel": "gather_mean_warp_kernel", "achieved": g_bytes / (g_ms * 1e-3) / 1e9,
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": g_bytes / (g_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
            "avg_kernel_ms": g_ms, "bytes_per_query_frame": g_bytes // rows,
            "torch_index_select_gbs": sel_gbs,
            "note": "standalone K4 on this step's indices; random 3 KB rows of the raw library; "
                    "torch_index_select_gbs = torch's row gather of the same rows (read + write), "
                    "the practical ceiling of this access pattern"}


def pack_roofline(env):
    """K1 alone (once per library, generate_voice_library.py / load time): the pack kernel on a fresh
    channel-major [D, n] chunk far larger than L2.  Algorithmic bytes per frame: D*(4 read + 4 raw + 2 packed + 2 second
    plane) + 12 (norm, err, err2) = 9,228 B with the second bf16 plane the product stores by default for a single
    library (7,688 B without it, reported beside)."""
    import torch
    from alive_vc_b200 import matching as M

    dev, peaks = env.dev, env.peaks
    pn = 250_000
    px = torch.randn(D, pn, device=dev)
    px_rows = px.t().contiguous()
    reps = 10

    _settle()

    def timed(dst, view):
        for _ in range(4):
            M.pack_into(dst, 0, view)
        torch.cuda.synchronize()
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        for _ in range(reps):
            M.pack_into(dst, 0, view)
        pe1.record()
        torch.cuda.synchronize()
        return pe0.elapsed_time(pe1) / reps

    out = {}
    for refine in (True, False):
        pdst = M.alloc_packed(pn, D, dev, refine=refine)
        per_frame = D * (12 if refine else 10) + (12 if refine else 8)
        ms_cm = timed(pdst, px)
        ms_rm = timed(pdst, px_rows.t())       # the same frames as a row-major producer hands them over (pack_rows)
        out[refine] = {"bytes_per_frame": per_frame,
                       "channel_major": {"kernel": "pack_cm2_kernel", "avg_kernel_ms": ms_cm,
                                         "achieved": pn * per_frame / (ms_cm * 1e-3) / 1e9,
                                         "frac": pn * per_frame / (ms_cm * 1e-3) / 1e9 / peaks["hbm_gbs"]},
                       "row_major": {"kernel": "pack_rm_kernel", "avg_kernel_ms": ms_rm,
                                     "achieved": pn * per_frame / (ms_rm * 1e-3) / 1e9,
                                     "frac": pn * per_frame / (ms_rm * 1e-3) / 1e9 / peaks["hbm_gbs"]}}
        del pdst
    # torch's own transposing copy of the same chunk ([768, n] -> [n, 768], read + write 4 B per element):
    # what a library kernel gets out of the channel-major access pattern
    t_out = torch.empty((pn, D), dtype=torch.float32, device=dev)
    for _ in range(2):
        t_out.copy_(px.t())
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for _ in range(reps):
        t_out.copy_(px.t())
    pe1.record()
    torch.cuda.synchronize()
    t_gbs = 2.0 * pn * D * 4 / (pe0.elapsed_time(pe1) / reps * 1e-3) / 1e9
    del t_out, px, px_rows
    torch.cuda.empty_cache()
    main = out[True]
    return {"bound": "hbm", "kernel": "pack_cm2_kernel", "achieved": main["channel_major"]["achieved"],
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": main["channel_major"]["frac"],
            "avg_kernel_ms": main["channel_major"]["avg_kernel_ms"], "bytes_per_frame": main["bytes_per_frame"], "frames": pn,
            "row_major": main["row_major"],
            "without_second_plane": {"bytes_per_frame": out[False]["bytes_per_frame"],
                                     "channel_major": out[False]["channel_major"], "row_major": out[False]["row_major"]},
            "torch_transpose_copy_gbs": t_gbs,
            "note": "standalone K1 on a channel-major [768, 250k] fp32 chunk (1.9 GB read per launch, no L2 reuse), writing raw "
                    "fp32 rows + both bf16 planes + norms; row_major = the same frames as [250k, 768] rows"}


def summary_of(res):
    """the per-workload entry of `workloads`"""
    keep = ("value", "unit", "ms_per_step", "steps", "latency_ms", "host_chunk_latency_ms", "gpu_launches", "parity")
    out = {k: res[k] for k in keep if k in res}
    out["e2e"] = res["e2e"]
    r = res["roofline"]
    out["roofline"] = {k: r[k] for k in ("bound", "kernel", "achieved", "peak", "unit", "frac", "avg_kernel_ms",
                                           "frac_of_sustained", "kernel_share_of_step") if k in r}
    out["config"] = {k: res["config"][k] for k in ("workload", "B", "T", "N", "api", "planes", "fallback_queries_last_step",
                                                    "exact_scan_queries_last_step")}
    out["clocks"] = res["clocks"]
    return out


def gpu_arm(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry

    env = Env()
    env.args = args
    env.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    env.rank = rank = int(os.environ.get("RANK", "0"))
    env.local_rank = local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1 (one process per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    env.dev = dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        entry.build()          # no-op when the in-tree .so is current (it travels with the snapshot)
    if world > 1:
        dist.barrier()         # nobody loads the library before rank 0 has (re)built it
    env.peaks = load_peaks()
    env.variant = args.variant
    from alive_vc_b200 import matching as _M
    _M.SCREEN_FORMAT = args.format

    res = measure(env, args.workload, args.steps, args.warmup, headline=True)
    ok = True
    line = None
    if rank == 0:
        ok = res["parity"]["ok"]
        line = {
            "metric": "query_frames_per_sec_matched_k4", "value": res["value"], "unit": "query_frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": res["scaling"], "vs_baseline": None,
            "dtype": "f16" if args.format == "fp16" else "bf16",
            "data": "synthetic", "config": res["config"], "clocks": res["clocks"], "e2e": res["e2e"],
            "gpu_launches": res["gpu_launches"], "roofline": res["roofline"], "parity": res["parity"],
        }
        for key in ("latency_ms", "host_chunk_latency_ms", "roofline_gather"):
            if key in res:
                line[key] = res[key]
        if world == 1:
            line["roofline_pack"] = pack_roofline(env)
            line["cpu_baseline"] = None if args.no_cpu else run_cpu_arm(args.workload, 3, 1)
            if args.torch_eager:
                line["torch_eager_gpu"] = run_torch_eager_gpu(args.workload, dev)
    # ---- the other BASELINE configs, same process, same measurement (default N=1 run of the headline) ----
    if world == 1 and args.workload == "cfg4" and not args.no_extras:
        extras = {}
        for w in EXTRA_WORKLOADS:
            try:
                r = measure(env, w, DEFAULT_STEPS[w], 3, headline=False)
            except Exception as e:       # noqa: BLE001 - an extra must never take the headline down with it
                extras[w] = {"error": f"{type(e).__name__}: {e}"[:300]}
                torch.cuda.empty_cache()
                continue
            extras[w] = summary_of(r)
            ok = ok and r["parity"]["ok"]
            if not args.no_cpu:
                extras[w]["cpu_baseline"] = run_cpu_arm(w, 2, 1)
        # clustered libraries: cfg1's shape, 100 clusters of ~1000 near-identical frames
        for noise, fmt in ((0.2, "auto"), (0.5, "auto"), (0.2, "bf16")):
            name = f"clustered_noise{noise}" + ("" if fmt == "auto" else f"_{fmt}")
            try:
                r = measure(env, "cfg1", 20, 3, headline=False, clustered={"clusters": 100, "noise": noise, "fmt": fmt})
                ent = summary_of(r)
                ent["slowdown_vs_random_cfg1"] = (r["ms_per_step"] / extras["cfg1"]["ms_per_step"]
                                                  if "ms_per_step" in extras.get("cfg1", {}) else None)
                extras[name] = ent
                ok = ok and r["parity"]["ok"]
            except Exception as e:       # noqa: BLE001
                extras[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
                torch.cuda.empty_cache()
        line["workloads"] = extras
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not ok:
        print("bench.py: PARITY FAILURE (see the `parity` entries above)", file=sys.stderr, flush=True)
        sys.exit(1)


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
    ap.add_argument("--no-extras", action="store_true",
                    help="default cfg4 run at N=1: skip the other BASELINE configs (`workloads` in the JSON line)")
    ap.add_argument("--exchange", default="peer", choices=["nccl", "peer"],
                    help="multi-GPU row exchange: one fused merge+gather kernel over CUDA-IPC peer memory (falls back "
                         "to NCCL when the mapping fails), or always the NCCL reduce-scatter/all-gather")
    ap.add_argument("--format", default="bf16", choices=["fp16", "bf16"],
                    help="16-bit format of the packed planes (tensor-core operands) of the synthetic BASELINE libraries: bf16 "
                         "(the brief's format and the faster one under the power cap) or IEEE fp16 (8x finer rounding; what "
                         "pack_library's probe picks for clustered libraries - the `clustered` entries use that probe)")
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
