"""Where one step of the row-sharded cfg4 match goes, on rank 0 under torchrun: every kernel (ours and NCCL's)
with its device time, from torch.profiler (CUPTI) - ncu cannot wrap a multi-rank command.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        tools/gpu_sharded_profile.py          # writes gpurun_out/launches_cfg4_n8_r02.csv
"""
import csv
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from alive_vc_b200.sharded import CudaShardBackend, ShardedLibrary, shard_bounds  # noqa: E402
import bench                                                   # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    T, N = 10_000, int(os.environ.get("N_TOTAL", 10_000_000))
    exchange = os.environ.get("EXCHANGE", "peer")
    lo, hi = shard_bounds(N, world, rank)
    lib = bench.build_library(lo, hi, 7, dev)
    sh = ShardedLibrary(CudaShardBackend(lib, "screen", 0), lib.n, lo, N, None, peer_memory=(exchange == "peer"))
    g = torch.Generator(device=dev).manual_seed(1)
    src = torch.randn(1, 768, T, device=dev, generator=g)
    dist.broadcast(src, 0)
    q_lo, q_hi = shard_bounds(T, world, rank)
    modes = {"replicated": lambda: sh.match(src, 4, 0.0),
             "scattered": lambda: sh.match(src[:, :, q_lo:q_hi], 4, 0.0, scattered=True, t_total=T)}
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    for mode, fn in modes.items():
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        step_ms = e0.elapsed_time(e1) / 5
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            fn()
            torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            rows = []
            for ev in prof.events():
                if ev.device_type == torch.autograd.DeviceType.CUDA:
                    rows.append((ev.time_range.start, ev.name, ev.time_range.elapsed_us()))
            rows.sort()
            path = os.path.join(out_dir, f"launches_cfg4_n{world}_{mode}_{exchange}_r02.csv")
            with open(path, "w", newline="") as f:
                w = csv.writer(f)
                w.writerow(["order", "kernel", "duration_us", f"# one step of ShardedLibrary.match ({mode} output, {exchange} exchange) "
                            f"on rank 0 of {world}; step = {step_ms:.3f} ms by CUDA events over 5 steps"])
                for i, (_, name, us) in enumerate(rows):
                    w.writerow([i, name[:120], f"{us:.1f}"])
            total = sum(r[2] for r in rows)
            print(f"[{mode}/{exchange}] step {step_ms:.3f} ms; {len(rows)} device activities, {total / 1e3:.3f} ms busy -> {path}", flush=True)
            top = sorted(rows, key=lambda r: -r[2])[:12]
            for _, name, us in top:
                print(f"    {us:10.1f} us  {name[:100]}", flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    sh.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
