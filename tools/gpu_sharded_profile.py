"""Phase breakdown of ShardedLibrary.match under torchrun (CUDA events between phases)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from alive_vc_b200.sharded import CudaShardBackend, ShardedLibrary, shard_bounds  # noqa: E402
import bench                                                   # noqa: E402


def main():
    rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    T, N = 10_000, int(os.environ.get("N_TOTAL", 10_000_000))
    lo, hi = shard_bounds(N, world, rank)
    lib = bench.build_library(lo, hi, 7, dev)
    sh = ShardedLibrary(CudaShardBackend(lib, "screen", 0), lib.n, lo, N, None)
    g = torch.Generator(device=dev).manual_seed(1)
    src = torch.randn(1, 768, T, device=dev, generator=g)
    dist.broadcast(src, 0)
    marks = []

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append((name, e))

    be = sh.backend
    orig = {n: getattr(be, n) for n in ("local_topk", "merge", "gather_rows", "mean_blend")}

    def wrap(name):
        def f(*a, **k):
            mark("pre_" + name)
            r = orig[name](*a, **k)
            mark("post_" + name)
            return r
        return f
    for n in orig:
        setattr(be, n, wrap(n))
    for it in range(4):
        marks.clear()
        dist.barrier(); torch.cuda.synchronize()
        mark("start")
        out = sh.match(src, 4, 0.0)
        mark("end")
        torch.cuda.synchronize()
    if rank == 0:
        prev = marks[0]
        for name, e in marks[1:]:
            print(f"{prev[0]:>18s} -> {name:<18s} {prev[1].elapsed_time(e):8.3f} ms")
            prev = (name, e)
        print("total", marks[0][1].elapsed_time(marks[-1][1]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
