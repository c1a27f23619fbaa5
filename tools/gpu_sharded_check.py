"""Run under torchrun on N GPUs: the row-sharded library over NCCL - the peer path (one all-gather of the top-k
records + ONE fused merge/gather kernel over CUDA-IPC peer memory), the NCCL row exchange, and the scattered form
(each rank passes its slice of the queries and gets its slice of the result) - must give bit-identical results to
the single-GPU path.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/gpu_sharded_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import alive_vc_b200 as A                                        # noqa: E402
from alive_vc_b200.sharded import ShardedLibrary, shard_bounds   # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for (B, T, N, k, alpha) in [(1, 500, 200_003, 4, 0.0), (2, 64, 50_000, 4, 0.25), (1, 33, 5, 4, 0.0), (1, 24, 50_000, 4, 0.0),
                                (1, 300, 160_000, 4, 0.0), (1, 2003, 1_000_000, 4, 0.0)]:
        g = torch.Generator(device=dev).manual_seed(123)
        src = torch.randn(B, 768, T, device=dev, generator=g)
        ref = torch.randn(1, 768, N, device=dev, generator=g)
        if N == 160_000:
            # tight clusters: nothing certifies, every shard runs its collect pass
            cent = torch.randn(768, 40, device=dev, generator=g)
            ref = (cent[:, torch.randint(0, 40, (N,), device=dev, generator=g)] + 0.2 * ref[0])[None]
            src = (cent[:, torch.randint(0, 40, (T,), device=dev, generator=g)] + 0.2 * src[0])[None]
        want, widx = A.match_features(src, ref.expand(B, -1, -1), k, alpha, return_indices=True)
        same = True
        for peer in (False, True):
            lib = ShardedLibrary.from_full(ref, mode="auto", peer_memory=peer)
            if peer and lib.peers is None and rank == 0:
                print(f"   (peer mapping unavailable: {lib.peer_error}; NCCL exchange used)", flush=True)
            out, idx = lib.match(src, k, alpha, return_indices=True)
            good = torch.equal(out, want) and torch.equal(idx, widx)
            if B == 1:
                lo, hi = shard_bounds(T, world, rank)
                o2, i2 = lib.match(src[:, :, lo:hi], k, alpha, return_indices=True, scattered=True, t_total=T)
                good = good and torch.equal(o2, want[:, :, lo:hi]) and torch.equal(i2, widx[:, lo:hi])
            if not good:
                print(f"   MISMATCH on rank {rank} with peer_memory={peer}", flush=True)
            same = same and good
            torch.cuda.synchronize()
            dist.barrier()
            lib.close()
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"sharded x{world} B={B} T={T} N={N} k={k} alpha={alpha}: bit-identical={bool(flag.item())}",
                  flush=True)
        ok &= bool(flag.item())
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
