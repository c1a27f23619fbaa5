"""The multi-GPU choreography of alive_vc_b200.sharded (shard bounds, all-gather of local
top-k, merge, zero-padded row exchange, mean+blend) exercised on CPU with world_size=2/3
`gloo` process groups.  The compute steps are supplied by an ORACLE-based backend defined
here (test infrastructure); the product backend is CUDA-only."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """numpy restatement of the per-rank steps (uses oracle/knn_oracle.py)."""

    def __init__(self, local_ref_dn: np.ndarray, row_base: int):
        self.ref = local_ref_dn            # [D, n_local]
        self.row_base = row_base

    def pack_queries(self, source):
        return source.numpy()              # [B,D,T]

    def local_topk(self, q, k):
        from oracle import knn_oracle as O
        B, D, T = q.shape
        qq = np.swapaxes(q, 1, 2).reshape(1, B * T, D).swapaxes(1, 2)     # [1,D,B*T]
        scores = O.cosine_scores_np(qq, self.ref[None])
        val, idx = O.topk_desc_np(scores, k)
        return torch.from_numpy(val[0]), torch.from_numpy(idx[0] + self.row_base)

    def merge(self, scores, idx, k):
        r, t, kk = scores.shape
        s = scores.permute(1, 0, 2).reshape(t, r * kk).numpy()
        i = idx.permute(1, 0, 2).reshape(t, r * kk).numpy()
        top_s = np.empty((t, k), np.float32)
        top_i = np.empty((t, k), np.int64)
        for q in range(t):
            key = np.where(i[q] < 0, -np.inf, s[q])
            order = np.lexsort((i[q], -key))
            top_s[q], top_i[q] = s[q][order[:k]], i[q][order[:k]]
        return torch.from_numpy(top_s), torch.from_numpy(top_i)

    def gather_rows(self, top_idx):
        t, k = top_idx.shape
        D, n = self.ref.shape
        rows = np.zeros((t, k, D), np.float32)
        loc = top_idx.numpy() - self.row_base
        own = (loc >= 0) & (loc < n)
        rows[own] = self.ref.T[loc[own]]
        return torch.from_numpy(rows)

    def mean_blend(self, rows, q, alpha, row0=0, out=None):
        r = rows.numpy()
        acc = r[:, 0].copy()
        for j in range(1, r.shape[1]):
            acc = (acc + r[:, j]).astype(np.float32)
        res = (acc / np.float32(r.shape[1])).astype(np.float32)
        B, D, T = q.shape
        qrows = np.swapaxes(q, 1, 2).reshape(B * T, D)
        out = (res * np.float32(1 - alpha)).astype(np.float32) + (qrows * np.float32(alpha)).astype(np.float32)
        return torch.from_numpy(out.astype(np.float32))


def _worker(rank, world, port, n_total, T, k, alpha, B):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from alive_vc_b200.sharded import ShardedLibrary, shard_bounds
        from oracle import knn_oracle as O
        rng = np.random.default_rng(77)
        src = rng.standard_normal((B, 768, T), dtype=np.float32)
        ref = rng.standard_normal((1, 768, n_total), dtype=np.float32)
        lo, hi = shard_bounds(n_total, world, rank)
        lib = ShardedLibrary(OracleBackend(ref[0][:, lo:hi], lo), hi - lo, lo, n_total)
        out, idx = lib.match(torch.from_numpy(src), k=k, alpha=alpha, return_indices=True)
        ref_b = np.broadcast_to(ref, (B,) + ref.shape[1:])
        want_out, want_idx, _ = O.match_features_np(src, ref_b, k, alpha, True)
        assert np.array_equal(idx.numpy(), want_idx), "sharded indices differ from the single-library oracle"
        assert np.array_equal(out.numpy(), want_out), "sharded features are not bit-identical"
        assert tuple(out.shape) == (B, 768, T) and tuple(out.stride()) == (T * 768, 1, 768)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,n_total,T,k,alpha,B", [
    (2, 1001, 37, 4, 0.0, 1),
    (2, 300, 16, 4, 0.25, 2),
    (3, 10, 5, 4, 0.0, 1),        # shards of 4/3/3 frames: some ranks hold fewer than k frames
    (3, 2, 4, 2, 0.5, 1),         # shards of 1/1/0 frames: one rank holds nothing
])
def test_sharded_match_equals_single_library(world, n_total, T, k, alpha, B):
    mp.spawn(_worker, args=(world, _free_port(), n_total, T, k, alpha, B), nprocs=world, join=True)


def _worker_scattered(rank, world, port, n_total, T, k, alpha):
    """scattered form: every rank passes its slice of the query frames and gets its slice of the result"""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from alive_vc_b200.sharded import ShardedLibrary, shard_bounds
        from oracle import knn_oracle as O
        rng = np.random.default_rng(78)
        src = rng.standard_normal((1, 768, T), dtype=np.float32)
        ref = rng.standard_normal((1, 768, n_total), dtype=np.float32)
        lo, hi = shard_bounds(n_total, world, rank)
        q_lo, q_hi = shard_bounds(T, world, rank)
        lib = ShardedLibrary(OracleBackend(ref[0][:, lo:hi], lo), hi - lo, lo, n_total)
        mine = torch.from_numpy(np.ascontiguousarray(src[:, :, q_lo:q_hi]))
        out, idx = lib.match(mine, k=k, alpha=alpha, return_indices=True, scattered=True, t_total=T)
        want_out, want_idx, _ = O.match_features_np(src, ref, k, alpha, True)
        assert tuple(out.shape) == (1, 768, q_hi - q_lo) and tuple(idx.shape) == (1, q_hi - q_lo, k)
        assert np.array_equal(idx.numpy(), want_idx[:, q_lo:q_hi]), "scattered indices differ from the oracle"
        assert np.array_equal(out.numpy(), want_out[:, :, q_lo:q_hi]), "scattered features are not bit-identical"
        # a slice of the wrong length is refused before anything is exchanged
        if q_hi - q_lo > 1:
            with pytest.raises(RuntimeError, match="must pass its"):
                lib.match(mine[:, :, 1:], k=k, alpha=alpha, scattered=True, t_total=T)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total,T,k,alpha", [
    (2, 1001, 37, 4, 0.0),        # 19 + 18 query frames: the shorter slice is padded for the all-gather
    (3, 300, 16, 4, 0.25),
    (3, 10, 2, 4, 0.0),           # fewer query frames than ranks: one rank passes an empty slice
])
def test_sharded_scattered_match(world, n_total, T, k, alpha):
    mp.spawn(_worker_scattered, args=(world, _free_port(), n_total, T, k, alpha), nprocs=world, join=True)
