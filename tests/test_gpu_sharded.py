"""The row-sharded CUDA path under test on ONE device.

R in-process ranks (alive_vc_b200.sharded.ThreadComm: one thread per rank, collectives through a barrier) drive
the REAL product backend - CudaShardBackend: alive_knn_match on the shard, one record per rank, alive_knn_merge_gather
over every rank's raw shard (the peer path; same-process pointers stand in for the CUDA-IPC mappings) or
alive_knn_merge_records + alive_knn_gather_rows + exact row sum + alive_knn_mean_blend (the NCCL path) - through the
same ShardedLibrary.match the multi-GPU bench runs.  Results must equal the oracle on the WHOLE library
(module/common.py:96-109) and, bit for bit, the unsharded call.  The multi-process NCCL/IPC plumbing itself is
checked by bench.py's `parity` block at every N and by tools/gpu_sharded_check.py under torchrun."""
import ctypes
import threading

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import alive_vc_b200 as A                                          # noqa: E402
from alive_vc_b200 import _cabi, matching as M                      # noqa: E402
from alive_vc_b200.sharded import (CudaShardBackend, PeerShards, ShardedLibrary, ThreadComm, record_bytes,   # noqa: E402
                                   shard_bounds)
from oracle import knn_oracle as O                                  # noqa: E402


def _cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _run_ranks(world, fn):
    """fn(comm) on `world` threads; returns the per-rank results, re-raises the first failure"""
    comms = ThreadComm.make(world)
    res, err = [None] * world, [None] * world

    def body(r):
        try:
            torch.cuda.set_device(0)
            res[r] = fn(comms[r])
        except BaseException as e:      # noqa: BLE001
            err[r] = e
            comms[r]._s.barrier.abort()

    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    for e in err:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in err:
        if e is not None:
            raise e
    return res


@pytest.mark.parametrize("peer", [True, False])
@pytest.mark.parametrize("world,n_total,T,k,alpha,B", [
    (2, 20_001, 300, 4, 0.0, 1),
    (3, 5_000, 64, 4, 0.25, 2),
    (8, 100_003, 257, 4, 0.0, 1),
    (3, 10, 5, 4, 0.0, 1),        # shards of 4/3/3 frames: some ranks hold fewer than k frames
    (3, 2, 4, 2, 0.5, 1),         # shards of 1/1/0 frames: one rank holds nothing
    (8, 9, 33, 8, 0.0, 1),        # k = 8 > every shard
])
def test_thread_ranks_equal_oracle_and_unsharded(world, n_total, T, k, alpha, B, peer):
    rng = np.random.default_rng(1000 * world + n_total + T)
    src = rng.standard_normal((B, 768, T), dtype=np.float32)
    ref = rng.standard_normal((1, 768, n_total), dtype=np.float32)
    s, r = _cuda(src), _cuda(ref)
    want_out, want_idx = A.match_features(s, r.expand(B, -1, -1), k, alpha, return_indices=True)
    ref_b = np.broadcast_to(ref, (B,) + ref.shape[1:])
    o_out, o_idx, _ = O.match_features_np(src, ref_b, k, alpha, True)
    scores = O.cosine_scores_np(src, ref_b)

    def rank_fn(comm):
        lib = ShardedLibrary.from_full(r, comm=comm, peer_memory=peer)
        assert (lib.peers is not None) == peer
        lo, hi = shard_bounds(n_total, comm.world, comm.rank)
        assert lib.n_local == hi - lo and lib.row_base == lo
        out, idx = lib.match(s, k, alpha, return_indices=True)
        res = out.clone(), idx.clone()
        torch.cuda.synchronize()
        comm.barrier()          # a rank's shard must outlive every peer's gather (ShardedLibrary.close protocol)
        return res

    for out, idx in _run_ranks(world, rank_fn):
        ok, _, _, bad = O.indices_match_mod_ties(idx.cpu().numpy(), o_idx, scores, 1e-6)
        assert ok, bad
        same = (idx.cpu().numpy() == o_idx).all(axis=2)
        o = out.cpu().numpy()
        assert np.array_equal(np.swapaxes(o, 1, 2)[same], np.swapaxes(o_out, 1, 2)[same])
        assert torch.equal(idx, want_idx) and torch.equal(out, want_out)          # bit-identical to one GPU
        assert tuple(out.shape) == (B, 768, T) and tuple(out.stride()) == (T * 768, 1, 768)


@pytest.mark.parametrize("peer", [True, False])
@pytest.mark.parametrize("world,n_total,T", [(2, 30_000, 301), (4, 50_000, 1000), (8, 70_001, 1003), (3, 4_000, 2)])
def test_scattered_queries_and_results(world, n_total, T, peer):
    """each rank hands over only ITS slice of the query frames and gets ITS slice of the result back"""
    rng = np.random.default_rng(world + n_total + T)
    src = rng.standard_normal((1, 768, T), dtype=np.float32)
    ref = rng.standard_normal((1, 768, n_total), dtype=np.float32)
    s, r = _cuda(src), _cuda(ref)
    want_out, want_idx = A.match_features(s, r, 4, 0.25, return_indices=True)

    def rank_fn(comm):
        lib = ShardedLibrary.from_full(r, comm=comm, peer_memory=peer)
        lo, hi = shard_bounds(T, comm.world, comm.rank)
        out, idx = lib.match(s[:, :, lo:hi], 4, 0.25, return_indices=True, scattered=True, t_total=T)
        res = lo, hi, out.clone(), idx.clone()
        torch.cuda.synchronize()
        comm.barrier()
        return res

    for lo, hi, out, idx in _run_ranks(world, rank_fn):
        assert tuple(out.shape) == (1, 768, hi - lo)
        assert torch.equal(out, want_out[:, :, lo:hi]) and torch.equal(idx, want_idx[:, lo:hi])


def test_clustered_library_sharded_collect_pass():
    """every shard runs its collect pass (tight clusters: nothing certifies) and the merged answer is still the
    single-library answer and the oracle's"""
    T, N, world = 200, 160_000, 4
    g = torch.Generator(device="cuda").manual_seed(123)
    cent = torch.randn(768, 40, device="cuda", generator=g)
    ref = (cent[:, torch.randint(0, 40, (N,), device="cuda", generator=g)] + 0.2 * torch.randn(768, N, device="cuda", generator=g))[None]
    src = (cent[:, torch.randint(0, 40, (T,), device="cuda", generator=g)] + 0.2 * torch.randn(768, T, device="cuda", generator=g))[None]
    want_out, want_idx = A.match_features(src, ref, 4, 0.0, return_indices=True, mode="exact")

    def rank_fn(comm):
        lib = ShardedLibrary.from_full(ref, comm=comm)
        out, idx = lib.match(src, 4, 0.0, return_indices=True)
        res = out.clone(), idx.clone()
        torch.cuda.synchronize()
        comm.barrier()
        return res

    for out, idx in _run_ranks(world, rank_fn):
        assert torch.equal(idx, want_idx) and torch.equal(out, want_out)
    sl = slice(0, 24)
    o_out, o_idx, _ = O.match_features_np(src[:, :, sl].cpu().numpy(), ref.cpu().numpy(), 4, 0.0, True)
    scores = O.cosine_scores_np(src[:, :, sl].cpu().numpy(), ref.cpu().numpy())
    ok, _, _, bad = O.indices_match_mod_ties(want_idx[:, sl].cpu().numpy(), o_idx, scores, 1e-6)
    assert ok, bad


def test_records_merge_and_fused_gather_entry_points():
    """the C entry points of the exchange step one by one: alive_knn_merge_records == alive_knn_merge,
    alive_knn_merge_gather == merge + alive_knn_gather_mean_peers == gather_rows (summed) + mean_blend"""
    c = _cabi.load()
    dev = torch.device("cuda")
    stream = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(5)
    R, T, k, N, D = 5, 333, 4, 40_000, 768
    ref = torch.randn(1, D, N, device=dev, generator=g)
    src = torch.randn(1, D, T, device=dev, generator=g)
    full = A.pack_library(ref)
    want_out, want_idx, want_sc = M.run_match(src, full, k, 0.3)
    shards, stride = [], record_bytes(T, k)
    gathered = torch.empty((R, stride), dtype=torch.uint8, device=dev)
    q = None
    for rnk in range(R):
        lo, hi = shard_bounds(N, R, rnk)
        sh = full.rows(lo, hi)
        shards.append(sh)
        be = CudaShardBackend(sh)
        q = be.pack_queries(src)
        be.local_topk(q, k, out_score=gathered[rnk, T * k * 8: T * k * 12].view(torch.float32).view(T, k),
                      out_idx=gathered[rnk, : T * k * 8].view(torch.int64).view(T, k))
    # (a) records merge == strided merge
    all_i = gathered[:, : T * k * 8].contiguous().view(torch.int64).view(R, T, k)
    all_s = gathered[:, T * k * 8: T * k * 12].contiguous().view(torch.float32).view(R, T, k)
    ts1, ti1 = torch.empty((T, k), device=dev), torch.empty((T, k), dtype=torch.int64, device=dev)
    ts2, ti2 = torch.empty((T, k), device=dev), torch.empty((T, k), dtype=torch.int64, device=dev)
    _cabi.check(c.alive_knn_merge(all_s.data_ptr(), all_i.data_ptr(), R, T, k, ts1.data_ptr(), ti1.data_ptr(), stream), "merge")
    _cabi.check(c.alive_knn_merge_records(gathered.data_ptr(), stride, R, T, k, ts2.data_ptr(), ti2.data_ptr(), stream), "merge_records")
    assert torch.equal(ti1, ti2) and torch.equal(ts1, ts2)
    assert torch.equal(ti1, want_idx[0]) and torch.equal(ts1, want_sc[0])
    # (b) fused merge + gather over the shards' blocks, a row range in the middle and the whole batch
    ptrs = torch.tensor([sh.raw.data_ptr() for sh in shards], dtype=torch.int64, device=dev)
    bounds = torch.tensor([sh.row_base for sh in shards] + [N], dtype=torch.int64, device=dev)
    for row0, rows in ((0, T), (100, 57), (T - 1, 1)):
        out = torch.full((rows, D), float("nan"), device=dev)
        ti = torch.full((rows, k), -7, dtype=torch.int64, device=dev)
        ts = torch.full((rows, k), float("nan"), device=dev)
        _cabi.check(c.alive_knn_merge_gather(gathered.data_ptr(), stride, R, T, k, row0, rows, ptrs.data_ptr(), bounds.data_ptr(),
                                             R, D, q.raw.data_ptr(), q.norms.data_ptr(), 0.3, out.data_ptr(), ts.data_ptr(),
                                             ti.data_ptr(), stream), "merge_gather")
        assert torch.equal(ti, want_idx[0, row0:row0 + rows]) and torch.equal(ts, want_sc[0, row0:row0 + rows])
        assert torch.equal(out, want_out[0, row0:row0 + rows])
    # (c) the two-step forms
    out_p = torch.empty((T, D), device=dev)
    _cabi.check(c.alive_knn_gather_mean_peers(ptrs.data_ptr(), bounds.data_ptr(), R, D, ti1.data_ptr(), T, k, q.raw.data_ptr(),
                                              None, 0.3, out_p.data_ptr(), stream), "gather_mean_peers")
    assert torch.equal(out_p, want_out[0])
    rows_sum = torch.zeros((T, k, D), device=dev)
    for sh in shards:
        rws = torch.empty((T, k, D), device=dev)
        _cabi.check(c.alive_knn_gather_rows(sh.raw.data_ptr(), sh.n, D, sh.row_base, ti1.data_ptr(), T, k, rws.data_ptr(), stream), "gather_rows")
        assert int((rws != 0).any(dim=2).sum()) == int(((ti1 >= sh.row_base) & (ti1 < sh.row_base + sh.n)).sum())
        rows_sum += rws                                   # adding zeros is exact
    out_b = torch.empty((T, D), device=dev)
    _cabi.check(c.alive_knn_mean_blend(rows_sum.data_ptr(), T, k, D, q.raw.data_ptr(), 0.3, out_b.data_ptr(), stream), "mean_blend")
    assert torch.equal(out_b, want_out[0])


@pytest.mark.parametrize("T,N,k,alpha,D", [(1000, 50_000, 4, 0.0, 768), (77, 3000, 8, 0.5, 768), (130, 5000, 16, 0.0, 256),
                                           (64, 2000, 4, 0.0, 1536), (50, 2000, 3, 0.0, 100), (33, 900, 40, 0.0, 768)])
def test_gather_mean_kernels_are_bit_exact(T, N, k, alpha, D):
    """K4 standalone (warp-per-query kernel, CTA kernel for other dims / long lists): with and without the query
    norms (which let it skip the query row), non-finite and negative-zero corner cases included, against the
    sequential float32 model of common.py:107-109"""
    g = torch.Generator(device="cuda").manual_seed(T + N + k)
    raw = torch.randn(N, D, device="cuda", generator=g)
    raw[:8] = 0.0
    raw[4:8] = -0.0
    q_raw = torch.randn(T, D, device="cuda", generator=g)
    q_raw[3, 5] = float("inf")
    q_raw[4, 7] = float("nan")
    idx = torch.randint(8, N, (T, k), device="cuda", generator=g)
    idx[0] = torch.arange(k, device="cuda") % 4                    # mean of +0 rows
    idx[1] = 4 + torch.arange(k, device="cuda") % 4                # mean of -0 rows: sign of zero decided by 0 * q
    q_norm = torch.linalg.vector_norm(q_raw.double(), dim=1).float()
    rows = raw[idx].cpu().numpy()                                   # [T, k, D]
    acc = rows[:, 0].copy()
    for j in range(1, k):
        acc = (acc + rows[:, j]).astype(np.float32)
    # true float32 division by k on the host (torch's CUDA `tensor / python_scalar` multiplies by a rounded reciprocal)
    with np.errstate(invalid="ignore"):
        want = ((acc / np.float32(k)).astype(np.float32) * np.float32(1.0 - alpha)).astype(np.float32) + \
            (q_raw.cpu().numpy() * np.float32(alpha)).astype(np.float32)
    want = torch.from_numpy(want.astype(np.float32)).cuda()
    c = _cabi.load()
    for qn in (q_norm, None):
        out = torch.empty((T, D), device="cuda")
        _cabi.check(c.alive_knn_gather_mean(raw.data_ptr(), N, D, idx.data_ptr(), T, k, q_raw.data_ptr(),
                                            qn.data_ptr() if qn is not None else None, alpha, out.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), "gather_mean")
        same = (out == want) | (out.isnan() & want.isnan())
        assert bool(same.all())
        assert torch.equal(torch.signbit(out[~out.isnan()]), torch.signbit(want[~want.isnan()]))
