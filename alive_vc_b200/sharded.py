"""Row-sharded voice library across the GPUs of one box (BASELINE cfg4).

The reference has no distributed code at all; this is the one exchange step the path
needs when the library does not fit / should not live on one GPU:

    rank r holds frames [lo_r, hi_r) packed (bf16 + raw fp32); every rank sees all queries
    1. local exact top-k (K2 + K2b + K3 on the shard, GLOBAL frame indices) written as ONE record per rank:
       idx [T, k] int64 | score [T, k] float32
    2. ONE all-gather of the records (T*k*12 B per rank over NVLink)
    3. peer path (default when every rank could map every other rank's raw shard through CUDA IPC):
       ONE kernel (alive_knn_merge_gather) merges the R lists of a query into the global top-k (score desc,
       frame asc) and reads the k winning raw frames straight from the GPUs that own them (NVLink peer loads),
       sums them in descending-score order, divides by k, blends - the arithmetic of the single-GPU K4, so the
       result is bit-identical.  A rank may produce all T rows, or only its own T/R slice (`scattered=True`).
       NCCL path (no peer mapping): merge, then each rank gathers the winning rows IT owns into a zero-initialised
       [T, k, D] block, reduce-scatter(sum) - adding zeros is exact -, mean + blend of T/R queries per rank,
       all-gather of the result.

`scattered=True` also shards the query traffic: a rank passes only ITS slice of the query frames (what it copied
from the host); the slices are all-gathered over NVLink instead of every rank pulling all T frames over PCIe.

One process per GPU (`torch.distributed`, backend nccl).  The collectives go through a small communicator object
and the compute steps through a backend object, so the choreography can be exercised without NCCL: on CPU with
`gloo` and an oracle backend (tests/test_sharded_gloo.py), and on ONE GPU with R in-process ranks (ThreadComm)
driving the real CUDA backend (tests/test_gpu_sharded.py).  The product backend is CUDA-only and has no fallback.
"""
from __future__ import annotations

import ctypes
import threading

import torch
import torch.distributed as dist

from . import _cabi
from . import matching as M


def shard_bounds(n_total: int, world: int, rank: int):
    """Frames [lo, hi) owned by `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def record_bytes(t: int, k: int) -> int:
    """Bytes of one rank's record (idx [t,k] int64 | score [t,k] float32), padded to 16."""
    return (t * k * 12 + 15) // 16 * 16


# ---------------------------------------------------------------------------------------------
# communicators
# ---------------------------------------------------------------------------------------------
class TorchDistComm:
    """torch.distributed process group (NCCL on the GPUs of one box; gloo in the CPU tests)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def all_gather(self, out: torch.Tensor, inp: torch.Tensor):
        """out (contiguous, world x the size of inp) = the ranks' `inp` in rank order"""
        dist.all_gather_into_tensor(out.view(-1), inp.reshape(-1), group=self.group)

    def all_reduce_sum(self, t: torch.Tensor):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def has_reduce_scatter(self) -> bool:
        try:
            return dist.get_backend(self.group) == "nccl"      # gloo keeps the all-reduce form
        except Exception:
            return False

    def reduce_scatter_sum(self, out: torch.Tensor, inp: torch.Tensor):
        dist.reduce_scatter_tensor(out, inp, op=dist.ReduceOp.SUM, group=self.group)

    def all_ok(self, ok: bool, device) -> bool:
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        return int(flag.item()) == 1

    def barrier(self):
        dist.barrier(group=self.group)

    def map_raw_shards(self, local: M.PackedFrames):
        """Every rank's raw [n_r, D] block as a pointer valid in THIS process: CUDA IPC over NVLink.
        Returns (pointers, row bases, handles to close).  Raises RuntimeError when a rank cannot export."""
        c = _cabi.load()
        dev = local.device
        handle = ctypes.create_string_buffer(64)
        off = ctypes.c_int64(0)
        export_error = ""
        if local.n > 0:
            with M._on(dev):
                if c.alive_knn_ipc_export(local.raw.data_ptr(), handle, ctypes.byref(off)) != 0:
                    export_error = c.alive_knn_last_error().decode("utf-8", "replace")
        # exchange (handle, byte offset, rows, row_base, export ok) as a small byte tensor over the
        # group - every rank takes part in this collective even if its own export failed
        mine = torch.zeros(64 + 32, dtype=torch.uint8)
        mine[:64] = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8)
        mine[64:] = torch.tensor([off.value, local.n, local.row_base, 0 if export_error else 1],
                                 dtype=torch.int64).view(torch.uint8)
        everyone = torch.empty((self.world, 64 + 32), dtype=torch.uint8, device=dev)
        self.all_gather(everyone, mine.to(dev))
        everyone = everyone.cpu()
        if not all(int(everyone[r, 64:].view(torch.int64)[3]) == 1 for r in range(self.world)):
            raise RuntimeError("CUDA IPC export failed on at least one rank" + (f": {export_error}" if export_error else ""))
        opened, ptrs, bases = [], [], []
        try:
            for r in range(self.world):
                r_off, r_n, r_base, _ = everyone[r, 64:].view(torch.int64).tolist()
                bases.append(r_base)
                if r == self.rank or r_n == 0:
                    ptrs.append(local.raw.data_ptr() if r == self.rank else 0)
                    continue
                base = ctypes.c_void_p(0)
                with M._on(dev):
                    _cabi.check(c.alive_knn_ipc_open(bytes(everyone[r, :64].tolist()), ctypes.byref(base)), "alive_knn_ipc_open")
                opened.append(base.value)
                ptrs.append(base.value + r_off)
        except RuntimeError:
            for b in opened:
                c.alive_knn_ipc_close(b)
            raise
        return ptrs, bases, opened


class ThreadComm:
    """R ranks as R threads of ONE process sharing one device and its default stream - the whole sharded
    choreography, real CUDA backend included, without NCCL or a second GPU (tests, and a smoke check of the
    C library's thread safety).  `ThreadComm.make(R)` returns the R per-rank communicators."""

    class _Shared:
        def __init__(self, world):
            self.world = world
            self.barrier = threading.Barrier(world, timeout=300)     # a rank that died breaks the barrier for the others
            self.slots = [None] * world

    def __init__(self, shared, rank):
        self._s, self.world, self.rank = shared, shared.world, rank

    @classmethod
    def make(cls, world: int):
        shared = cls._Shared(world)
        return [cls(shared, r) for r in range(world)]

    def _exchange(self, value):
        self._s.slots[self.rank] = value
        self._s.barrier.wait()
        got = list(self._s.slots)
        self._s.barrier.wait()            # nobody overwrites a slot before everyone has read it
        return got

    def all_gather(self, out, inp):
        parts = self._exchange(inp)
        out.view(self.world, -1).copy_(torch.stack([p.reshape(-1) for p in parts]))
        if out.is_cuda:
            torch.cuda.current_stream(out.device).synchronize()
        self._s.barrier.wait()            # inputs may be reused after this point

    def all_reduce_sum(self, t):
        parts = self._exchange(t)
        total = parts[0].clone()
        for p in parts[1:]:
            total += p                    # rank order: the same sum on every rank (adding zeros is exact anyway)
        if t.is_cuda:
            torch.cuda.current_stream(t.device).synchronize()
        self._s.barrier.wait()
        t.copy_(total)

    def has_reduce_scatter(self) -> bool:
        return True

    def reduce_scatter_sum(self, out, inp):
        parts = self._exchange(inp)
        per = inp.shape[0] // self.world
        total = parts[0][self.rank * per:(self.rank + 1) * per].clone()
        for p in parts[1:]:
            total += p[self.rank * per:(self.rank + 1) * per]
        out.copy_(total)
        if out.is_cuda:
            torch.cuda.current_stream(out.device).synchronize()
        self._s.barrier.wait()

    def all_ok(self, ok, device):
        return all(self._exchange(bool(ok)))

    def barrier(self):
        self._s.barrier.wait()

    def map_raw_shards(self, local):
        """same process, same device: the other ranks' blocks are plain pointers"""
        got = self._exchange((local.raw.data_ptr() if local.n > 0 else 0, local.row_base, local))
        return [g[0] for g in got], [g[1] for g in got], []


# ---------------------------------------------------------------------------------------------
# the product backend
# ---------------------------------------------------------------------------------------------
class _Queries:
    def __init__(self, source):
        self.source = source          # [B, D, T] float32
        self.raw = None               # [B*T, D] float32 row-major (K1's copy, inside the match workspace)
        self.norms = None             # [B*T]
        self.workspace = None


class CudaShardBackend:
    """The product backend: hand-written kernels through the C ABI."""

    writes_records = True          # local_topk(out_score=, out_idx=) writes the lists in place

    def __init__(self, local: M.PackedFrames, mode: str = "auto", variant: int = 0):
        self.local = local
        self.mode = mode
        self.variant = variant

    @property
    def device(self):
        return self.local.device

    def pack_queries(self, source):
        """Queries are packed inside alive_knn_match; keep the fp32 source and (after local_topk)
        the workspace that holds the row-major raw query frames needed by the blend."""
        return _Queries(source if source.dtype == torch.float32 else source.float())

    def _ensure_raw(self, q):
        if q.raw is None:       # this shard held no frames, so alive_knn_match never packed the queries
            p = M.pack_queries(q.source, fmt=self.local.format)
            q.raw, q.norms = p.raw, p.norms

    def local_topk(self, q, k, out_score=None, out_idx=None):
        """([T,k] float32, [T,k] int64 global indices); k <= frames on this shard.  With out_* the lists are
        written in place (the views of this rank's record)."""
        B, D, T = q.source.shape
        if out_idx is not None:
            out_idx, out_score = out_idx.view(B, T, k), out_score.view(B, T, k)
        sink = {}
        _, idx, score = M.run_match(q.source, self.local, k, 0.0, self.mode, self.variant, want_out=False,
                                    top_idx=out_idx, top_score=out_score, info_sink=sink)
        ws, off = sink["workspace"], sink["offsets"]
        rows = B * T
        q.raw = ws[off[0]: off[0] + rows * D * 4].view(torch.float32).view(rows, D)
        q.norms = ws[off[1]: off[1] + rows * 4].view(torch.float32)
        q.workspace = ws
        return score.view(rows, k), idx.view(rows, k)

    def match_single(self, source, k, alpha):
        """world == 1: the plain one-call pipeline (no exchange step needed)."""
        return M.run_match(source if source.dtype == torch.float32 else source.float(), self.local, k, alpha,
                           self.mode, self.variant)

    def merge(self, scores, idx, k):
        r, t, kk = scores.shape
        top_s = torch.empty((t, k), dtype=torch.float32, device=scores.device)
        top_i = torch.empty((t, k), dtype=torch.int64, device=scores.device)
        if kk != k:
            raise RuntimeError("merge expects k entries per rank")
        with M._on(scores.device):
            rc = _cabi.load().alive_knn_merge(scores.data_ptr(), idx.data_ptr(), r, t, k, top_s.data_ptr(),
                                              top_i.data_ptr(), M._stream_ptr(scores.device))
        _cabi.check(rc, "alive_knn_merge")
        M._count(1)
        return top_s, top_i

    def merge_records(self, gathered, stride, ranks, t, k):
        """global top-k straight from the all-gathered records ([ranks, stride] uint8)"""
        dev = gathered.device
        top_s = torch.empty((t, k), dtype=torch.float32, device=dev)
        top_i = torch.empty((t, k), dtype=torch.int64, device=dev)
        with M._on(dev):
            rc = _cabi.load().alive_knn_merge_records(gathered.data_ptr(), stride, ranks, t, k, top_s.data_ptr(),
                                                      top_i.data_ptr(), M._stream_ptr(dev))
        _cabi.check(rc, "alive_knn_merge_records")
        M._count(1)
        return top_s, top_i

    def merge_gather(self, gathered, stride, ranks, t, k, row0, rows, peers, q, alpha, want_idx):
        """alive_knn_merge_gather: merge + peer-memory gather + mean + blend for query rows [row0, row0+rows)"""
        dev = gathered.device
        d = self.local.d
        self._ensure_raw(q)
        out = torch.empty((rows, d), dtype=torch.float32, device=dev)
        top_i = torch.empty((rows, k), dtype=torch.int64, device=dev) if want_idx else None
        with M._on(dev):
            rc = _cabi.load().alive_knn_merge_gather(
                gathered.data_ptr(), stride, ranks, t, k, row0, rows, peers.ptrs.data_ptr(), peers.bounds.data_ptr(),
                peers.shards, d, q.raw.data_ptr(), q.norms.data_ptr() if q.norms is not None else None, float(alpha),
                out.data_ptr(), None, top_i.data_ptr() if want_idx else None, M._stream_ptr(dev))
        _cabi.check(rc, "alive_knn_merge_gather")
        M._count(1)
        return out, top_i

    def gather_rows(self, top_idx):
        t, k = top_idx.shape
        if self.local.n == 0:      # a shard without frames contributes zeros
            return torch.zeros((t, k, self.local.d), dtype=torch.float32, device=top_idx.device)
        rows = torch.empty((t, k, self.local.d), dtype=torch.float32, device=top_idx.device)
        with M._on(top_idx.device):
            rc = _cabi.load().alive_knn_gather_rows(self.local.raw.data_ptr(), self.local.n, self.local.d,
                                                    self.local.row_base, top_idx.data_ptr(), t, k, rows.data_ptr(),
                                                    M._stream_ptr(top_idx.device))
        _cabi.check(rc, "alive_knn_gather_rows")
        M._count(1)
        return rows

    def mean_blend(self, rows, q, alpha, row0=0, out=None):
        """mean + blend of rows [t,k,d] for query rows [row0, row0+t) -> out [t,d]"""
        t, k, d = rows.shape
        self._ensure_raw(q)
        if out is None:
            out = torch.empty((t, d), dtype=torch.float32, device=rows.device)
        if t > 0:
            with M._on(rows.device):
                rc = _cabi.load().alive_knn_mean_blend(rows.data_ptr(), t, k, d, q.raw[row0:].data_ptr(), float(alpha),
                                                       out.data_ptr(), M._stream_ptr(rows.device))
            _cabi.check(rc, "alive_knn_mean_blend")
            M._count(1)
        return out


class PeerShards:
    """Every rank's raw [n_r, D] shard addressable from this process (CUDA IPC over NVLink between processes,
    plain pointers between the in-process ranks of ThreadComm), so the final step can read the k winning
    frames of each query wherever they live: no collective after the all-gather of the top-k records."""

    def __init__(self, local: M.PackedFrames, n_total: int, comm):
        ptrs, bases, self._opened = comm.map_raw_shards(local)
        dev = local.device
        self.shards = comm.world
        self.ptrs = torch.tensor(ptrs, dtype=torch.int64, device=dev)
        self.bounds = torch.tensor(list(bases) + [n_total], dtype=torch.int64, device=dev)
        self._keep = local          # the exporting side must keep its allocation alive

    def gather_mean(self, top_idx, q_raw, alpha, d, q_norm=None):
        t, k = top_idx.shape
        out = torch.empty((t, d), dtype=torch.float32, device=top_idx.device)
        with M._on(top_idx.device):
            rc = _cabi.load().alive_knn_gather_mean_peers(self.ptrs.data_ptr(), self.bounds.data_ptr(), self.shards, d,
                                                          top_idx.data_ptr(), t, k, q_raw.data_ptr(),
                                                          q_norm.data_ptr() if q_norm is not None else None, float(alpha),
                                                          out.data_ptr(), M._stream_ptr(top_idx.device))
        _cabi.check(rc, "alive_knn_gather_mean_peers")
        M._count(1)
        return out

    def close(self):
        c = _cabi.load()
        for b in self._opened:
            c.alive_knn_ipc_close(b)
        self._opened = []


class ShardedLibrary:
    """A voice library whose frames are split by rows over the ranks of `group`.

    peer_memory: True = map every rank's raw shard into every process (CUDA IPC) and finish with ONE fused
    merge + gather kernel; falls back to the NCCL row exchange when any rank cannot (all ranks together,
    never a mix); False = always the NCCL row exchange."""

    def __init__(self, backend, n_local: int, row_base: int, n_total: int, group=None, peer_memory: bool = True,
                 comm=None):
        self.backend = backend
        self.peers = None
        self.peer_error = None
        self.n_local = n_local
        self.row_base = row_base
        self.n_total = n_total
        self.comm = comm if comm is not None else TorchDistComm(group)
        self.group = group
        self.world = self.comm.world
        self.rank = self.comm.rank
        self._stage = {}
        if peer_memory and self.world > 1 and hasattr(backend, "merge_gather"):
            # Either every rank maps every peer or all of them use the NCCL exchange (never a mix).
            try:
                peers = PeerShards(backend.local, n_total, self.comm)
                ok = True
            except RuntimeError as e:          # e.g. an allocator whose blocks cannot be exported
                peers, ok = None, False
                self.peer_error = str(e)
            if self.comm.all_ok(ok, backend.local.device):
                self.peers = peers
            elif peers is not None:
                peers.close()

    def close(self):
        """Unmap the peers' shards (call on every rank, after a barrier, before the libraries are freed)."""
        if self.peers is not None:
            self.peers.close()
            self.peers = None

    @classmethod
    def from_local_frames(cls, frames_dn: torch.Tensor, row_base: int, n_total: int, group=None,
                          mode: str = "auto", variant: int = 0, peer_memory: bool = True, comm=None):
        """`frames_dn` [D, n_local]: this rank's frames, global rows [row_base, row_base+n_local)."""
        local = M.pack_frames(frames_dn)
        local.row_base = row_base
        return cls(CudaShardBackend(local, mode, variant), local.n, row_base, n_total, group, peer_memory, comm)

    @classmethod
    def from_full(cls, reference: torch.Tensor, group=None, mode: str = "auto", variant: int = 0,
                  peer_memory: bool = True, comm=None):
        """Every rank passes the same [1, D, N] library; each keeps only its row range."""
        c = comm if comm is not None else TorchDistComm(group)
        ref = reference[0] if reference.dim() == 3 else reference
        lo, hi = shard_bounds(ref.shape[1], c.world, c.rank)
        return cls.from_local_frames(ref[:, lo:hi], lo, ref.shape[1], group, mode, variant, peer_memory, c)

    # -- the exchange step ------------------------------------------------------------------------
    def _gather_queries(self, source, t_total):
        """scattered queries: this rank's [1, D, T_r] slice -> every rank's slices [R, D, per] (padding = 1.0:
        an ordinary frame that certifies, never a zero frame)."""
        R = self.world
        per = (t_total + R - 1) // R
        lo, hi = shard_bounds(t_total, R, self.rank)
        if source.dim() != 3 or source.shape[0] != 1 or source.shape[2] != hi - lo:
            raise RuntimeError(f"scattered match: rank {self.rank} must pass its [1, D, {hi - lo}] slice of the {t_total} "
                               f"query frames (got {tuple(source.shape)})")
        D = source.shape[1]
        key = (D, per, source.device)
        stage = self._stage.get(key)
        if stage is None:
            stage = torch.ones((D, per), dtype=torch.float32, device=source.device)
            self._stage[key] = stage
        stage[:, :hi - lo].copy_(source[0])
        full = torch.empty((R, D, per), dtype=torch.float32, device=source.device)
        self.comm.all_gather(full, stage)
        return full, per, hi - lo

    def match(self, source: torch.Tensor, k: int = 4, alpha: float = 0.0, return_indices: bool = False,
              scattered: bool = False, t_total: int | None = None):
        """Same contract as match_features(source, whole_library): source [B, D, T] replicated
        on every rank -> [B, D, T] on every rank (transposed view of a contiguous [B,T,D]).

        scattered=True (needs `t_total`): `source` is THIS rank's slice [1, D, T_r] of the T = t_total query frames
        (contiguous runs in rank order, sizes as shard_bounds(t_total, world, rank)); returns this rank's slice
        [1, D, T_r] of the result (+ its [1, T_r, k] indices)."""
        if source.dim() != 3:
            raise RuntimeError("ShardedLibrary.match expects source [B, D, T]")
        if not isinstance(k, int) or k < 1 or k > self.n_total:
            raise RuntimeError("selected index k out of range")
        be = self.backend
        src_dtype = source.dtype
        if scattered and self.world > 1:
            if t_total is None:
                raise RuntimeError("scattered match needs t_total (the number of query frames over all ranks)")
            if source.dtype != torch.float32:
                source = source.float()
            source, per, mine = self._gather_queries(source, t_total)
            row0, rows_out = self.rank * per, mine
        else:
            scattered = False
        B, D, T = source.shape
        if self.world == 1 and hasattr(be, "match_single") and self.n_local >= k:
            out_btd, idx, _ = be.match_single(source, k, alpha)
            out = out_btd.transpose(1, 2)
            if out.dtype != src_dtype:
                out = out.to(src_dtype)
            return (out, idx) if return_indices else out
        q = be.pack_queries(source)
        t = B * T
        dev = source.device
        if not scattered:
            row0, rows_out = 0, t
        # 1. local exact top-k, written as this rank's record (padded with -inf / -1 when the shard holds fewer
        #    than k frames)
        stride = record_bytes(t, k)
        rec = torch.empty((stride,), dtype=torch.uint8, device=dev)
        rec_i = rec[: t * k * 8].view(torch.int64).view(t, k)
        rec_s = rec[t * k * 8: t * k * 12].view(torch.float32).view(t, k)
        k_loc = min(k, self.n_local)
        if k_loc == k:
            if getattr(be, "writes_records", False):
                be.local_topk(q, k, out_score=rec_s, out_idx=rec_i)          # written in place
            else:                                   # a backend that returns fresh tensors (the CPU test backend)
                s, i = be.local_topk(q, k)
                rec_s.copy_(s)
                rec_i.copy_(i)
        else:
            rec_s.fill_(float("-inf"))
            rec_i.fill_(-1)
            if k_loc > 0:
                s, i = be.local_topk(q, k_loc)
                rec_s[:, :k_loc] = s
                rec_i[:, :k_loc] = i
        # 2. ONE all-gather of the records
        if self.world == 1:
            gathered = rec.view(1, stride)
        else:
            gathered = torch.empty((self.world, stride), dtype=torch.uint8, device=dev)
            self.comm.all_gather(gathered, rec)
        if self.peers is not None:
            # 3'. one kernel: merge + read the k winning frames of every query from whichever GPU owns them
            out_rows, top_i = be.merge_gather(gathered, stride, self.world, t, k, row0, rows_out, self.peers, q, alpha,
                                              return_indices)
        else:
            # 3. merge;  4. owned rows, zeros elsewhere;  5. exact sum over ranks;  6. mean + blend
            if hasattr(be, "merge_records"):
                _, top_i = be.merge_records(gathered, stride, self.world, t, k)
            else:
                all_i = gathered[:, : t * k * 8].contiguous().view(torch.int64).view(self.world, t, k)
                all_s = gathered[:, t * k * 8: t * k * 12].contiguous().view(torch.float32).view(self.world, t, k)
                _, top_i = be.merge(all_s, all_i, k)
            rows = be.gather_rows(top_i)
            if self.world > 1 and self.comm.has_reduce_scatter():
                # reduce-scatter the zero-padded rows over the QUERY axis (half the traffic of an
                # all-reduce), finish T/R queries per rank, all-gather the [T, D] result
                R = self.world
                per_q = (t + R - 1) // R
                if per_q * R != t:
                    rows = torch.cat([rows, rows.new_zeros((per_q * R - t, k, D))], 0)
                mine_rows = torch.empty((per_q, k, D), dtype=torch.float32, device=dev)
                self.comm.reduce_scatter_sum(mine_rows, rows)
                q0 = self.rank * per_q
                valid = max(0, min(per_q, t - q0))
                out_slice = torch.zeros((per_q, D), dtype=torch.float32, device=dev) if valid < per_q else \
                    torch.empty((per_q, D), dtype=torch.float32, device=dev)
                be.mean_blend(mine_rows[:valid], q, alpha, row0=q0, out=out_slice)
                out_full = torch.empty((per_q * R, D), dtype=torch.float32, device=dev)
                self.comm.all_gather(out_full, out_slice)
                out_rows = out_full[:t]
            else:
                if self.world > 1:
                    self.comm.all_reduce_sum(rows)
                out_rows = be.mean_blend(rows, q, alpha)
            if scattered:
                out_rows, top_i = out_rows[row0:row0 + rows_out], top_i[row0:row0 + rows_out]
        if scattered:
            out = out_rows.view(1, rows_out, D).transpose(1, 2)
            top_i = top_i.view(1, rows_out, k) if top_i is not None else None
        else:
            out = out_rows.view(B, T, D).transpose(1, 2)
            top_i = top_i.view(B, T, k) if top_i is not None else None
        if out.dtype != src_dtype:
            out = out.to(src_dtype)
        if return_indices:
            return out, top_i
        return out
