"""Row-sharded voice library across the GPUs of one box (BASELINE cfg4).

The reference has no distributed code at all; this is the one exchange step the path
needs when the library does not fit / should not live on one GPU:

    rank r holds frames [lo_r, hi_r) packed (bf16 + raw fp32); every rank sees all queries
    1. local exact top-k            (K2 + K2b + K3 on the shard, GLOBAL frame indices)
    2. all-gather of (score, index) [T, k] per rank   - T*k*12 B per rank over NVLink
    3. merge -> global top-k        (score desc, frame asc)  on every rank
    4. each rank gathers the winning rows IT owns into a zero-initialised [T, k, D] block
    5. all-reduce(sum) of that block - adding zeros is exact, so every rank now holds the
       k winning raw frames of every query in descending-score order
    6. mean + blend exactly as the single-GPU K4 -> bit-identical result on every rank

One process per GPU (`torch.distributed`, backend nccl); the collectives are issued
through torch.distributed on the current stream's NCCL communicator.  The compute steps
go through a small backend object so the choreography can be exercised on CPU with
`gloo` in tests (tests inject an oracle-based backend; the product backend below is
CUDA-only and has no fallback).
"""
from __future__ import annotations

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


class CudaShardBackend:
    """The product backend: hand-written kernels through the C ABI."""

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

    def local_topk(self, q, k):
        """([T,k] float32, [T,k] int64 global indices); k <= frames on this shard."""
        B, D, T = q.source.shape
        _, idx, score = M.run_match(q.source, self.local, k, 0.0, self.mode, self.variant, want_out=False)
        ws = M.last_info._workspace
        q.raw = ws[: B * T * D * 4].view(torch.float32).view(B * T, D)     # offset 0 of the layout = q_raw
        q.workspace = ws
        return score.view(B * T, k), idx.view(B * T, k)

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
        rc = _cabi.load().alive_knn_merge(scores.data_ptr(), idx.data_ptr(), r, t, k, top_s.data_ptr(),
                                          top_i.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc, "alive_knn_merge")
        M._count(1)
        return top_s, top_i

    def gather_rows(self, top_idx):
        t, k = top_idx.shape
        if self.local.n == 0:      # a shard without frames contributes zeros
            return torch.zeros((t, k, self.local.d), dtype=torch.float32, device=top_idx.device)
        rows = torch.empty((t, k, self.local.d), dtype=torch.float32, device=top_idx.device)
        rc = _cabi.load().alive_knn_gather_rows(self.local.raw.data_ptr(), self.local.n, self.local.d,
                                                self.local.row_base, top_idx.data_ptr(), t, k, rows.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc, "alive_knn_gather_rows")
        M._count(1)
        return rows

    def mean_blend(self, rows, q, alpha, row0=0, out=None):
        """mean + blend of rows [t,k,d] for query rows [row0, row0+t) -> out [t,d]"""
        t, k, d = rows.shape
        if q.raw is None:       # this shard held no frames, so alive_knn_match never packed the queries
            q.raw = M.pack_queries(q.source).raw
        if out is None:
            out = torch.empty((t, d), dtype=torch.float32, device=rows.device)
        if t > 0:
            rc = _cabi.load().alive_knn_mean_blend(rows.data_ptr(), t, k, d, q.raw[row0:].data_ptr(), float(alpha),
                                                   out.data_ptr(), torch.cuda.current_stream().cuda_stream)
            _cabi.check(rc, "alive_knn_mean_blend")
            M._count(1)
        return out


class PeerShards:
    """Every rank's raw [n_r, D] shard mapped into this process through CUDA IPC (NVLink peer
    memory), so the final gather can read the k winning frames of each query wherever they live:
    one kernel, no collective after the top-k merge (alive_knn_gather_mean_peers)."""

    def __init__(self, local: M.PackedFrames, n_total: int, group=None):
        import ctypes
        c = _cabi.load()
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        dev = local.device
        handle = ctypes.create_string_buffer(64)
        off = ctypes.c_int64(0)
        export_error = ""
        if local.n > 0 and c.alive_knn_ipc_export(local.raw.data_ptr(), handle, ctypes.byref(off)) != 0:
            export_error = c.alive_knn_last_error().decode("utf-8", "replace")
        # exchange (handle, byte offset, rows, row_base, export ok) as a small byte tensor over the
        # group - every rank takes part in this collective even if its own export failed
        mine = torch.zeros(64 + 32, dtype=torch.uint8)
        mine[:64] = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8)
        mine[64:] = torch.tensor([off.value, local.n, local.row_base, 0 if export_error else 1],
                                 dtype=torch.int64).view(torch.uint8)
        everyone = torch.empty((world, 64 + 32), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(everyone, mine.to(dev), group=group)
        everyone = everyone.cpu()
        if not all(int(everyone[r, 64:].view(torch.int64)[3]) == 1 for r in range(world)):
            raise RuntimeError("CUDA IPC export failed on at least one rank" + (f": {export_error}" if export_error else ""))
        self._opened = []
        ptrs, bounds = [], []
        for r in range(world):
            r_off, r_n, r_base, _ = everyone[r, 64:].view(torch.int64).tolist()
            bounds.append(r_base)
            if r == rank or r_n == 0:
                ptrs.append(local.raw.data_ptr() if r == rank else 0)
                continue
            base = ctypes.c_void_p(0)
            _cabi.check(c.alive_knn_ipc_open(bytes(everyone[r, :64].tolist()), ctypes.byref(base)), "alive_knn_ipc_open")
            self._opened.append(base.value)
            ptrs.append(base.value + r_off)
        bounds.append(n_total)
        self.shards = world
        self.ptrs = torch.tensor(ptrs, dtype=torch.int64, device=dev)
        self.bounds = torch.tensor(bounds, dtype=torch.int64, device=dev)
        self._keep = local          # the exporting side must keep its allocation alive

    def gather_mean(self, top_idx, q_raw, alpha, d):
        t, k = top_idx.shape
        out = torch.empty((t, d), dtype=torch.float32, device=top_idx.device)
        rc = _cabi.load().alive_knn_gather_mean_peers(self.ptrs.data_ptr(), self.bounds.data_ptr(), self.shards, d,
                                                      top_idx.data_ptr(), t, k, q_raw.data_ptr(), float(alpha),
                                                      out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc, "alive_knn_gather_mean_peers")
        M._count(1)
        return out

    def close(self):
        c = _cabi.load()
        for b in self._opened:
            c.alive_knn_ipc_close(b)
        self._opened = []


class _Queries:
    def __init__(self, source):
        self.source = source
        self.raw = None
        self.workspace = None


class ShardedLibrary:
    """A voice library whose frames are split by rows over the ranks of `group`."""

    def __init__(self, backend, n_local: int, row_base: int, n_total: int, group=None, peer_memory: bool = False):
        self.backend = backend
        self.peers = None
        self.n_local = n_local
        self.row_base = row_base
        self.n_total = n_total
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if peer_memory and self.world > 1:
            # raw shards of every rank mapped over NVLink (CUDA IPC): the gather needs no collective.
            # Either every rank maps every peer or all of them use the NCCL exchange (never a mix).
            try:
                peers = PeerShards(backend.local, n_total, group)
                ok = 1
            except RuntimeError as e:          # e.g. an allocator whose blocks cannot be exported
                peers, ok = None, 0
                self.peer_error = str(e)
            flag = torch.tensor([ok], dtype=torch.int32, device=backend.local.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 1:
                self.peers = peers
            elif peers is not None:
                peers.close()

    def _reduce_scatter_ok(self) -> bool:
        """reduce_scatter_tensor exists on NCCL; gloo (CPU tests) keeps the all-reduce form."""
        try:
            return dist.get_backend(self.group) == "nccl"
        except Exception:
            return False

    @classmethod
    def from_local_frames(cls, frames_dn: torch.Tensor, row_base: int, n_total: int, group=None,
                          mode: str = "auto", variant: int = 0, peer_memory: bool = False):
        """`frames_dn` [D, n_local]: this rank's frames, global rows [row_base, row_base+n_local)."""
        local = M.pack_frames(frames_dn)
        local.row_base = row_base
        return cls(CudaShardBackend(local, mode, variant), local.n, row_base, n_total, group, peer_memory)

    @classmethod
    def from_full(cls, reference: torch.Tensor, group=None, mode: str = "auto", variant: int = 0,
                  peer_memory: bool = False):
        """Every rank passes the same [1, D, N] library; each keeps only its row range."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        ref = reference[0] if reference.dim() == 3 else reference
        lo, hi = shard_bounds(ref.shape[1], world, rank)
        return cls.from_local_frames(ref[:, lo:hi], lo, ref.shape[1], group, mode, variant, peer_memory)

    def match(self, source: torch.Tensor, k: int = 4, alpha: float = 0.0, return_indices: bool = False):
        """Same contract as match_features(source, whole_library): source [B, D, T] replicated
        on every rank -> [B, D, T] on every rank (transposed view of a contiguous [B,T,D])."""
        if source.dim() != 3:
            raise RuntimeError("ShardedLibrary.match expects source [B, D, T]")
        if not isinstance(k, int) or k < 1 or k > self.n_total:
            raise RuntimeError("selected index k out of range")
        B, D, T = source.shape
        be = self.backend
        if self.world == 1 and hasattr(be, "match_single") and self.n_local >= k:
            out_btd, idx, _ = be.match_single(source, k, alpha)
            out = out_btd.transpose(1, 2)
            if out.dtype != source.dtype:
                out = out.to(source.dtype)
            return (out, idx) if return_indices else out
        q = be.pack_queries(source)
        t = B * T
        dev = source.device
        # 1. local exact top-k (pad with -inf / -1 when the shard holds fewer than k frames)
        k_loc = min(k, self.n_local)
        loc_s = torch.full((t, k), float("-inf"), dtype=torch.float32, device=dev)
        loc_i = torch.full((t, k), -1, dtype=torch.int64, device=dev)
        if k_loc > 0:
            s, i = be.local_topk(q, k_loc)
            loc_s[:, :k_loc] = s
            loc_i[:, :k_loc] = i
        if self.world == 1:
            top_s, top_i = loc_s, loc_i
        else:
            # 2. all-gather candidates
            all_s = torch.empty((self.world * t, k), dtype=torch.float32, device=dev)
            all_i = torch.empty((self.world * t, k), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(all_s, loc_s, group=self.group)
            dist.all_gather_into_tensor(all_i, loc_i, group=self.group)
            # 3. merge
            top_s, top_i = be.merge(all_s.view(self.world, t, k), all_i.view(self.world, t, k), k)
        if self.peers is not None:
            # 4'. one kernel reads the k winning frames of every query from whichever GPU owns them
            if q.raw is None:
                q.raw = M.pack_queries(q.source).raw
            out_rows = self.peers.gather_mean(top_i, q.raw, alpha, D)
            out = out_rows.view(B, T, D).transpose(1, 2)
            if out.dtype != source.dtype:
                out = out.to(source.dtype)
            return (out, top_i.view(B, T, k)) if return_indices else out
        # 4. owned rows, zeros elsewhere;  5. exact sum over ranks;  6. mean + blend
        rows = be.gather_rows(top_i)
        if self.world > 1 and self._reduce_scatter_ok():
            # NCCL: reduce-scatter the zero-padded rows over the QUERY axis (half the traffic of an
            # all-reduce), finish T/R queries per rank, all-gather the [T, D] result
            R = self.world
            per = (t + R - 1) // R
            if per * R != t:
                rows = torch.cat([rows, rows.new_zeros((per * R - t, k, D))], 0)
            mine = torch.empty((per, k, D), dtype=torch.float32, device=dev)
            dist.reduce_scatter_tensor(mine, rows, op=dist.ReduceOp.SUM, group=self.group)
            row0 = self.rank * per
            valid = max(0, min(per, t - row0))
            out_slice = torch.zeros((per, D), dtype=torch.float32, device=dev) if valid < per else \
                torch.empty((per, D), dtype=torch.float32, device=dev)
            be.mean_blend(mine[:valid], q, alpha, row0=row0, out=out_slice)
            out_full = torch.empty((per * R, D), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(out_full, out_slice, group=self.group)
            out_rows = out_full[:t]
        else:
            if self.world > 1:
                dist.all_reduce(rows, op=dist.ReduceOp.SUM, group=self.group)
            out_rows = be.mean_blend(rows, q, alpha)
        out = out_rows.view(B, T, D).transpose(1, 2)
        if out.dtype != source.dtype:
            out = out.to(source.dtype)
        if return_indices:
            return out, top_i.view(B, T, k)
        return out
