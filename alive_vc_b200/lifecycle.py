"""Callers either side of the match: building a voice library on the GPU, matching
all overlapped windows of an utterance in one call, the host-buffer realtime loop, and the row-major
frame format for producers that emit [T, 768] (pack_rows / match_rows).

Reference call sites mirrored here (nothing below re-implements the encoders/decoder around them):
    generate_voice_library.py:30-42   512 random frames written into random slots of `tokens`, saved
    inference.py:67-84                library = cat([CE(target), VL.tokens], dim=2)
    inference.py:96-134               per 3x-overlapped window: feat = match_features(CE(spec), tgt)
    realtime_inference.py:130-191     per audio block: host chunk -> device -> match -> host
All device work goes through the C ABI (include/alive_knn.h); torch is used for buffers and copies.
"""
from __future__ import annotations

import ctypes

import torch

from . import matching as M


class LibraryBuilder:
    """Accumulates library frames on the GPU and emits them in the packed layout.

    `put(slots, frames)` has the semantics of the reference's generation loop
    (`VL.tokens.data[:, :, n] = t`, generate_voice_library.py:36-38): writes are applied in order, the
    last write to a slot wins, untouched slots keep their previous content.  `append(frames)` grows
    the library by whole utterances (inference.py:76 / realtime_inference.py:88 concatenate encoder
    output the same way), so a library of arbitrary N can be built from a corpus without ever leaving
    the device.  `packed()` runs K1 (alive_knn_pack) once over the accumulated frames.

    Slots that were never written have no content: the reference's loop starts from `VoiceLibrary()`,
    whose tokens are random-normal (voice_library.py:9), so its untouched slots are noise frames.  Ask for
    the same start with `init="randn"` (all `capacity` slots filled like `VoiceLibrary(num_tokens=capacity)`)
    or pass the initial `tokens=`; otherwise `tokens()` / `packed()` / `save()` raise while a slot below
    the highest written one is still empty (an all-zero frame would rank first in every match, like the
    NaN similarity it produces in the reference).
    """

    def __init__(self, d: int = 768, capacity: int = 512, device="cuda", tokens: torch.Tensor | None = None,
                 init: str | None = None, generator: torch.Generator | None = None):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("alive_vc_b200: LibraryBuilder needs a CUDA device (no CPU path)")
        if d < 1 or capacity < 0:
            raise ValueError("LibraryBuilder: d >= 1 and capacity >= 0 expected")
        if init not in (None, "randn"):
            raise ValueError("LibraryBuilder: init must be None or 'randn'")
        self.d = d
        self._rows = torch.zeros((max(capacity, 1), d), dtype=torch.float32, device=dev)   # [cap, D] row-major
        self._written = torch.zeros((max(capacity, 1),), dtype=torch.bool, device=dev)
        self._n = 0
        self._packed = None
        if tokens is not None:
            self.append(tokens)
        elif init == "randn" and capacity > 0:
            # VoiceLibrary(num_tokens=capacity).tokens (voice_library.py:9), generated in its [1, D, N] layout
            self.append(torch.randn(1, d, capacity, device=dev, generator=generator))

    # -- geometry -----------------------------------------------------------------------------
    def __len__(self) -> int:
        return self._n

    @property
    def device(self):
        return self._rows.device

    def _reserve(self, n: int):
        if n <= self._rows.shape[0]:
            return
        cap = max(n, 2 * self._rows.shape[0])
        grown = torch.zeros((cap, self.d), dtype=torch.float32, device=self.device)
        grown[:self._n] = self._rows[:self._n]
        self._rows = grown
        mask = torch.zeros((cap,), dtype=torch.bool, device=self.device)
        mask[:self._n] = self._written[:self._n]
        self._written = mask

    def unwritten_slots(self) -> torch.Tensor:
        """Indices of the slots below len(self) that no put/append has filled."""
        return (~self._written[:self._n]).nonzero().reshape(-1)

    def _require_complete(self):
        missing = self.unwritten_slots()
        if missing.numel():
            raise RuntimeError(
                f"LibraryBuilder: {missing.numel()} of {self._n} slots were never written (first: "
                f"{missing[:8].tolist()}); start from tokens=... or init='randn' (what VoiceLibrary() holds in "
                "generate_voice_library.py:30) or fill them")

    @staticmethod
    def _as_dn(frames: torch.Tensor, d: int) -> torch.Tensor:
        """[1, D, n] / [D, n] / [D] -> [D, n] float32 view."""
        if frames.dim() == 3:
            if frames.shape[0] != 1:
                raise RuntimeError("LibraryBuilder: frames must be [1, D, n], [D, n] or [D]")
            frames = frames[0]
        if frames.dim() == 1:
            frames = frames.unsqueeze(1)
        if frames.dim() != 2 or frames.shape[0] != d:
            raise RuntimeError(f"LibraryBuilder: expected {d} channels, got shape {tuple(frames.shape)}")
        return frames.detach().float()

    # -- writes -------------------------------------------------------------------------------
    def append(self, frames: torch.Tensor) -> "LibraryBuilder":
        """Add the n frames of a [1, D, n] / [D, n] tensor after the current last frame."""
        f = self._as_dn(frames, self.d).to(self.device)
        n = f.shape[1]
        self._reserve(self._n + n)
        self._rows[self._n:self._n + n].copy_(f.t())
        self._written[self._n:self._n + n] = True
        self._n += n
        self._packed = None
        return self

    def put(self, slots, frames: torch.Tensor) -> "LibraryBuilder":
        """`tokens[:, :, slots[i]] = frames[:, i]` for i = 0..m-1 IN ORDER (last write wins)."""
        f = self._as_dn(frames, self.d).to(self.device)
        slots = torch.as_tensor(slots, dtype=torch.int64, device=self.device).reshape(-1)
        m = slots.numel()
        if f.shape[1] != m:
            raise RuntimeError(f"LibraryBuilder.put: {m} slots but {f.shape[1]} frames")
        if m == 0:
            return self
        lo, hi = int(slots.min()), int(slots.max())
        if lo < 0:
            raise IndexError("LibraryBuilder.put: negative slot")
        self._reserve(hi + 1)
        self._n = max(self._n, hi + 1)
        # position of the last write to every slot; only those writes land
        last = torch.full((hi + 1,), -1, dtype=torch.int64, device=self.device)
        last.scatter_reduce_(0, slots, torch.arange(m, device=self.device), reduce="amax", include_self=True)
        written = (last >= 0).nonzero().reshape(-1)
        self._rows[written] = f.t()[last[written]]
        self._written[written] = True
        self._packed = None
        return self

    # -- reads --------------------------------------------------------------------------------
    def tokens(self) -> torch.Tensor:
        """The library as the reference stores it: [1, D, N] float32 (module/voice_library.py:9)."""
        self._require_complete()
        return self._rows[:self._n].t().unsqueeze(0).contiguous()

    def packed(self) -> M.PackedFrames:
        if self._n == 0:
            raise RuntimeError("selected index k out of range")      # an empty library cannot be matched
        self._require_complete()
        if self._packed is None:
            self._packed = M.pack_frames(self._rows[:self._n].t())    # [D, N] view of the row-major block
        return self._packed

    def save(self, path: str):
        """`path` = the reference's own checkpoint ({"tokens": [1, D, N]}, strict-loadable by the unmodified
        reference), `path + ".alive_knn"` = the packed layout (matching.save_packed_library)."""
        M.save_packed_library(self.packed(), path, include_legacy_tokens=True)


def match_windows(windows, lib: M.PackedFrames, k: int = 4, alpha: float = 0.0, mode: str = "auto"):
    """All windows of an utterance against one packed library in ONE pipeline launch.

    The reference converts an utterance as 3x-overlapped windows and runs the match once per window
    (inference.py:96-134, `feat = match_features(feat, tgt, ...)` at :129).  Every window is matched
    against the same library and frames are matched independently, so the windows can be laid end to
    end: `windows` is a [W, D, Tw] tensor or a sequence of [1, D, T_i] / [D, T_i] tensors of ragged
    length.  Returns a list of [1, D, T_i] tensors (transposed views of one contiguous [sum T, D]
    block), bit-identical to W separate match_features calls.
    """
    if isinstance(windows, torch.Tensor):
        if windows.dim() != 3:
            raise RuntimeError("match_windows expects [W, D, Tw] or a sequence of [1, D, T_i]")
        pieces = [windows[w] for w in range(windows.shape[0])]
    else:
        pieces = [w[0] if w.dim() == 3 else w for w in windows]
    if not pieces:
        return []
    for p in pieces:
        if p.dim() != 2 or p.shape[0] != lib.d:
            raise RuntimeError(f"match_windows: every window must have {lib.d} channels")
    lens = [int(p.shape[1]) for p in pieces]
    total = sum(lens)
    dev = lib.device
    if total == 0:
        return [torch.empty((1, lib.d, 0), dtype=torch.float32, device=dev) for _ in pieces]
    # frames of all windows end to end, already in the kernels' [T, D] row-major order
    rows = torch.empty((total, lib.d), dtype=torch.float32, device=dev)
    at = 0
    for p, n in zip(pieces, lens):
        rows[at:at + n].copy_(p.t())
        at += n
    out, _, _ = M.match_packed(rows.t().unsqueeze(0), lib, k, float(alpha), mode)     # [1, total, D]
    res, at = [], 0
    for n in lens:
        res.append(out[:, at:at + n].transpose(1, 2))
        at += n
    return res


def pack_rows(frames_nd: torch.Tensor) -> M.PackedFrames:
    """Pack a library a producer already holds ROW-major: [N, D] (or [1, N, D]) float32 CUDA frames,
    e.g. content-encoder output kept as `[T, 768]` instead of the reference's `[1, 768, T]`
    (module/content_encoder.py:21-25 emits channel-major only because Conv1d does).  K1 reads each
    3 KB frame as one coalesced row - no transpose pass, no channel-major staging copy."""
    if frames_nd.dim() == 3:
        if frames_nd.shape[0] != 1:
            raise RuntimeError("pack_rows expects [N, D] or [1, N, D]")
        frames_nd = frames_nd[0]
    if frames_nd.dim() != 2:
        raise RuntimeError("pack_rows expects [N, D] or [1, N, D]")
    return M.pack_frames(frames_nd.t())        # a [D, N] VIEW with stride_d == 1: K1's row-major branch


def match_rows(frames: torch.Tensor, lib, k: int = 4, alpha: float = 0.0, *, return_indices: bool = False,
               mode: str = "auto"):
    """The match for a producer/consumer pair that works on ROW-major frames.

    `frames` is [T, D] or [B, T, D] (any strides; the natural output of a channels-last encoder),
    `lib` a PackedFrames (pack_rows / pack_library / LibraryBuilder.packed()) or a row-major [N, D]
    tensor.  Returns the matched features in the SAME row-major shape, contiguous - the block the
    kernels write, handed over without the transposed view `match_features` has to return to look like
    common.py:108.  Values and indices are bit-identical to
    `match_features(frames.transpose(-1, -2), tokens)`; the per-call query transpose of
    common.py:100 never happens in either direction (the C ABI takes the strides as they are).
    """
    squeeze = frames.dim() == 2
    f = frames.unsqueeze(0) if squeeze else frames
    if f.dim() != 3:
        raise RuntimeError("match_rows expects [T, D] or [B, T, D]")
    if not isinstance(lib, M.PackedFrames):
        lib = M.cached_pack(lib, (lib[0] if lib.dim() == 3 else lib).t(), tag=1)
    if f.shape[2] != lib.d:
        raise RuntimeError(f"feature dims differ: queries {f.shape[2]}, library {lib.d}")
    M._require_cuda(f, "frames")
    src = f if f.dtype == torch.float32 else f.float()
    B, T, _ = src.shape
    if T == 0:
        out = torch.empty((B, 0, lib.d), dtype=torch.float32, device=src.device)
        idx = torch.empty((B, 0, k), dtype=torch.int64, device=src.device)
    else:
        with torch.no_grad():
            out, idx, _ = M.match_packed(src.transpose(1, 2), lib, k, float(alpha), mode)   # a view: no copy
    if out.dtype != frames.dtype:
        out = out.to(frames.dtype)
    if squeeze:
        out, idx = out[0], idx[0]
    if torch.is_grad_enabled() and frames.requires_grad:
        out = M._BlendGrad.apply(frames, out, float(alpha))      # d(out)/d(frames) = alpha, as in match_features
    return (out, idx) if return_indices else out


class HostStreamingMatcher:
    """The realtime loop's chunk step with HOST buffers (realtime_inference.py:130-191 moves every block
    host -> device -> host): pinned input and output buffers and ONE CUDA graph per chunk.

        hm = HostStreamingMatcher(pack_library(tgt), T=24)
        out = hm(chunk_cpu)            # [B, D, T] float32 CPU tensor -> [B, D, T] view of the pinned result

    Match only (no `pre` / `post`): the graph is one H2D copy node and the match pipeline; the last kernel of the
    pipeline writes the matched frames (contiguous 3 KB rows) straight into the pinned result over PCIe (unified
    addressing: zero_copy="out", the default) - no D2H copy node; a call is a memcpy into the pinned buffer, one
    graph launch and one wait.  zero_copy="both" also lets K1 read the chunk in place (no H2D node either),
    zero_copy=False keeps both copy nodes.  `hm.src_host` may be filled in place (`hm.run()`).
    With zero_copy="out"/"both" the wait is an EARLY one (`early=True`, alive_knn_arm_notify): a one-thread kernel
    right behind the finish kernel raises a flag in pinned host memory once the certified rows have landed there, and
    `result()` returns on that flag.  The six launches of the fallback chain (idle whenever every query certifies) are
    not part of the graph at all (ALIVE_KNN_MODE_DEFER_FALLBACK); if the flag says some query was uncertified,
    `result()` enqueues them (alive_knn_match_fallback) and waits for the stream.

    With caller modules - `pre` (e.g. the content encoder, realtime_inference.py:150: spectrogram chunk -> [B, D, T]
    features) and/or `post` (e.g. the decoder, :166) - the SAME graph holds  H2D copy -> pre -> match -> post -> D2H copy
    (encoder -> match -> decoder in one capture).  `pre` / `post` must be capture-safe (static shapes,
    no host synchronisation) like any module run under torch.cuda.graph; `in_shape` is the shape of the host chunk
    `pre` consumes (default [batch, D, T]), the output shape is whatever `post` returns.
    """

    def __init__(self, lib: M.PackedFrames, T: int, k: int = 4, alpha: float = 0.0, batch: int = 1,
                 mode: str = "auto", variant: int = 0, r_max: int = M.DEFAULT_R_MAX, pre=None, post=None,
                 in_shape=None, in_dtype=torch.float32, zero_copy="out", early=True):
        dev = lib.device
        self.lib, self.pre, self.post = lib, pre, post
        if zero_copy not in (False, "out", "both"):
            raise ValueError("zero_copy must be False, 'out' or 'both'")
        # zero-copy needs the match to be the first / last thing in the graph
        self.zc_in = zero_copy == "both" and pre is None
        self.zc_out = zero_copy in ("out", "both") and post is None
        self.early = bool(early) and self.zc_out
        self._c = M._cabi.load()
        self.flag_host = torch.zeros(1, dtype=torch.int64).pin_memory()
        self._flag_ptr = self.flag_host.data_ptr()
        self._flag_last = ctypes.c_uint64(0)
        self.inner = M.StreamingMatcher(lib, T, k, alpha, batch, mode, variant, r_max, use_graph=False)
        in_shape = tuple(in_shape) if in_shape is not None else (batch, lib.d, T)
        self.src_host = torch.zeros(in_shape, dtype=in_dtype).pin_memory()
        self.stream = torch.cuda.Stream(device=dev)
        self.done = torch.cuda.Event()
        with torch.cuda.device(dev):
            self.in_dev = None if self.zc_in else torch.zeros(in_shape, dtype=in_dtype, device=dev)
            self.flag_ctr = torch.zeros(1, dtype=torch.int64, device=dev)
            if post is None:
                self.out_host = torch.zeros((batch, T, lib.d), dtype=torch.float32).pin_memory()
            else:
                with torch.cuda.stream(self.stream), torch.no_grad():
                    probe = self._device_step(None)            # warm-up outside capture; fixes the output shape
                self.stream.synchronize()
                self.out_host = torch.zeros(tuple(probe.shape), dtype=probe.dtype).pin_memory()
            with torch.cuda.stream(self.stream), torch.no_grad():
                self._enqueue()                            # warm-up outside capture
            self.stream.synchronize()
            self.launches_per_call = M.last_info.launches  # kernels of OURS per graph replay
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream), torch.no_grad():
                self._enqueue()
            # per-chunk host work = two C calls (alive_knn_graph_launch: cudaGraphLaunch + cudaEventRecord;
            # alive_knn_event_wait: poll) on raw handles
            self.done.record(self.stream)
            self.stream.synchronize()
            self._flag_last.value = int(self.flag_host[0].item()) & 0xFFFFFFFFFFFFFFFF      # after the warm-up run
            self._exec = self.graph.raw_cuda_graph_exec()
            self._stream_h, self._event_h, self._dev_index = self.stream.cuda_stream, self.done.cuda_event, dev.index

    def _device_step(self, out):
        """[pre ->] match [-> post]; `out`: where the match writes its [B, T, D] block (None = the device buffer).
        Returns what has to reach the host (None when the match already wrote it there)."""
        i = self.inner
        if self.zc_in:
            src = self.src_host                                 # K1 reads the pinned chunk in place
        else:
            feat = self.pre(self.in_dev) if self.pre is not None else self.in_dev
            if feat.dtype != torch.float32:
                feat = feat.float()                             # (fp16 encoder output under autocast)
            src = feat
        if self.early:
            M._cabi.check(self._c.alive_knn_arm_notify(self._flag_ptr, self.flag_ctr.data_ptr()), "alive_knn_arm_notify")
        try:
            M.run_match(src, self.lib, i.k, i.alpha, i.mode, i.variant, i.r_max, workspace=i.workspace,
                        out=out if out is not None else i.out, top_idx=i.top_idx, top_score=i.top_score, host_buffers=True,
                        defer_fallback=self.early)
        finally:
            if self.early:                                      # (consumed by the call above unless it raised before it)
                self._c.alive_knn_arm_notify(None, None)
        if self.post is not None:
            return self.post(i.out.transpose(1, 2))             # [B, D, T] view, as match_features returns it
        return None if out is not None else i.out

    def _enqueue(self):
        if not self.zc_in:
            self.in_dev.copy_(self.src_host, non_blocking=True)
        res = self._device_step(self.out_host if self.zc_out else None)
        if res is not None:
            self.out_host.copy_(res, non_blocking=True)

    def run(self):
        """Launch on the chunk already sitting in `self.src_host`; returns immediately (see `result`)."""
        if self._dev_index is not None and torch.cuda.current_device() != self._dev_index:
            with torch.cuda.device(self._dev_index):
                rc = self._c.alive_knn_graph_launch(self._exec, self._stream_h, self._event_h)
        else:
            rc = self._c.alive_knn_graph_launch(self._exec, self._stream_h, self._event_h)
        if rc:
            M._cabi.check(rc, "alive_knn_graph_launch")
        M.launch_count += self.launches_per_call

    def submit(self, chunk: torch.Tensor):
        """Stage `chunk` (CPU) and launch; returns immediately (see `result`)."""
        self.src_host.copy_(chunk)
        self.run()

    def result(self) -> torch.Tensor:
        if self.early:
            rc = self._c.alive_knn_flag_wait(self._flag_ptr, self._flag_last, ctypes.byref(self._flag_last), 10_000_000)
            if rc:
                M._cabi.check(rc, "alive_knn_flag_wait")
            if not (self._flag_last.value & 1):
                return self.out_host.transpose(1, 2)            # every row certified: the result is complete
            # some query was left uncertified: the fallback chain was not part of the graph - enqueue it now
            i = self.inner
            with torch.cuda.device(self.lib.device), torch.cuda.stream(self.stream):
                M.run_match_fallback(i.src.shape[0], i.src.shape[2], self.lib, i.k, i.alpha, i.mode, i.variant, i.r_max, i.workspace,
                                     self.out_host, i.top_idx, i.top_score)
                self.done.record(self.stream)
        rc = self._c.alive_knn_event_wait(self._event_h)
        if rc:
            M._cabi.check(rc, "alive_knn_event_wait")
        return self.out_host.transpose(1, 2) if self.post is None else self.out_host

    def __call__(self, chunk: torch.Tensor) -> torch.Tensor:
        self.submit(chunk)
        return self.result()


class HostPipeline:
    """Throughput path with HOST buffers (offline batch conversion: features arrive from the host utterance after
    utterance, results go back): every step is  host -> device copy, match, device -> host copy  of one [B, D, T] batch,
    and the copies of neighbouring steps overlap the match - double-buffered device staging, the copies on two side
    streams, the matches back to back on the caller's stream.

        hp = HostPipeline(lib, B, T)
        for src_host, out_host in batches:         # pinned [B, D, T] in, pinned [B, T, D] out
            hp.step(src_host, out_host)            # returns at once
        hp.drain()                                 # every out_host is complete

    `out_host[b]` receives the contiguous [T, D] block the kernels write (match_features' result is its transpose).
    Results are those of match_packed on the same frames."""

    def __init__(self, lib: M.PackedFrames, B: int, T: int, k: int = 4, alpha: float = 0.0, mode: str = "auto",
                 variant: int = 0, r_max: int = M.DEFAULT_R_MAX, depth: int = 2, match_fn=None):
        """`match_fn(src_dev [B, D, T]) -> rows [B, T, D]` (a device tensor, any strides) replaces the local match - e.g.
        a ShardedLibrary's scattered match, whose collectives then run on the caller's stream between the copies."""
        dev = lib.device
        self.lib, self.k, self.alpha, self.mode, self.variant, self.r_max, self.depth = lib, k, float(alpha), mode, variant, r_max, depth
        self.match_fn = match_fn
        with torch.cuda.device(dev):
            self.s_in, self.s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
            self.src = [torch.empty((B, lib.d, T), dtype=torch.float32, device=dev) for _ in range(depth)]
            self.out = [None if match_fn is not None else torch.empty((B, T, lib.d), dtype=torch.float32, device=dev)
                        for _ in range(depth)]
            self.top_idx = torch.empty((B, T, k), dtype=torch.int64, device=dev)
            self.top_score = torch.empty((B, T, k), dtype=torch.float32, device=dev)
            self.ev_in = [torch.cuda.Event() for _ in range(depth)]      # the step's input has landed in src[i]
            self.ev_run = [torch.cuda.Event() for _ in range(depth)]     # the step's match has finished (src[i] is free, out[i] is ready)
            self.ev_out = [torch.cuda.Event() for _ in range(depth)]     # out[i] has left for the host
            self.workspace = None
        self.n = 0

    def step(self, src_host: torch.Tensor, out_host: torch.Tensor):
        if not (src_host.is_pinned() and out_host.is_pinned()):
            raise RuntimeError("alive_vc_b200: HostPipeline needs pinned (page-locked) host buffers")
        i = self.n % self.depth
        dev = self.lib.device
        main = torch.cuda.current_stream(dev)
        first = self.n < self.depth
        if not first:
            self.s_in.wait_event(self.ev_run[i])          # the match that last read src[i] is done
        else:
            self.s_in.wait_stream(main)                    # (allocation order)
        with torch.cuda.stream(self.s_in):
            self.src[i].copy_(src_host, non_blocking=True)
            self.ev_in[i].record(self.s_in)
        main.wait_event(self.ev_in[i])
        if not first:
            main.wait_event(self.ev_out[i])                # out[i] of the step before last has been read out
        if self.match_fn is not None:
            self.out[i] = self.match_fn(self.src[i])       # (kept referenced until its copy has been issued AND the slot recycled)
        else:
            info = {}
            M.run_match(self.src[i], self.lib, self.k, self.alpha, self.mode, self.variant, self.r_max, workspace=self.workspace,
                        out=self.out[i], top_idx=self.top_idx, top_score=self.top_score, info_sink=info)
            self.workspace = info["workspace"]
        self.ev_run[i].record(main)
        self.s_out.wait_event(self.ev_run[i])
        with torch.cuda.stream(self.s_out):
            out_host.copy_(self.out[i], non_blocking=True)
            self.ev_out[i].record(self.s_out)
        self.n += 1

    def drain(self):
        """The caller's stream waits for every copy issued so far (no host synchronisation)."""
        main = torch.cuda.current_stream(self.lib.device)
        for i in range(min(self.n, self.depth)):
            main.wait_event(self.ev_out[i])


class RowsContentEncoder(torch.nn.Module):
    """The producer step around a content encoder with the reference's structure
    (module/content_encoder.py:8-25: `input_layer` -> `mid_layers` -> `output_layer`, a 1x1 Conv1d to 768 channels;
    [B, n_fft/2+1, T] spectrogram -> [B, 768, T]).  The encoder's own layers run untouched; only its LAST layer is
    evaluated channels-last - the same 1x1 convolution written as a linear map on [B, T, C] - so the features leave
    the encoder as row-major [B, T, 768] frames, the layout the match consumes, and K1 (alive_knn_pack) runs as the
    encoder's epilogue: `forward` returns the packed queries for `match_packed_queries` (no per-call query pack
    inside the match, no transposed views either side), `rows` the plain [B, T, 768] frames for `match_rows`.
    An encoder without that structure is wrapped as `encoder(x).transpose(1, 2)` (a view; K1 takes the strides)."""

    def __init__(self, encoder: torch.nn.Module):
        super().__init__()
        self.encoder = encoder

    def rows(self, spec: torch.Tensor) -> torch.Tensor:
        enc = self.encoder
        out_layer = getattr(enc, "output_layer", None)
        if (isinstance(out_layer, torch.nn.Conv1d) and out_layer.kernel_size == (1,) and out_layer.groups == 1 and
                hasattr(enc, "input_layer") and hasattr(enc, "mid_layers")):
            x = enc.mid_layers(enc.input_layer(spec))                                   # content_encoder.py:22-23
            return torch.nn.functional.linear(x.transpose(1, 2), out_layer.weight[:, :, 0], out_layer.bias)   # :24
        return enc(spec).transpose(1, 2)

    def forward(self, spec: torch.Tensor, fmt=None) -> M.PackedFrames:
        """`fmt`: the 16-bit format of the libraries the frames will be matched against (None = matching.SCREEN_FORMAT)"""
        with torch.no_grad():
            rows = self.rows(spec)
            B, T, D = rows.shape
            # [D, B*T] view with stride_d == 1: K1's row-major kernel; both planes (queries are small)
            return M.pack_frames(rows.reshape(B * T, D).float().t(), refine=True, fmt=fmt)
