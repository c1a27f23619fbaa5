"""Host side of the kNN voice-library matching path.

Mirrors the reference's functional interface
    module/common.py:96-109   match_features(source, reference, k=4, alpha=0.0)
with the same argument meaning, return layout and error behaviour, and runs it
as hand-written sm_100a kernels behind the C ABI in include/alive_knn.h:

    pack (K1)  ->  fused tcgen05 similarity + running top lists (K2)
               ->  certificate + prune (K2b)  ->  exact fp32/fp64 rescoring (K3)
               ->  exact scan for uncertified queries  ->  gather + mean + blend (K4)

PyTorch is used for device memory, streams and autograd plumbing only.  There
is no CPU path: tensors must live on a CUDA device and the in-tree
libalive_knn.so must be built, otherwise a RuntimeError is raised.
"""
from __future__ import annotations

import contextlib
import ctypes
import weakref
from dataclasses import dataclass, field
from typing import Optional

import torch

from . import _cabi

LIST_LEN = 8            # ALIVE_KNN_LIST_LEN
MAX_K = 64              # ALIVE_KNN_MAX_K
DEFAULT_R_MAX = 256     # survivors rescored per query before falling back to the exact scan
EXACT_BELOW_N = 0       # the tensor-core screen handles any library size (exact scan: k > 8 or d % 64 != 0)

_sm_count: dict = {}


def _num_sms(device: torch.device) -> int:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _sm_count:
        _sm_count[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _sm_count[idx]


def _stream_ptr(device=None) -> int:
    """torch's current stream ON `device` (default: the current device)."""
    return torch.cuda.current_stream(device).cuda_stream


def _on(device):
    """The C ABI launches on the CURRENT CUDA device and never switches it: make the device the buffers
    live on current for the duration of the call (a no-op context when it already is)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if torch.cuda.current_device() == idx:
        return contextlib.nullcontext()
    return torch.cuda.device(idx)


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(
            f"alive_vc_b200: `{name}` is on {t.device}; this implementation has no CPU path "
            "(sm_100a CUDA kernels only) - move the tensors to a B200 device")


@dataclass
class PackedFrames:
    """Frames in the layout the kernels consume (see DESIGN.md "Data layout in HBM")."""
    n: int
    d: int
    raw: torch.Tensor        # [n,d] float32, raw frames, row-major
    norms: torch.Tensor      # [n]   float32
    packed: torch.Tensor     # [n,d] bfloat16, frames / norm
    err: torch.Tensor        # [n]   float32, ||bf16(x/|x|) - x/|x|||_2
    stats: torch.Tensor      # [4]   int32 (uint32 bit patterns): max err, non-finite row count, max err2, reserved
    row_base: int = 0        # global index of frame 0 (sharded libraries)
    items: int = 1           # > 1: `items` independent libraries of n // items frames each, back to back
    lo: Optional[torch.Tensor] = None     # [n,d] second plane rn16(x/|x| - packed): refined collect pass
    err2: Optional[torch.Tensor] = None   # [n] float32 ||x/|x| - packed - lo||_2
    format: int = _cabi.FORMAT_BF16       # 16-bit format of packed / lo (see SCREEN_FORMAT)
    _handle: object = field(default=None, repr=False)

    @property
    def n_item(self) -> int:
        return self.n // self.items

    def handle(self) -> "_cabi.Library":
        """alive_knn_library_t view of these buffers (rebuilt if row_base changed)."""
        h = self._handle
        if h is None or h.row_base != self.row_base or h.items != self.items:
            h = _cabi.Library(self.packed.data_ptr(), self.raw.data_ptr(), self.norms.data_ptr(),
                              self.stats.data_ptr(), self.n // self.items, self.d, self.row_base, self.items,
                              self.lo.data_ptr() if self.lo is not None else None, self.format)
            self._handle = h
        return h

    @property
    def device(self):
        return self.raw.device

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.raw, self.norms, self.packed, self.err, self.stats, self.lo,
                                                          self.err2) if t is not None)

    def rows(self, lo_row: int, hi_row: int) -> "PackedFrames":
        """frames [lo_row, hi_row) as a row shard (views, global indices kept through row_base)"""
        return PackedFrames(n=hi_row - lo_row, d=self.d, raw=self.raw[lo_row:hi_row], norms=self.norms[lo_row:hi_row],
                            packed=self.packed[lo_row:hi_row], err=self.err[lo_row:hi_row], stats=self.stats,
                            row_base=self.row_base + lo_row,
                            lo=self.lo[lo_row:hi_row] if self.lo is not None else None,
                            err2=self.err2[lo_row:hi_row] if self.err2 is not None else None, format=self.format)


# The second bf16 plane costs 2 B per element (+33 % of a packed library) and buys the refined collect pass: clustered
# libraries (near-duplicate frames: silence, sustained vowels) stay off the exhaustive scan.  "auto": on for a single
# library when the plane fits comfortably (at most a quarter of the free device memory), off for sets of per-speaker
# libraries (pack_libraries: BASELINE cfg5 fills the GPU without it); True / False force it.
REFINE_DEFAULT = "auto"

# 16-bit format of the packed planes (the tensor-core operands); results are identical either way (everything after
# the screen is exact), only the speed differs:
#   "bf16"  the format the project brief names, and the faster one on well-spread libraries: fp16 multipliers draw more
#           power, and under the 1 kW cap the fp16 screen of BASELINE cfg4 runs 7 % slower (1340 vs 1440 TFLOP/s, SM
#           clock 1230 vs 1350 MHz - measured);
#   "fp16"  normalised frames live in [-1, 1], IEEE half rounds them 8x finer at the same tensor-core rate: the
#           certificate's band is 8x narrower, so CLUSTERED libraries (near-duplicate frames: silence, sustained vowels)
#           certify in the first pass instead of taking the collect pass (cfg1 shape, clusters of 1000 frames at noise
#           0.2: 0.17 ms against 0.53 ms);
#   "auto"  (default for pack_library / pack_frames of a single library) pack as bf16, probe the library with 256 of its
#           own frames, repack as fp16 when more than an eighth of them cannot be certified.  Costs one small screen and
#           one host synchronisation at PACK time.
# Buffers that are filled piecewise (alloc_packed + pack_into) and query packs use the explicit format ("auto" = bf16).
SCREEN_FORMAT = "auto"
_FORMATS = {"bf16": _cabi.FORMAT_BF16, "fp16": _cabi.FORMAT_FP16, "auto": _cabi.FORMAT_BF16,
            _cabi.FORMAT_BF16: _cabi.FORMAT_BF16, _cabi.FORMAT_FP16: _cabi.FORMAT_FP16}
_PLANE_DTYPE = {_cabi.FORMAT_BF16: torch.bfloat16, _cabi.FORMAT_FP16: torch.float16}
AUTO_PROBE_MIN_FRAMES = 4096
AUTO_PROBE_QUERIES = 256


def _format_of(fmt) -> int:
    return _FORMATS[SCREEN_FORMAT if fmt is None else fmt]


def probe_uncertified_fraction(lib: "PackedFrames", samples: int = AUTO_PROBE_QUERIES) -> float:
    """Fraction of `samples` evenly spaced LIBRARY frames, matched against the library itself (k = 4: the frame and its
    three nearest neighbours), that the first-pass certificate cannot clear - the signature of a clustered library.
    Synchronises the host (a pack-time decision)."""
    global last_info
    if lib.items != 1:
        raise RuntimeError("probe_uncertified_fraction expects a single library")
    m = min(samples, lib.n)
    idx = torch.linspace(0, lib.n - 1, m, device=lib.device).long()
    q = PackedFrames(n=m, d=lib.d, raw=lib.raw[idx], norms=lib.norms[idx], packed=lib.packed[idx], err=lib.err[idx],
                     stats=lib.stats, lo=lib.lo[idx] if lib.lo is not None else None,
                     err2=lib.err2[idx] if lib.err2 is not None else None, format=lib.format)
    saved = last_info
    match_packed_queries(q, lib, min(4, lib.n), 0.0, mode="screen", want_out=False)
    frac = last_info.fallback_queries() / float(m)
    last_info = saved
    return frac


def _want_refine(refine, n: int, d: int, device, items: int = 1) -> bool:
    if refine is None:
        refine = REFINE_DEFAULT
    if refine == "auto":
        if items > 1 or d % 64 != 0:
            return False
        try:
            free, _ = torch.cuda.mem_get_info(device)
        except Exception:
            return True
        return n * d * 2 <= free // 4
    return bool(refine)


def alloc_packed(n: int, d: int, device, refine=None, items: int = 1, fmt=None) -> PackedFrames:
    want_lo = _want_refine(refine, n, d, device, items)
    f = _format_of(fmt)
    return PackedFrames(
        n=n, d=d,
        raw=torch.empty((n, d), dtype=torch.float32, device=device),
        norms=torch.empty((n,), dtype=torch.float32, device=device),
        packed=torch.empty((n, d), dtype=_PLANE_DTYPE[f], device=device),
        err=torch.empty((n,), dtype=torch.float32, device=device),
        stats=torch.zeros((4,), dtype=torch.int32, device=device),
        items=items,
        lo=torch.empty((n, d), dtype=_PLANE_DTYPE[f], device=device) if want_lo else None,
        err2=torch.empty((n,), dtype=torch.float32, device=device) if want_lo else None,
        format=f,
    )


def pack_into(dst: PackedFrames, row0: int, frames_dn: torch.Tensor):
    """K1 on a [D, n] float32 view (any strides) -> rows [row0, row0+n) of `dst`."""
    lib = _cabi.load()
    d, n = frames_dn.shape
    if n == 0:
        return
    assert frames_dn.dtype == torch.float32 and frames_dn.is_cuda
    assert d == dst.d and row0 + n <= dst.n and frames_dn.device == dst.device
    with _on(dst.device):
        rc = lib.alive_knn_pack(
            frames_dn.data_ptr(), n, d, frames_dn.stride(1), frames_dn.stride(0),
            dst.raw[row0:].data_ptr(), dst.norms[row0:].data_ptr(), dst.packed[row0:].data_ptr(),
            dst.err[row0:].data_ptr(), dst.stats.data_ptr(),
            dst.lo[row0:].data_ptr() if dst.lo is not None else None,
            dst.err2[row0:].data_ptr() if dst.err2 is not None else None, dst.format, _stream_ptr(dst.device))
    _cabi.check(rc, "alive_knn_pack")
    _count(1)


def pack_frames(frames_dn: torch.Tensor, refine=None, fmt=None) -> PackedFrames:
    """Normalise-and-pack a [D, N] float32 CUDA view (the reference's channel-major
    layout, any strides).  Done ONCE per library (generate_voice_library.py / load time)
    instead of once per call as common.py:101-104 does.  `refine`: also store the second bf16 plane
    (True / False / None = REFINE_DEFAULT, see there); `fmt`: "bf16" / "fp16" / "auto" (None = SCREEN_FORMAT)."""
    _require_cuda(frames_dn, "frames")
    if frames_dn.dtype != torch.float32:
        frames_dn = frames_dn.float()
    d, n = frames_dn.shape
    choice = SCREEN_FORMAT if fmt is None else fmt
    out = alloc_packed(n, d, frames_dn.device, refine, fmt=choice)
    pack_into(out, 0, frames_dn)
    if choice == "auto" and n >= AUTO_PROBE_MIN_FRAMES and d % 64 == 0:
        bad_rows = int(out.stats[1].item())
        if bad_rows == 0 and probe_uncertified_fraction(out) > 0.125:
            keep_lo = out.lo is not None
            del out                                   # clustered: the 8x finer rounding of fp16 pays for itself
            out = alloc_packed(n, d, frames_dn.device, keep_lo, fmt="fp16")
            pack_into(out, 0, frames_dn)
    return out


def pack_library(reference: torch.Tensor, refine=None, fmt=None) -> PackedFrames:
    """[1, D, N] (or [D, N]) library tensor -> PackedFrames."""
    if reference.dim() == 3:
        if reference.shape[0] != 1:
            raise RuntimeError("pack_library expects a single library [1, D, N] (see pack_libraries)")
        reference = reference[0]
    return pack_frames(reference, refine, fmt)


def pack_libraries(reference: torch.Tensor, refine=None, fmt=None) -> PackedFrames:
    """[B, D, N] -> ONE PackedFrames holding B independent libraries of N frames back to back
    (items = B): batch item b of a query tensor is matched against library b only, all of them in
    a single launch (BASELINE cfg5; train_decoder.py:134-135's per-utterance libraries)."""
    _require_cuda(reference, "reference")
    if reference.dim() != 3:
        raise RuntimeError("pack_libraries expects [B, D, N]")
    if reference.dtype != torch.float32:
        reference = reference.float()
    B, D, N = reference.shape
    out = alloc_packed(B * N, D, reference.device, refine, items=B, fmt=fmt)
    for b in range(B):
        pack_into(out, b * N, reference[b])
    return out


# ---------------------------------------------------------------------------------------
# library lifecycle: the reference's own checkpoint + a sidecar with the packed layout
# ---------------------------------------------------------------------------------------
PACKED_FORMAT_VERSION = 2
SIDECAR_SUFFIX = ".alive_knn"


def _fingerprint(raw: torch.Tensor):
    """Cheap content check tying a sidecar to its tokens file (fp64 sum and sum of squares, on the device)."""
    r = raw.double()
    return [float(r.sum()), float((r * r).sum())]


def save_packed_library(lib: PackedFrames, path: str, include_legacy_tokens: bool = True):
    """Persist a packed library as TWO files:

    `path`               exactly what generate_voice_library.py:42 / fine_tune.py:199 write: the state dict
                         `{"tokens": [1, D, N] float32}` and nothing else, so the UNMODIFIED reference reads it with
                         its strict `VL.load_state_dict(torch.load(path))` (inference.py:81, realtime_inference.py:93,
                         fine_tune.py:126) - given `VoiceLibrary(num_tokens=N)`.  A set of B per-speaker libraries
                         (`pack_libraries`, items = B) is stored as tokens [B, D, N].
    `path + ".alive_knn"` the packed layout K1 produced (bf16 rows, norms, error norms, stats) plus n, d, items,
                         row_base and a fingerprint of the tokens; the raw fp32 rows are NOT stored twice (they are
                         the transposed tokens).
    `include_legacy_tokens=False` skips the first file (the sidecar alone cannot be loaded)."""
    n_item = lib.n_item
    tokens = lib.raw.view(lib.items, n_item, lib.d).transpose(1, 2).contiguous().cpu()     # [items, D, N]
    if include_legacy_tokens:
        torch.save({"tokens": tokens}, path)
    torch.save({
        "alive_knn_packed_version": PACKED_FORMAT_VERSION,
        "n": lib.n, "d": lib.d, "row_base": lib.row_base, "items": lib.items, "format": lib.format,
        "norms": lib.norms.cpu(), "packed": lib.packed.cpu(), "err": lib.err.cpu(), "stats": lib.stats.cpu(),
        "lo": lib.lo.cpu() if lib.lo is not None else None, "err2": lib.err2.cpu() if lib.err2 is not None else None,
        "fingerprint": _fingerprint(lib.raw),
    }, path + SIDECAR_SUFFIX)


def load_packed_library(path: str, device="cuda") -> PackedFrames:
    """Inverse of save_packed_library.  `path` is a reference voice-library checkpoint (`{"tokens": [1, D, N]}`,
    or [B, D, N] for a set of libraries); when the sidecar `path + ".alive_knn"` exists and its fingerprint matches
    the tokens, the packed layout is taken from it, otherwise the tokens are packed on the fly (K1) - so every
    checkpoint the reference wrote loads too."""
    import os
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("alive_vc_b200: packed libraries live on a CUDA device (no CPU path)")
    blob = torch.load(path, map_location="cpu", weights_only=True)
    if "alive_knn_packed_version" in blob:        # round-1 single-file format (packed arrays next to `tokens`)
        if blob["alive_knn_packed_version"] != 1:
            raise RuntimeError(f"{path}: unsupported packed format version {blob['alive_knn_packed_version']}")
        stats = torch.zeros((4,), dtype=torch.int32)
        stats[:2] = blob["stats"][:2]
        return PackedFrames(n=int(blob["n"]), d=int(blob["d"]), raw=blob["raw"].to(dev), norms=blob["norms"].to(dev),
                            packed=blob["packed"].to(dev), err=blob["err"].to(dev), stats=stats.to(dev),
                            row_base=int(blob["row_base"]))
    if "tokens" not in blob:
        raise RuntimeError(f"{path}: not a voice-library checkpoint (no `tokens` key)")
    tokens = blob["tokens"]
    if tokens.dim() != 3:
        raise RuntimeError(f"{path}: tokens must be [1, D, N] (or [B, D, N]), got {tuple(tokens.shape)}")
    items, d, n_item = tokens.shape
    side = path + SIDECAR_SUFFIX
    if os.path.exists(side):
        sc = torch.load(side, map_location="cpu", weights_only=True)
        raw = tokens.to(dev).float().transpose(1, 2).contiguous().view(items * n_item, d)
        ok = (sc.get("alive_knn_packed_version") == PACKED_FORMAT_VERSION and int(sc["n"]) == items * n_item and
              int(sc["d"]) == d and int(sc["items"]) == items and sc["fingerprint"] == _fingerprint(raw))
        if ok:
            return PackedFrames(n=items * n_item, d=d, raw=raw, norms=sc["norms"].to(dev), packed=sc["packed"].to(dev),
                                err=sc["err"].to(dev), stats=sc["stats"].to(dev), row_base=int(sc["row_base"]),
                                items=items, lo=sc["lo"].to(dev) if sc.get("lo") is not None else None,
                                err2=sc["err2"].to(dev) if sc.get("err2") is not None else None,
                                format=int(sc.get("format", _cabi.FORMAT_BF16)))
        del raw      # stale sidecar (the tokens were edited since): pack again
    tok = tokens.to(dev)
    return pack_library(tok) if items == 1 else pack_libraries(tok)


# ---------------------------------------------------------------------------------------
# pack cache: invisible to callers (SURVEY §8(b) "Ownership").  Keyed on the tensor OBJECT that owns the
# memory (the base of a view) plus the view's geometry, validated against the version counter.
# ---------------------------------------------------------------------------------------
_pack_cache: dict = {}
_PACK_CACHE_MAX = 16
# Libraries below this many elements are re-packed on every call instead of cached (K1 on 512 x 768 tokens is
# one launch of a few microseconds): writes through `.data` (`VL.tokens.data[:, :, n] = t`,
# generate_voice_library.py:38), which do NOT bump the version counter, are then always seen - the default
# 512-token VoiceLibrary included.  Larger tensors written through `.data` (or by a custom kernel / a graph
# replay) need an explicit clear_pack_cache(tensor) / VoiceLibrary.invalidate().
PACK_CACHE_MIN_ELEMENTS = 1 << 20


def _version_of(t: torch.Tensor):
    try:
        return t._version
    except RuntimeError:          # "Inference tensors do not track version counter": nothing to validate against
        return None


def _owner_of(t: torch.Tensor) -> torch.Tensor:
    """The object whose lifetime and version counter govern `t`: its base when `t` is a view."""
    try:
        base = t._base
    except RuntimeError:
        base = None
    return t if base is None else base


def _geometry(t: torch.Tensor):
    return (t.data_ptr(), tuple(t.shape), tuple(t.stride()), t.dtype, t.device.index)


def _cached(reference: torch.Tensor, tag, build) -> PackedFrames:
    version = _version_of(reference)
    if version is None or reference.numel() < PACK_CACHE_MIN_ELEMENTS:
        return build()            # inference tensors (realtime_inference.py:143) and small libraries: never cached
    owner = _owner_of(reference)
    oid = (id(owner), tag, _geometry(reference))
    ent = _pack_cache.get(oid)
    if ent is not None:
        ref, old_version, packed = ent
        if ref() is owner and old_version == version:
            return packed
        del _pack_cache[oid]
    packed = build()
    if len(_pack_cache) >= _PACK_CACHE_MAX:
        for dead in [k for k, (r, _, _) in _pack_cache.items() if r() is None]:
            del _pack_cache[dead]
        while len(_pack_cache) >= _PACK_CACHE_MAX:
            del _pack_cache[next(iter(_pack_cache))]

    def _drop(_ref, owner_id=id(owner)):
        for key in [k for k in _pack_cache if k[0] == owner_id]:
            _pack_cache.pop(key, None)

    try:
        _pack_cache[oid] = (weakref.ref(owner, _drop), version, packed)
    except TypeError:
        pass
    return packed


def cached_pack(reference: torch.Tensor, frames_dn: torch.Tensor, tag=0) -> PackedFrames:
    """Pack `frames_dn` (the [D, N] float32 frames of `reference`) unless the same view of the same tensor
    object was packed before and its version counter has not moved since."""
    return _cached(reference, tag, lambda: pack_frames(frames_dn))


def cached_pack_many(reference: torch.Tensor, reference_bdn: torch.Tensor) -> PackedFrames:
    """pack_libraries with the same invisible cache as cached_pack."""
    return _cached(reference, "many", lambda: pack_libraries(reference_bdn))


def clear_pack_cache(tensor: Optional[torch.Tensor] = None):
    """Forget every cached packed library - or, given a tensor, the ones made from it or from views of the same
    base.  Needed after writes the version counter does not see (`t.data[...] = x`, custom kernels, CUDA-graph
    replays writing into the library) when the tensor holds at least PACK_CACHE_MIN_ELEMENTS elements."""
    if tensor is None:
        _pack_cache.clear()
        return
    owner_id = id(_owner_of(tensor))
    for key in [k for k in _pack_cache if k[0] == owner_id]:
        _pack_cache.pop(key, None)


# ---------------------------------------------------------------------------------------
# top-k search
# ---------------------------------------------------------------------------------------
@dataclass
class SearchInfo:
    """Per-call bookkeeping (device tensors; reading them synchronises)."""
    mode: str
    plan: Optional[dict] = None
    fb_count: Optional[torch.Tensor] = None     # [items] int32: queries the first screen could not certify
    exact_count: Optional[torch.Tensor] = None  # [1] int32: of those, queries that needed the exhaustive scan
    collect: bool = False                       # a collect pass ran between the two (alive_knn_match, off[6] area)
    sel_n: Optional[torch.Tensor] = None        # [T] int32: survivors rescored per query (-1 = exact scan)
    launches: int = 0

    def fallback_queries(self) -> int:
        return int(self.fb_count.sum().item()) if self.fb_count is not None else 0

    def exact_scan_queries(self) -> int:
        if self.exact_count is None or not self.collect:
            return self.fallback_queries()
        return int(self.exact_count.sum().item())


last_info: Optional[SearchInfo] = None

# bookkeeping used by bench.py: kernels launched through the C ABI, and an optional list that
# receives (start_event, end_event) pairs bracketing every alive_knn_search launch
launch_count: int = 0
search_events: Optional[list] = None


def _count(n: int):
    global launch_count
    launch_count += n


def make_plan(t: int, n: int, d: int, device, variant: int = 0, fmt=None) -> _cabi.Plan:
    """`fmt`: the 16-bit format of the operands the plan will be launched on (PackedFrames.format; None = SCREEN_FORMAT,
    what freshly packed frames have)"""
    plan = _cabi.Plan()
    rc = _cabi.load().alive_knn_plan(t, n, d, _num_sms(device), variant, ctypes.byref(plan))
    _cabi.check(rc, "alive_knn_plan")
    plan.format = _format_of(fmt)
    return plan


def exact_topk(q: PackedFrames, lib: PackedFrames, k: int, top_score=None, top_idx=None,
               q_list=None, q_count=None):
    """common.py:102-105 by exhaustive exact scan (fp64-accumulated similarities)."""
    c = _cabi.load()
    dev = q.device
    t = q.n
    if top_score is None:
        top_score = torch.empty((t, k), dtype=torch.float32, device=dev)
        top_idx = torch.empty((t, k), dtype=torch.int64, device=dev)
    ws_bytes = c.alive_knn_exact_workspace_bytes(t, lib.n, k, 1)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    with _on(dev):
        rc = c.alive_knn_exact(q.raw.data_ptr(), q.norms.data_ptr(), t, lib.raw.data_ptr(), lib.norms.data_ptr(),
                               lib.n, lib.d, k,
                               q_list.data_ptr() if q_list is not None else None,
                               q_count.data_ptr() if q_count is not None else None,
                               lib.row_base, ws.data_ptr(), top_score.data_ptr(), top_idx.data_ptr(), 0.0, None, 1,
                               _stream_ptr(dev))
    _cabi.check(rc, "alive_knn_exact")
    _count(2)
    return top_score, top_idx


def search_topk(q: PackedFrames, lib: PackedFrames, k: int, mode: str = "auto",
                r_max: int = DEFAULT_R_MAX, variant: int = 0):
    """Exact top-k (score desc, frame asc) of every query frame against the library.
    Returns (top_score [T,k] float32, top_idx [T,k] int64 incl. lib.row_base)."""
    global last_info
    c = _cabi.load()
    if q.d != lib.d:
        raise RuntimeError(f"feature dims differ: queries {q.d}, library {lib.d}")
    if k < 1 or k > lib.n:
        raise RuntimeError("selected index k out of range")   # torch.topk's message (common.py:105)
    if k > MAX_K:
        raise RuntimeError(f"alive_vc_b200 supports k <= {MAX_K} (got {k})")
    t, n, d = q.n, lib.n, lib.d
    dev = q.device
    if mode == "auto":
        mode = "exact" if (k > LIST_LEN or n < EXACT_BELOW_N or d % 64 != 0) else "screen"
    if mode == "exact":
        top_score, top_idx = exact_topk(q, lib, k)
        last_info = SearchInfo(mode="exact", launches=2)
        return top_score, top_idx
    if mode != "screen":
        raise ValueError(f"unknown mode {mode!r}")
    if k > LIST_LEN:
        raise RuntimeError(f"the screened path needs k <= {LIST_LEN}")

    if q.format != lib.format:
        raise RuntimeError("queries and library must be packed in the same 16-bit format")
    plan = make_plan(t, n, d, dev, variant, lib.format)
    with _on(dev):
        return _search_topk_screen(c, q, lib, k, r_max, plan, dev)


def _search_topk_screen(c, q, lib, k, r_max, plan, dev):
    global last_info
    t, d = q.n, lib.d
    stream = _stream_ptr(dev)
    cand_score = torch.empty((t, plan.lists, LIST_LEN), dtype=torch.float32, device=dev)
    cand_idx = torch.empty((t, plan.lists, LIST_LEN), dtype=torch.int32, device=dev)
    if search_events is not None:
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record()
    rc = c.alive_knn_search(q.packed.data_ptr(), lib.packed.data_ptr(), ctypes.byref(plan),
                            cand_score.data_ptr(), cand_idx.data_ptr(), stream)
    _cabi.check(rc, "alive_knn_search")
    if search_events is not None:
        ev1.record()
        search_events.append((ev0, ev1))

    sel_idx = torch.empty((t, r_max), dtype=torch.int32, device=dev)
    sel_n = torch.empty((t,), dtype=torch.int32, device=dev)
    fb_list = torch.empty((t,), dtype=torch.int32, device=dev)
    fb_count = torch.empty((1,), dtype=torch.int32, device=dev)
    rc = c.alive_knn_prune(cand_score.data_ptr(), cand_idx.data_ptr(), t, plan.lists, k, d, q.err.data_ptr(),
                           q.norms.data_ptr(), lib.stats.data_ptr(), r_max, sel_idx.data_ptr(), sel_n.data_ptr(),
                           fb_list.data_ptr(), fb_count.data_ptr(), stream)
    _cabi.check(rc, "alive_knn_prune")

    top_score = torch.empty((t, k), dtype=torch.float32, device=dev)
    top_idx = torch.empty((t, k), dtype=torch.int64, device=dev)
    rc = c.alive_knn_rescore(q.raw.data_ptr(), q.norms.data_ptr(), t, lib.raw.data_ptr(), lib.norms.data_ptr(), d,
                             sel_idx.data_ptr(), sel_n.data_ptr(), r_max, k, lib.row_base,
                             top_score.data_ptr(), top_idx.data_ptr(), stream)
    _cabi.check(rc, "alive_knn_rescore")
    _count(3)   # search + prune + rescore
    # queries the certificate could not clear: exhaustive scan (device-side list, no host sync)
    exact_topk(q, lib, k, top_score, top_idx, fb_list, fb_count)
    last_info = SearchInfo(mode="screen", plan=plan.as_dict(), fb_count=fb_count, sel_n=sel_n, launches=7)
    return top_score, top_idx


def gather_mean(lib: PackedFrames, top_idx: torch.Tensor, q: PackedFrames, alpha: float, out: torch.Tensor):
    """K4: common.py:107-109 into `out` [T,D] float32."""
    t, k = top_idx.shape
    with _on(out.device):
        rc = _cabi.load().alive_knn_gather_mean(lib.raw.data_ptr(), lib.n, lib.d, top_idx.data_ptr(), t, k,
                                                q.raw.data_ptr(), q.norms.data_ptr() if q.norms is not None else None,
                                                float(alpha), out.data_ptr(), _stream_ptr(out.device))
    _cabi.check(rc, "alive_knn_gather_mean")
    _count(1)
    return out


def pack_queries(source: torch.Tensor, fmt=None) -> PackedFrames:
    """[B, D, T] float32 CUDA -> PackedFrames with B*T rows (row b*T + t), both planes; `fmt` = the format of the
    library they will be matched against (None = SCREEN_FORMAT)."""
    B, D, T = source.shape
    out = alloc_packed(B * T, D, source.device, refine=True, fmt=fmt)
    for b in range(B):
        pack_into(out, b * T, source[b])
    return out


_MODES = {"auto": 0, "screen": 1, "exact": 2}


def _collect_launches(refined: bool) -> int:
    """kernels of a screened call with the collect pass behind it (the query pack not counted): search, finish,
    [refine_prep], collect search, collect_rescore, exact_partial / _rows / _final"""
    return 7 if refined else 6


_layout_cache: dict = {}


def _layout(rows: int, lib: PackedFrames, k: int, r_max: int, mode: int, variant: int, device):
    """Workspace layout of alive_knn_match for this shape (a pure function of its arguments: memoised, so a
    steady-state call makes ONE C call)."""
    key = (rows, lib.n_item, lib.d, k, r_max, mode, _num_sms(device), variant, lib.items)
    hit = _layout_cache.get(key)
    if hit is not None:
        return hit
    off = (ctypes.c_int64 * 14)()
    rc = _cabi.load().alive_knn_match_layout(*key[:8], lib.items, off)
    _cabi.check(rc, "alive_knn_match_layout")
    if len(_layout_cache) > 256:
        _layout_cache.clear()
    _layout_cache[key] = list(off)
    return _layout_cache[key]


def run_match(source: torch.Tensor, lib: PackedFrames, k: int = 4, alpha: float = 0.0, mode: str = "auto",
              variant: int = 0, r_max: int = DEFAULT_R_MAX, want_out: bool = True, workspace=None,
              out=None, top_idx=None, top_score=None, info_sink: Optional[dict] = None, host_buffers: bool = False,
              defer_fallback: bool = False):
    """The whole path in ONE C call (alive_knn_match): pack the B*T query frames of `source`
    [B,D,T] (any strides, float32, CUDA), search, certify, rescore, exact-scan the uncertified,
    gather+mean+blend.  Returns (out [B,T,D] or None, top_idx [B,T,k] int64, top_score [B,T,k]).
    `info_sink` (a dict) receives the workspace and its layout offsets (q_raw at offsets[0], q_norm at [1]).
    `host_buffers`: `source` (and a given `out`) may be PINNED HOST tensors - the kernels read / write them in place
    over PCIe (unified addressing; HostStreamingMatcher's zero-copy chunk path); everything else lives on lib.device.
    `defer_fallback`: enqueue pack -> search -> finish only (ALIVE_KNN_MODE_DEFER_FALLBACK); the caller must learn whether
    a query was left uncertified (alive_knn_arm_notify) and then call run_match_fallback with the same arguments."""
    global last_info
    c = _cabi.load()
    B, D, T = source.shape
    if D != lib.d:
        raise RuntimeError(f"feature dims differ: queries {D}, library {lib.d}")
    if not isinstance(k, int) or k < 1 or k > lib.n_item:
        raise RuntimeError("selected index k out of range")   # torch.topk's message (common.py:105)
    if k > MAX_K:
        raise RuntimeError(f"alive_vc_b200 supports k <= {MAX_K} (got {k})")
    if lib.items > 1 and lib.items != B:
        raise RuntimeError(f"a packed set of {lib.items} libraries needs a query batch of {lib.items} (got {B})")
    assert source.dtype == torch.float32
    if host_buffers:
        for tns in (source, out):
            if tns is not None and not tns.is_cuda and not tns.is_pinned():
                raise RuntimeError("alive_vc_b200: host buffers must be pinned (page-locked) to be read by the kernels")
        dev = lib.device
        if source.is_cuda and source.device != dev:
            raise RuntimeError(f"source is on {source.device} but the packed library is on {dev}")
    else:
        _require_cuda(source, "source")
        dev = source.device
        if dev != lib.device:
            raise RuntimeError(f"source is on {dev} but the packed library is on {lib.device}")
    rows = B * T
    m = _MODES[mode]
    if m == 0:
        m = 2 if (k > LIST_LEN or lib.n_item < EXACT_BELOW_N or lib.d % 64 != 0) else 1
    off = _layout(rows, lib, k, r_max, m, variant, dev)
    if workspace is None:
        workspace = torch.empty((off[11],), dtype=torch.uint8, device=dev)
    if want_out and out is None:
        out = torch.empty((B, T, D), dtype=torch.float32, device=dev)
    if top_idx is None:
        top_idx = torch.empty((B, T, k), dtype=torch.int64, device=dev)
        top_score = torch.empty((B, T, k), dtype=torch.float32, device=dev)
    ev0 = ev1 = None
    with _on(dev):
        if search_events is not None and m == 1:
            ev0 = torch.cuda.Event(enable_timing=True)
            ev1 = torch.cuda.Event(enable_timing=True)
            ev0.record()      # forces creation of the cudaEvent_t handles; re-recorded inside the C call
            ev1.record()
            search_events.append((ev0, ev1))
        rc = c.alive_knn_match(source.data_ptr(), B, T, source.stride(0), source.stride(2), source.stride(1),
                               ctypes.byref(lib.handle()), k, float(alpha), r_max,
                               m | (_cabi.MODE_DEFER_FALLBACK if defer_fallback else 0), _num_sms(dev), variant,
                               workspace.data_ptr(), workspace.numel(), out.data_ptr() if want_out else None,
                               top_idx.data_ptr(), top_score.data_ptr(),
                               ev0.cuda_event if ev0 is not None else None,
                               ev1.cuda_event if ev1 is not None else None, _stream_ptr(dev))
    _cabi.check(rc, "alive_knn_match")
    n_launch = 1 + ((_collect_launches(lib.lo is not None) if off[7] > off[6] else 4) if m == 1 else 2)
    if defer_fallback and m == 1:
        n_launch = 3                                     # pack, search, finish
    _count(n_launch)
    last_info = SearchInfo(mode="screen" if m == 1 else "exact",
                           fb_count=workspace[off[9]:off[9] + 4 * lib.items].view(torch.int32),
                           exact_count=workspace[off[9] + 4 * lib.items:off[9] + 8 * lib.items].view(torch.int32),
                           collect=m == 1 and off[7] > off[6],
                           sel_n=workspace[off[7]:off[7] + 4 * rows].view(torch.int32) if m == 1 else None,
                           launches=n_launch)
    last_info._workspace = workspace
    last_info._offsets = off
    if info_sink is not None:          # callers on several threads cannot rely on the module-level last_info
        info_sink["workspace"], info_sink["offsets"] = workspace, off
    return (out if want_out else None), top_idx, top_score


def run_match_fallback(B: int, T: int, lib: PackedFrames, k: int, alpha: float, mode: str, variant: int, r_max: int,
                       workspace: torch.Tensor, out, top_idx: torch.Tensor, top_score: torch.Tensor):
    """The second half of a run_match(..., defer_fallback=True) call: the fallback chain for the queries its finish
    kernel could not certify (alive_knn_match_fallback; same sizes, same workspace, same result buffers)."""
    c = _cabi.load()
    m = _MODES[mode]
    if m == 0:
        m = 2 if (k > LIST_LEN or lib.n_item < EXACT_BELOW_N or lib.d % 64 != 0) else 1
    if m != 1:
        return                      # the exact mode has no second half
    dev = lib.device
    with _on(dev):
        rc = c.alive_knn_match_fallback(None, None, None, None, None, None, B, T, ctypes.byref(lib.handle()), k, float(alpha),
                                        r_max, m, _num_sms(dev), variant, workspace.data_ptr(), workspace.numel(),
                                        out.data_ptr() if out is not None else None, top_idx.data_ptr(),
                                        top_score.data_ptr(), _stream_ptr(dev))
    _cabi.check(rc, "alive_knn_match_fallback")
    off = _layout(B * T, lib, k, r_max, m, variant, dev)
    _count((_collect_launches(lib.lo is not None) if off[7] > off[6] else 4) - 2)


def match_packed(source: torch.Tensor, lib: PackedFrames, k: int = 4, alpha: float = 0.0,
                 mode: str = "auto", variant: int = 0, r_max: int = DEFAULT_R_MAX):
    """All B*T query frames of `source` [B,D,T] against ONE packed library - or, when `lib` holds
    B libraries (pack_libraries), batch item b against library b.  Returns (out [B,T,D] float32
    contiguous, top_idx [B,T,k] int64 relative to the item's own library, top_score [B,T,k])."""
    out, idx, score = run_match(source, lib, k, alpha, mode, variant, r_max)
    if lib.items > 1:
        idx = idx - (torch.arange(lib.items, device=idx.device, dtype=idx.dtype) * lib.n_item).view(-1, 1, 1)
    return out, idx, score


def match_packed_queries(q: PackedFrames, lib: PackedFrames, k: int = 4, alpha: float = 0.0, mode: str = "auto",
                         variant: int = 0, r_max: int = DEFAULT_R_MAX, batch: int = 1, want_out: bool = True,
                         workspace=None, out=None, top_idx=None, top_score=None):
    """The match for query frames that are ALREADY packed: `q` = pack_frames / pack_rows /
    pack_queries output of the producer (K1 ran as the encoder's epilogue - once, however many libraries the
    utterance is then matched against).  alive_knn_match_packed: no K1 launch, everything else as run_match.
    `q` holds batch * T frames (frame b*T + t); returns (out [batch, T, D] float32, top_idx [batch, T, k] int64 -
    global frame indices when `lib` holds several libraries -, top_score [batch, T, k]), bit-identical to
    run_match on the frames `q` was packed from."""
    global last_info
    c = _cabi.load()
    if q.d != lib.d:
        raise RuntimeError(f"feature dims differ: queries {q.d}, library {lib.d}")
    if not isinstance(k, int) or k < 1 or k > lib.n_item:
        raise RuntimeError("selected index k out of range")
    if k > MAX_K:
        raise RuntimeError(f"alive_vc_b200 supports k <= {MAX_K} (got {k})")
    if batch < 1 or q.n % batch != 0:
        raise RuntimeError(f"{q.n} packed query frames do not split into a batch of {batch}")
    if lib.items > 1 and lib.items != batch:
        raise RuntimeError(f"a packed set of {lib.items} libraries needs a query batch of {lib.items} (got {batch})")
    dev = q.device
    if dev != lib.device:
        raise RuntimeError(f"the packed queries are on {dev} but the packed library is on {lib.device}")
    if q.format != lib.format:
        raise RuntimeError("queries and library must be packed in the same 16-bit format (pack_frames(..., fmt=lib.format))")
    rows, D, T = q.n, q.d, q.n // batch
    m = _MODES[mode]
    if m == 0:
        m = 2 if (k > LIST_LEN or lib.n_item < EXACT_BELOW_N or lib.d % 64 != 0) else 1
    off = _layout(rows, lib, k, r_max, m, variant, dev)
    if workspace is None:
        workspace = torch.empty((off[11],), dtype=torch.uint8, device=dev)
    if want_out and out is None:
        out = torch.empty((batch, T, D), dtype=torch.float32, device=dev)
    if top_idx is None:
        top_idx = torch.empty((batch, T, k), dtype=torch.int64, device=dev)
        top_score = torch.empty((batch, T, k), dtype=torch.float32, device=dev)
    with _on(dev):
        q_lo = q.lo.data_ptr() if (q.lo is not None and q.err2 is not None) else None
        rc = c.alive_knn_match_packed(q.raw.data_ptr(), q.norms.data_ptr(), q.packed.data_ptr(), q.err.data_ptr(),
                                      q_lo, q.err2.data_ptr() if q_lo is not None else None, batch, T,
                                      ctypes.byref(lib.handle()), k, float(alpha), r_max, m, _num_sms(dev), variant,
                                      workspace.data_ptr(), workspace.numel(), out.data_ptr() if want_out else None,
                                      top_idx.data_ptr(), top_score.data_ptr(), _stream_ptr(dev))
    _cabi.check(rc, "alive_knn_match_packed")
    launches = (_collect_launches(lib.lo is not None and q_lo is not None) if off[7] > off[6] else 4) if m == 1 else 2
    _count(launches)
    last_info = SearchInfo(mode="screen" if m == 1 else "exact",
                           fb_count=workspace[off[9]:off[9] + 4 * lib.items].view(torch.int32),
                           exact_count=workspace[off[9] + 4 * lib.items:off[9] + 8 * lib.items].view(torch.int32),
                           collect=m == 1 and off[7] > off[6],
                           sel_n=workspace[off[7]:off[7] + 4 * rows].view(torch.int32) if m == 1 else None,
                           launches=launches)
    last_info._workspace = workspace
    last_info._offsets = off
    return (out if want_out else None), top_idx, top_score


class StreamingMatcher:
    """Fixed-shape matcher for the realtime loop (realtime_inference.py:130-191): the library
    is packed once, every buffer is pre-allocated, and the whole pipeline (pack queries ->
    search -> prune -> rescore -> exact -> gather) is captured in ONE CUDA graph that is
    replayed per chunk - no allocation, no host synchronisation, one graph launch.

        sm = StreamingMatcher(pack_library(tgt), T=32)
        out = sm(chunk)          # chunk [B, D, T] float32 CUDA -> [B, D, T] (static buffer, reused)
    """

    def __init__(self, lib: PackedFrames, T: int, k: int = 4, alpha: float = 0.0, batch: int = 1,
                 mode: str = "auto", variant: int = 0, r_max: int = DEFAULT_R_MAX, use_graph: bool = True):
        dev = lib.device
        self.lib, self.k, self.alpha, self.mode, self.variant, self.r_max = lib, k, float(alpha), mode, variant, r_max
        self.src = torch.zeros((batch, lib.d, T), dtype=torch.float32, device=dev)
        self.out = torch.empty((batch, T, lib.d), dtype=torch.float32, device=dev)
        self.top_idx = torch.empty((batch, T, k), dtype=torch.int64, device=dev)
        self.top_score = torch.empty((batch, T, k), dtype=torch.float32, device=dev)
        m = _MODES[mode]
        if m == 0:
            m = 2 if (k > LIST_LEN or lib.n < EXACT_BELOW_N or lib.d % 64 != 0) else 1
        off = _layout(batch * T, lib, k, r_max, m, variant, dev)
        self.workspace = torch.empty((off[11],), dtype=torch.uint8, device=dev)
        self.graph = None
        self._run()                                   # eager warm-up (one-time attribute setup)
        self._launches = last_info.launches           # kernels per replay = kernels of one eager call
        torch.cuda.synchronize(dev)
        if use_graph:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                self._run()
            torch.cuda.current_stream(dev).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run()
            self.graph = g

    def _run(self):
        run_match(self.src, self.lib, self.k, self.alpha, self.mode, self.variant, self.r_max,
                  workspace=self.workspace, out=self.out, top_idx=self.top_idx, top_score=self.top_score)

    def __call__(self, source: torch.Tensor) -> torch.Tensor:
        self.src.copy_(source, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
            _count(self._launches)
        else:
            self._run()
        return self.out.transpose(1, 2)


_SCALAR_NAMES = {torch.float32: "Float", torch.float16: "Half", torch.bfloat16: "BFloat16", torch.float64: "Double"}


def _check_inputs(source: torch.Tensor, reference: torch.Tensor, k: int):
    # shape / k errors first, with the reference's own messages, then the device check
    if source.dim() != 3 or reference.dim() != 3:
        raise RuntimeError("match_features expects source [B, D, T] and reference [B, D, N]")
    B, D, T = source.shape
    if reference.shape[0] != B or reference.shape[1] != D:
        # message of torch.bmm in the reference (common.py:104)
        raise RuntimeError(
            f"Expected size for first two dimensions of batch2 tensor to be: [{B}, {D}] "
            f"but got: [{reference.shape[0]}, {reference.shape[1]}].")
    N = reference.shape[2]
    if not isinstance(k, int) or k < 1 or k > N:
        raise RuntimeError("selected index k out of range")   # torch.topk (common.py:105)
    _require_cuda(source, "source")
    _require_cuda(reference, "reference")
    if source.device != reference.device:
        raise RuntimeError(f"source is on {source.device} but reference is on {reference.device}")
    _check_dtypes(source, reference)


def _check_dtypes(source: torch.Tensor, reference: torch.Tensor):
    """Mixed dtypes only work in the reference under autocast (realtime_inference.py:144: fp16 encoder output
    against an fp32 library); outside it torch.bmm at common.py:104 raises - so does this."""
    if source.dtype != reference.dtype and not torch.is_autocast_enabled("cuda"):
        raise RuntimeError(f"expected scalar type {_SCALAR_NAMES.get(reference.dtype, reference.dtype)} "
                           f"but found {_SCALAR_NAMES.get(source.dtype, source.dtype)}")


def match_indices(source: torch.Tensor, reference: torch.Tensor, k: int = 4, mode: str = "auto",
                  variant: int = 0):
    """The neighbour indices the reference computes at common.py:105 but never returns:
    ([B,T,k] int64, [B,T,k] float32 similarities)."""
    _check_inputs(source, reference, k)
    with torch.no_grad():
        out = _match_impl(source, reference, k, 0.0, mode, variant, want_out=False)
    return out[1], out[2]


def _match_impl(source, reference, k, alpha, mode, variant, want_out=True):
    """source [B, D, T]; reference [B, D, N] (one library per batch item) or a library shared by every item
    ([1, D, N], or a stride-0 expand of it as VoiceLibrary.match builds at voice_library.py:16-19).
    Returns (out [B, T, D] float32, idx [B, T, k] int64, score [B, T, k] float32)."""
    B, D, T = source.shape
    src32 = source if source.dtype == torch.float32 else source.float()
    ref32 = reference if reference.dtype == torch.float32 else reference.float()
    dev = source.device
    if T == 0:
        return (torch.empty((B, 0, D), dtype=torch.float32, device=dev),
                torch.empty((B, 0, k), dtype=torch.int64, device=dev),
                torch.empty((B, 0, k), dtype=torch.float32, device=dev))
    if reference.shape[0] == 1 or ref32.stride(0) == 0:
        lib = cached_pack(reference, ref32[0])
    else:
        # a different library per batch item (train_decoder.py:134-135, BASELINE cfg5): all items are
        # packed back to back and matched in ONE pipeline launch
        lib = cached_pack_many(reference, ref32)
    out, idx, score = run_match(src32, lib, k, alpha, mode, variant, want_out=want_out)
    if lib.items > 1:
        idx = idx - (torch.arange(lib.items, device=idx.device, dtype=idx.dtype) * lib.n_item).view(-1, 1, 1)
    if out is None:
        out = torch.empty((B, 0, D), dtype=torch.float32, device=dev)
    return out, idx, score


class _BlendGrad(torch.autograd.Function):
    """d(out)/d(frames) of common.py:109 for the row-major entry point (lifecycle.match_rows): the matched term
    carries no gradient to the query frames (indices are not differentiable), the blend contributes alpha * g."""

    @staticmethod
    def forward(ctx, source, out_bdt, alpha):
        ctx.alpha = alpha
        return out_bdt.view_as(out_bdt)

    @staticmethod
    def backward(ctx, g):
        return g * ctx.alpha, None, None


def match_features(source: torch.Tensor, reference: torch.Tensor, k: int = 4, alpha: float = 0.0,
                   *, return_indices: bool = False, mode: str = "auto", variant: int = 0):
    """Drop-in for `module.common.match_features` (module/common.py:96-109).

    source [B, D, T], reference [B, D, N] (same B; D = 768 in ALiVE-VC) -> [B, D, T],
    returned exactly like the reference does: a transposed view of a contiguous
    [B, T, D] block (strides (T*D, 1, D)).  Each source frame is replaced by the mean of
    the RAW library frames of its k most cosine-similar neighbours, blended as
    `result*(1-alpha) + source*alpha`.  `reference` never receives a gradient and
    `source` receives `alpha * grad`, as in the reference (`torch.no_grad` at :98).

    The work is one registered custom op (`torch.ops.alive_vc_b200.knn_match`, alive_vc_b200/ops.py: fake
    kernel + autograd formula), so callers may be traced / torch.compile'd; it works under
    `torch.inference_mode()` and autocast like the call at realtime_inference.py:143-165.

    dtypes: the result has `torch.promote_types(source.dtype, reference.dtype)` like the reference's last
    line; mixed dtypes outside autocast raise like its bmm.  Similarities are always evaluated from the
    float32 values of the inputs (under fp16 autocast the reference's own bmm runs in half precision; this
    path returns the float32-exact neighbours of the same inputs).

    Raises RuntimeError("selected index k out of range") when k > N or the library is
    empty, and the bmm size error on a batch mismatch, like the reference.
    `return_indices=True` additionally returns the [B,T,k] int64 indices (extension used
    for parity checks; the reference computes but never returns them).
    """
    from . import ops
    _check_inputs(source, reference, k)
    out_btd, idx, _ = ops.knn_match(source, reference, k, float(alpha), mode, variant, False)
    out = out_btd.transpose(1, 2)
    want = torch.promote_types(source.dtype, reference.dtype)
    if out.dtype != want:
        out = out.to(want)
    if return_indices:
        return out, idx
    return out
