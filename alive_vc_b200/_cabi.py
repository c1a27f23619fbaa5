"""ctypes binding of the alive_knn C ABI (include/alive_knn.h) and its build recipe.

The shared library is built IN-TREE (`alive_vc_b200/libalive_knn.so`) with nvcc for
sm_100a only.  There is no fallback: if the library is missing or a call fails,
an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libalive_knn.so")
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["api.cu", "pack.cu", "search_sm100.cu", "select.cu", "gather.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]

# every symbol include/alive_knn.h declares
MODE_DEFER_FALLBACK = 0x100      # ALIVE_KNN_MODE_DEFER_FALLBACK
FORMAT_BF16, FORMAT_FP16 = 0, 1      # ALIVE_KNN_FORMAT_*

EXPORTS = [
    "alive_knn_last_error", "alive_knn_abi_version", "alive_knn_pack", "alive_knn_plan", "alive_knn_plan_batched",
    "alive_knn_search", "alive_knn_prune", "alive_knn_rescore", "alive_knn_exact_workspace_bytes",
    "alive_knn_exact", "alive_knn_merge", "alive_knn_gather_mean", "alive_knn_gather_rows",
    "alive_knn_mean_blend", "alive_knn_scatter_grad", "alive_knn_match_layout", "alive_knn_match",
    "alive_knn_finish", "alive_knn_gather_mean_peers", "alive_knn_ipc_export", "alive_knn_ipc_open",
    "alive_knn_ipc_close", "alive_knn_merge_records", "alive_knn_merge_gather", "alive_knn_match_packed",
    "alive_knn_graph_launch", "alive_knn_event_wait", "alive_knn_arm_notify", "alive_knn_flag_wait",
    "alive_knn_match_fallback",
]


class Plan(ctypes.Structure):
    """mirror of alive_knn_plan_t"""
    _fields_ = [
        ("t", ctypes.c_int32), ("n", ctypes.c_int64), ("d", ctypes.c_int32),
        ("ctas_per_unit", ctypes.c_int32), ("m_units", ctypes.c_int32), ("n_tiles", ctypes.c_int32),
        ("segments", ctypes.c_int32), ("tiles_per_segment", ctypes.c_int32), ("lists", ctypes.c_int32),
        ("grid", ctypes.c_int32), ("items", ctypes.c_int32), ("format", ctypes.c_int32), ("kernel", ctypes.c_int32),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class Library(ctypes.Structure):
    """mirror of alive_knn_library_t"""
    _fields_ = [
        ("packed", ctypes.c_void_p), ("raw", ctypes.c_void_p), ("norms", ctypes.c_void_p),
        ("stats", ctypes.c_void_p), ("n", ctypes.c_int64), ("d", ctypes.c_int32), ("row_base", ctypes.c_int64),
        ("items", ctypes.c_int32), ("lo", ctypes.c_void_p), ("format", ctypes.c_int32),
    ]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


OBJ_DIR = os.path.join(CSRC, "_obj")
_HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(_ROOT, "include", "alive_knn.h")]


def _obj_path(src: str) -> str:
    return os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(p) > t for p in deps)


def needs_build() -> bool:
    return _stale(LIB_PATH, [os.path.join(CSRC, s) for s in SOURCES] + _HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into libalive_knn.so for sm_100a (nvcc cross-compiles without a GPU).
    Every source is compiled to its own object (in parallel, only when it or a header changed), then linked."""
    if not force and not needs_build():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor

    os.makedirs(OBJ_DIR, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        obj = _obj_path(src)
        path = os.path.join(CSRC, src)
        if not force and not _stale(obj, [path] + _HEADERS):
            return ""
        cmd = [nvcc_path(), *compile_flags, "-I", os.path.join(_ROOT, "include"), "-c", "-o", obj, path]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        return res.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        logs = list(pool.map(compile_one, SOURCES))
    tmp = LIB_PATH + ".tmp"
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp, *[_obj_path(s) for s in SOURCES]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB_PATH)
    if verbose:
        print("\n".join(logs))
    return LIB_PATH


_lib = None
_lock = threading.Lock()

_vp = ctypes.c_void_p
_i32 = ctypes.c_int32
_i64 = ctypes.c_int64
_f32 = ctypes.c_float


def _declare(lib):
    lib.alive_knn_last_error.restype = ctypes.c_char_p
    lib.alive_knn_last_error.argtypes = []
    lib.alive_knn_abi_version.restype = ctypes.c_int
    lib.alive_knn_abi_version.argtypes = []
    lib.alive_knn_pack.restype = ctypes.c_int
    lib.alive_knn_pack.argtypes = [_vp, _i64, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]
    lib.alive_knn_plan.restype = ctypes.c_int
    lib.alive_knn_plan.argtypes = [_i32, _i64, _i32, _i32, _i32, ctypes.POINTER(Plan)]
    lib.alive_knn_plan_batched.restype = ctypes.c_int
    lib.alive_knn_plan_batched.argtypes = [_i32, _i32, _i64, _i32, _i32, _i32, ctypes.POINTER(Plan)]
    lib.alive_knn_search.restype = ctypes.c_int
    lib.alive_knn_search.argtypes = [_vp, _vp, ctypes.POINTER(Plan), _vp, _vp, _vp]
    lib.alive_knn_prune.restype = ctypes.c_int
    lib.alive_knn_prune.argtypes = [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]
    lib.alive_knn_rescore.restype = ctypes.c_int
    lib.alive_knn_rescore.argtypes = [_vp, _vp, _i32, _vp, _vp, _i32, _vp, _vp, _i32, _i32, _i64, _vp, _vp, _vp]
    lib.alive_knn_exact_workspace_bytes.restype = ctypes.c_size_t
    lib.alive_knn_exact_workspace_bytes.argtypes = [_i32, _i64, _i32, _i32]
    lib.alive_knn_exact.restype = ctypes.c_int
    lib.alive_knn_exact.argtypes = [_vp, _vp, _i32, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _f32,
                                    _vp, _i32, _vp]
    lib.alive_knn_finish.restype = ctypes.c_int
    lib.alive_knn_finish.argtypes = [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i64,
                                     _f32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]
    lib.alive_knn_merge.restype = ctypes.c_int
    lib.alive_knn_merge.argtypes = [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]
    lib.alive_knn_gather_mean.restype = ctypes.c_int
    lib.alive_knn_gather_mean.argtypes = [_vp, _i64, _i32, _vp, _i32, _i32, _vp, _vp, _f32, _vp, _vp]
    lib.alive_knn_gather_rows.restype = ctypes.c_int
    lib.alive_knn_gather_rows.argtypes = [_vp, _i64, _i32, _i64, _vp, _i32, _i32, _vp, _vp]
    lib.alive_knn_mean_blend.restype = ctypes.c_int
    lib.alive_knn_mean_blend.argtypes = [_vp, _i32, _i32, _i32, _vp, _f32, _vp, _vp]
    lib.alive_knn_scatter_grad.restype = ctypes.c_int
    lib.alive_knn_scatter_grad.argtypes = [_vp, _vp, _i32, _i32, _i32, _f32, _vp, _i64, _vp]
    lib.alive_knn_gather_mean_peers.restype = ctypes.c_int
    lib.alive_knn_gather_mean_peers.argtypes = [_vp, _vp, _i32, _i32, _vp, _i32, _i32, _vp, _vp, _f32, _vp, _vp]
    lib.alive_knn_merge_records.restype = ctypes.c_int
    lib.alive_knn_merge_records.argtypes = [_vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp]
    lib.alive_knn_merge_gather.restype = ctypes.c_int
    lib.alive_knn_merge_gather.argtypes = [_vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _f32,
                                           _vp, _vp, _vp, _vp]
    lib.alive_knn_ipc_export.restype = ctypes.c_int
    lib.alive_knn_ipc_export.argtypes = [_vp, ctypes.c_char_p, ctypes.POINTER(_i64)]
    lib.alive_knn_ipc_open.restype = ctypes.c_int
    lib.alive_knn_ipc_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(_vp)]
    lib.alive_knn_ipc_close.restype = ctypes.c_int
    lib.alive_knn_ipc_close.argtypes = [_vp]
    lib.alive_knn_match_layout.restype = ctypes.c_int
    lib.alive_knn_match_layout.argtypes = [_i32, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, ctypes.POINTER(_i64)]
    lib.alive_knn_graph_launch.restype = ctypes.c_int
    lib.alive_knn_graph_launch.argtypes = [_vp, _vp, _vp]
    lib.alive_knn_event_wait.restype = ctypes.c_int
    lib.alive_knn_event_wait.argtypes = [_vp]
    lib.alive_knn_arm_notify.restype = ctypes.c_int
    lib.alive_knn_arm_notify.argtypes = [_vp, _vp]
    lib.alive_knn_flag_wait.restype = ctypes.c_int
    lib.alive_knn_flag_wait.argtypes = [_vp, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64), _i64]
    lib.alive_knn_match_packed.restype = ctypes.c_int
    lib.alive_knn_match_packed.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, ctypes.POINTER(Library), _i32, _f32,
                                           _i32, _i32, _i32, _i32, _vp, ctypes.c_size_t, _vp, _vp, _vp, _vp]
    lib.alive_knn_match_fallback.restype = ctypes.c_int
    lib.alive_knn_match_fallback.argtypes = list(lib.alive_knn_match_packed.argtypes)
    lib.alive_knn_match.restype = ctypes.c_int
    lib.alive_knn_match.argtypes = [_vp, _i32, _i32, _i64, _i64, _i64, ctypes.POINTER(Library), _i32, _f32, _i32,
                                    _i32, _i32, _i32, _vp, ctypes.c_size_t, _vp, _vp, _vp, _vp, _vp, _vp]


def load():
    """dlopen the in-tree library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(nvcc, sm_100a). alive_vc_b200 has no CPU or PyTorch fallback.")
            # ALIVE_KNN_LIB: an instrumented build of the same sources (tests/gpu_tools/finish_phases.py)
            lib = ctypes.CDLL(os.environ.get("ALIVE_KNN_LIB") or LIB_PATH)
            _declare(lib)
            if lib.alive_knn_abi_version() != 7:
                raise RuntimeError("libalive_knn.so ABI version mismatch")
            _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().alive_knn_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what}: {msg}" if msg else f"{what} failed with code {rc}")
