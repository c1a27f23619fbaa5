"""ctypes loader of the C restatement (oracle/knn_oracle.c).  TEST INFRASTRUCTURE ONLY - see knn_oracle.py.

`build()` compiles it with gcc into oracle/_build/ (git-ignored); `match_features_c` has the signature of
`knn_oracle.match_features_np`.  Used by tests/test_oracle.py to cross-check the numpy oracle against the
golden vectors, and available to the GPU parity tests for sizes the numpy oracle is slow on.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "knn_oracle.c")
OUT_DIR = os.path.join(_HERE, "_build")
LIB = os.path.join(OUT_DIR, "libknn_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = LIB + ".tmp"
    cmd = ["gcc", "-O3", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", "-o", tmp, SRC, "-lm"]
    try:
        res = subprocess.run(cmd, capture_output=True, text=True)
        err = res.stdout + res.stderr if res.returncode != 0 else None
    except OSError as e:                       # no gcc on this box
        err = str(e)
    if err is None:
        os.replace(tmp, LIB)
        return LIB
    if os.path.exists(LIB):                    # a prebuilt library travelled with the snapshot: use it
        return LIB
    raise RuntimeError("gcc failed:\n" + " ".join(cmd) + "\n" + err)


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        lib.alive_oracle_match2.restype = ctypes.c_int
        lib.alive_oracle_match2.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int32] * 6 + [
            ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]
        _lib = lib
    return _lib


def match_features_c(source, reference, k: int = 4, alpha: float = 0.0, return_indices: bool = False,
                     rows: bool = False):
    """module/common.py:96-109 through the C restatement: source [B,D,T], reference [B,D,N] (or [1,D,N], the
    shared library of VoiceLibrary.match) -> out [B,D,T] float32 (+ idx [B,T,k] int64, val [B,T,k]).
    rows=True: the same frames handed over frame-major - source [B,T,D], reference [B,N,D] (or [1,N,D]),
    out [B,T,D] (what the GPU side keeps; spares the large parity tests two multi-GB transposes)."""
    src = np.ascontiguousarray(source, dtype=np.float32)
    ref = np.ascontiguousarray(reference, dtype=np.float32)
    if rows:
        B, T, D = src.shape
        RB, N, D2 = ref.shape
    else:
        B, D, T = src.shape
        RB, D2, N = ref.shape
    if D2 != D or (RB != B and RB != 1):
        raise RuntimeError("batch1 and batch2 must have same batch size")
    out = np.empty(src.shape, dtype=np.float32)
    idx = np.empty((B, T, k), dtype=np.int64)
    val = np.empty((B, T, k), dtype=np.float32)
    rc = _load().alive_oracle_match2(src.ctypes.data, ref.ctypes.data, B, RB, D, T, N, k, float(alpha),
                                     out.ctypes.data, idx.ctypes.data, val.ctypes.data, 1 if rows else 0)
    if rc == -1:
        raise RuntimeError("selected index k out of range")
    if rc == -2:
        raise RuntimeError("batch1 and batch2 must have same batch size")
    if rc != 0:
        raise RuntimeError(f"alive_oracle_match failed with code {rc}")
    if return_indices:
        return out, idx, val
    return out
