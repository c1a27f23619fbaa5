"""C-ABI surface checks that need no GPU: the library builds for sm_100a, loads, exports
every symbol include/alive_knn.h declares, and its host-only planner behaves."""
import ctypes
import os
import re

import pytest

from alive_vc_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _cabi.build_library()
    return _cabi.load()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "alive_knn.h")).read()
    declared = sorted(set(re.findall(r"\b(alive_knn_[a-z_0-9]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    assert sorted(_cabi.EXPORTS) == declared
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in alive_knn.h but not exported"
    assert lib.alive_knn_abi_version() == 7


def test_sass_is_blackwell_native():
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None and not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump not available")
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([exe, "-sass", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass, "no tcgen05.mma in the SASS"
    assert "LDTM" in sass, "no tcgen05.ld in the SASS"
    assert "UTMALDG" in sass, "no TMA tensor loads in the SASS"
    assert "HMMA.16816" not in sass, "legacy mma.sync path present"
    assert "sm_100a" in subprocess.run([exe, "-lelf", _cabi.LIB_PATH], capture_output=True, text=True).stdout


@pytest.mark.parametrize("t,n", [(1, 1), (1, 300), (24, 3512), (450, 3512), (1000, 100000), (32, 200000),
                                 (100000, 1000000), (10000, 10000000), (10000, 1250000), (1000, 500000),
                                 (129, 4097), (257, 255)])
@pytest.mark.parametrize("variant", [1, 2])
def test_plan_covers_the_problem(lib, t, n, variant):
    p = _cabi.Plan()
    assert lib.alive_knn_plan(t, n, 768, 148, variant, ctypes.byref(p)) == 0
    assert p.t == t and p.n == n and p.d == 768 and p.ctas_per_unit == variant
    assert p.m_units * 128 * variant >= t > (p.m_units - 1) * 128 * variant
    assert p.n_tiles * 256 >= n > (p.n_tiles - 1) * 256
    assert p.segments * p.tiles_per_segment >= p.n_tiles          # every tile belongs to a segment
    assert (p.segments - 1) * p.tiles_per_segment < p.n_tiles     # no empty segment
    assert p.lists == 2 * p.segments
    assert 0 < p.grid <= 148 and p.grid % variant == 0
    assert p.grid // variant <= p.m_units * p.segments
    # wave quantisation: idle tile slots stay under 15% once the problem is big enough to fill the GPU
    units = p.m_units * p.segments
    slots = 148 // variant
    if units >= 4 * slots:
        waves = -(-units // slots)
        assert waves * p.tiles_per_segment * slots <= 1.15 * p.m_units * p.n_tiles + slots


@pytest.mark.parametrize("t,n", [(1, 1), (24, 3512), (32, 200000), (32, 128 * 148), (7, 128 * 148 + 1), (32, 10_000_000)])
def test_skinny_plan(lib, t, n):
    """variant 3 / the default for t <= 32: 128-frame tiles dealt round-robin to the CTAs, one list per CTA."""
    for variant in (3, 0):
        p = _cabi.Plan()
        assert lib.alive_knn_plan(t, n, 768, 148, variant, ctypes.byref(p)) == 0
        assert p.kernel == 1 and p.items == 1 and p.m_units == 1 and p.ctas_per_unit == 1
        assert p.n_tiles * 128 >= n > (p.n_tiles - 1) * 128
        assert p.grid == min(148, p.n_tiles) and p.lists == p.grid == p.segments
        assert p.tiles_per_segment == -(-p.n_tiles // p.grid)
    p = _cabi.Plan()
    assert lib.alive_knn_plan(33, n, 768, 148, 0, ctypes.byref(p)) == 0 and p.kernel == 0      # too many queries
    assert lib.alive_knn_plan(33, n, 768, 148, 3, ctypes.byref(p)) != 0
    assert lib.alive_knn_plan_batched(2, t, n, 768, 148, 3, ctypes.byref(p)) != 0               # single item only
    assert lib.alive_knn_plan_batched(2, t, n, 768, 148, 0, ctypes.byref(p)) == 0 and p.kernel == 0


def test_plan_rejects_bad_arguments(lib):
    p = _cabi.Plan()
    assert lib.alive_knn_plan(0, 10, 768, 148, 1, ctypes.byref(p)) != 0
    assert b"t must be" in lib.alive_knn_last_error()
    assert lib.alive_knn_plan(4, 10, 100, 148, 1, ctypes.byref(p)) != 0
    assert b"multiple of 64" in lib.alive_knn_last_error()
    assert lib.alive_knn_plan(4, 10, 768, 148, 4, ctypes.byref(p)) != 0
    assert lib.alive_knn_plan(4, 2 ** 31, 768, 148, 1, ctypes.byref(p)) != 0


def test_argument_validation_without_gpu(lib):
    # NULL pointers are rejected before any CUDA call is made
    assert lib.alive_knn_pack(None, 4, 768, 1, 4, None, None, None, None, None, None, None, 0, None) != 0
    assert b"NULL" in lib.alive_knn_last_error()
    assert lib.alive_knn_exact_workspace_bytes(1000, 100000, 4, 1) > 0


@pytest.mark.parametrize("rows,n,k,mode", [(32, 200_000, 4, 1), (1000, 100_000, 4, 0), (10_000, 10_000_000, 4, 1),
                                           (24, 512, 4, 0), (7, 64, 16, 0), (450, 3512, 4, 2)])
def test_match_workspace_layout_is_consistent(lib, rows, n, k, mode):
    """alive_knn_match_layout is host-only: 14 256-byte aligned offsets (0..11 ascending, the second query plane
    and its error norms - 12, 13 - sit between q_err and the candidate lists), large enough for every buffer the
    pipeline carves out of the single workspace."""
    off = (ctypes.c_int64 * 14)()
    assert lib.alive_knn_match_layout(rows, n, 768, k, 64, mode, 148, 0, 1, off) == 0
    o = list(off)
    assert o[0] == 0 and all(x % 256 == 0 for x in o)
    assert all(o[i] <= o[i + 1] for i in range(11))
    assert o[3] < o[12] < o[13] < o[4] + 1 and o[13] - o[12] >= rows * 768 * 2 and o[4] - o[13] >= rows * 4
    assert o[1] - o[0] >= rows * 768 * 4          # q_raw
    assert o[2] - o[1] >= rows * 4                # q_norm
    assert o[3] - o[2] >= rows * 768 * 2          # q_packed
    assert o[4] - o[3] >= rows * 4                # q_err
    screen = mode == 1 or (mode == 0 and k <= 8)
    if screen:
        p = _cabi.Plan()
        assert lib.alive_knn_plan(rows, n, 768, 148, 0, ctypes.byref(p)) == 0
        assert o[5] - o[4] >= rows * p.lists * 8 * 4 and o[6] - o[5] >= rows * p.lists * 8 * 4
    assert o[11] - o[10] >= lib.alive_knn_exact_workspace_bytes(rows, n, k, 1)


def test_match_rejects_bad_arguments_before_any_launch(lib):
    lb = _cabi.Library(1, 1, 1, 1, 3, 768, 0, 1)     # n = 3 frames, one item
    # k > n: the reference's torch.topk message (common.py:105)
    rc = lib.alive_knn_match(1, 1, 4, 768 * 4, 1, 4, ctypes.byref(lb), 4, 0.0, 64, 0, 148, 0, 256, 1 << 20, None, 1, 1,
                             None, None, None)
    assert rc != 0 and b"selected index k out of range" in lib.alive_knn_last_error()
    rc = lib.alive_knn_match(None, 1, 4, 768 * 4, 1, 4, ctypes.byref(lb), 1, 0.0, 64, 0, 148, 0, 256, 1 << 20, None, 1, 1,
                             None, None, None)
    assert rc != 0 and b"NULL" in lib.alive_knn_last_error()


def test_batched_plan_and_layout(lib):
    """items > 1 (BASELINE cfg5): one plan / one workspace for all (query batch, library) pairs"""
    p = _cabi.Plan()
    assert lib.alive_knn_plan_batched(64, 1000, 500_000, 768, 148, 0, ctypes.byref(p)) == 0
    assert p.items == 64 and p.t == 1000 and p.n == 500_000 and p.ctas_per_unit == 2
    assert p.m_units == 4 and p.segments * p.tiles_per_segment >= p.n_tiles and p.lists == 2 * p.segments
    units = 64 * p.m_units * p.segments
    waves = -(-units // 74)
    assert waves * p.tiles_per_segment * 74 <= 1.05 * 64 * p.m_units * p.n_tiles     # < 5 % idle tile slots
    off = (ctypes.c_int64 * 14)()
    assert lib.alive_knn_match_layout(64 * 1000, 500_000, 768, 4, 64, 1, 148, 0, 64, off) == 0
    o = list(off)
    assert o[10] - o[9] >= 64 * 4                          # one uncertified-query counter per item
    assert lib.alive_knn_match_layout(1001, 500_000, 768, 4, 64, 1, 148, 0, 64, off) != 0   # rows % items
    assert lib.alive_knn_plan_batched(0, 10, 10, 768, 148, 0, ctypes.byref(p)) != 0
