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
    assert lib.alive_knn_abi_version() == 1


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


def test_plan_rejects_bad_arguments(lib):
    p = _cabi.Plan()
    assert lib.alive_knn_plan(0, 10, 768, 148, 1, ctypes.byref(p)) != 0
    assert b"t must be" in lib.alive_knn_last_error()
    assert lib.alive_knn_plan(4, 10, 100, 148, 1, ctypes.byref(p)) != 0
    assert b"multiple of 64" in lib.alive_knn_last_error()
    assert lib.alive_knn_plan(4, 10, 768, 148, 3, ctypes.byref(p)) != 0
    assert lib.alive_knn_plan(4, 2 ** 31, 768, 148, 1, ctypes.byref(p)) != 0


def test_argument_validation_without_gpu(lib):
    # NULL pointers are rejected before any CUDA call is made
    assert lib.alive_knn_pack(None, 4, 768, 1, 4, None, None, None, None, None, None) != 0
    assert b"NULL" in lib.alive_knn_last_error()
    assert lib.alive_knn_exact_workspace_bytes(1000, 100000, 4) > 0
