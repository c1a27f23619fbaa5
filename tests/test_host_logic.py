"""Host-side behaviour that must hold without a GPU: the reference's error behaviour, the
loud failure when no CUDA device backs the call (there is no CPU fallback), checkpoint
compatibility of VoiceLibrary, shard arithmetic."""
import os

import pytest
import torch

import alive_vc_b200 as A
from alive_vc_b200.sharded import shard_bounds
from oracle.gen_golden import GOLDEN_DIR

ERRS = dict(l.rstrip("\n").split("\t") for l in open(os.path.join(GOLDEN_DIR, "errors.txt")))


def test_k_larger_than_library_raises_like_reference():
    with pytest.raises(RuntimeError) as e:
        A.match_features(torch.randn(1, 768, 3), torch.randn(1, 768, 2), k=4)
    assert str(e.value) == ERRS["k_gt_n"]


def test_empty_library_raises_like_reference():
    with pytest.raises(RuntimeError) as e:
        A.match_features(torch.randn(1, 768, 3), torch.zeros(1, 768, 0), k=4)
    assert str(e.value) == ERRS["empty_library"]


def test_batch_mismatch_raises_like_reference():
    with pytest.raises(RuntimeError) as e:
        A.match_features(torch.randn(2, 768, 3), torch.randn(1, 768, 20), k=4)
    assert str(e.value) == ERRS["batch_mismatch"]


def test_cpu_tensors_fail_loudly_no_fallback():
    with pytest.raises(RuntimeError, match="no CPU path"):
        A.match_features(torch.randn(1, 768, 3), torch.randn(1, 768, 20), k=4)
    vl = A.VoiceLibrary()
    with pytest.raises(RuntimeError, match="no CPU path"):
        vl(torch.randn(1, 768, 5))
    with pytest.raises(RuntimeError, match="no CPU path"):
        A.pack_library(torch.randn(1, 768, 20))


def test_voice_library_checkpoint_layout():
    vl = A.VoiceLibrary()
    sd = vl.state_dict()
    assert list(sd.keys()) == ["tokens"] and tuple(sd["tokens"].shape) == (1, 768, 512)
    assert sd["tokens"].dtype == torch.float32 and vl.hubert_dim == 768
    # a reference-format checkpoint ({"tokens": [1,768,N]}) loads into a same-sized module
    other = A.VoiceLibrary(num_tokens=100)
    other.load_state_dict({"tokens": torch.zeros(1, 768, 100)})
    with pytest.raises(RuntimeError):        # size mismatch, like the reference module
        other.load_state_dict({"tokens": torch.zeros(1, 768, 512)})
    with pytest.raises(RuntimeError, match="selected index k out of range"):
        A.VoiceLibrary(num_tokens=2).match(torch.randn(1, 768, 5), k=4)


@pytest.mark.parametrize("n,world", [(10, 1), (10, 2), (10, 3), (7, 8), (10_000_000, 8), (0, 4)])
def test_shard_bounds_partition(n, world):
    prev = 0
    for r in range(world):
        lo, hi = shard_bounds(n, world, r)
        assert lo == prev and hi >= lo and hi - lo in (n // world, n // world + 1)
        prev = hi
    assert prev == n


def test_lifecycle_host_logic_without_a_gpu():
    """LibraryBuilder has no CPU path; match_windows of nothing is nothing."""
    from alive_vc_b200.lifecycle import LibraryBuilder, match_windows
    with pytest.raises(RuntimeError, match="CUDA"):
        LibraryBuilder(device="cpu")
    assert match_windows([], None) == []


def test_match_rows_host_checks_without_a_gpu():
    """Row-major entry points: shape errors first, then the loud no-CPU-path failure."""
    from alive_vc_b200.lifecycle import match_rows, pack_rows
    with pytest.raises(RuntimeError, match=r"\[T, D\] or \[B, T, D\]"):
        match_rows(torch.randn(768), torch.randn(20, 768))
    with pytest.raises(RuntimeError, match=r"\[N, D\]"):
        pack_rows(torch.randn(2, 20, 768))
    with pytest.raises(RuntimeError, match="no CPU path"):
        pack_rows(torch.randn(20, 768))
    with pytest.raises(RuntimeError, match="no CPU path"):
        match_rows(torch.randn(5, 768), torch.randn(20, 768))


def test_custom_ops_are_registered_and_traceable_without_a_gpu():
    """north_star: "a thin C-ABI torch custom op".  The ops exist in the dispatcher with a fake (meta) kernel, so
    shape propagation / tracing needs neither a GPU nor the CUDA library."""
    from alive_vc_b200 import ops   # noqa: F401  (registers the ops)
    op = torch.ops.alive_vc_b200.knn_match.default
    assert "reference_grad" in str(op._schema)
    src = torch.empty(3, 768, 11, device="meta", dtype=torch.float16)
    ref = torch.empty(1, 768, 500, device="meta")
    out, idx, score = op(src, ref, 4, 0.0, "auto", 0, False)
    assert tuple(out.shape) == (3, 11, 768) and out.dtype == torch.float32
    assert tuple(idx.shape) == (3, 11, 4) and idx.dtype == torch.int64 and tuple(score.shape) == (3, 11, 4)
    g = torch.ops.alive_vc_b200.knn_scatter_grad.default(torch.empty(33, 768, device="meta"),
                                                         torch.empty(33, 4, device="meta", dtype=torch.int64), 500, 0.25)
    assert tuple(g.shape) == (500, 768)
    # no CPU kernel: the op itself refuses CPU tensors (the public wrappers raise their own message first)
    with pytest.raises((NotImplementedError, RuntimeError)):
        op(torch.randn(1, 768, 3), torch.randn(1, 768, 20), 4, 0.0, "auto", 0, False)


def test_record_layout_of_the_sharded_exchange():
    from alive_vc_b200.sharded import record_bytes
    assert record_bytes(10_000, 4) == 480_000 and record_bytes(5, 3) == 192 and record_bytes(1, 1) == 16
    for t, k in ((7, 3), (1250, 4), (33, 8)):
        assert record_bytes(t, k) >= t * k * 12 and record_bytes(t, k) % 16 == 0
