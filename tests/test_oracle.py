"""The oracle (oracle/knn_oracle.py) against golden vectors produced by the
UNMODIFIED reference (oracle/gen_golden.py, run where /root/reference exists).
CPU only; no reference checkout needed at test time."""
import os

import numpy as np
import pytest

from oracle import knn_oracle as O
from oracle.gen_golden import CASES, GOLDEN_DIR, make_case_inputs


def _load(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(name):
    spec = CASES[name]
    g = _load(name)
    src, ref = make_case_inputs(spec)
    if spec["kind"] == "vl":
        out, idx, val = O.voice_library_match_np(ref, src, spec["k"], spec["alpha"], True)
        ref_b = np.broadcast_to(ref, (src.shape[0],) + ref.shape[1:])
    else:
        out, idx, val = O.match_features_np(src, ref, spec["k"], spec["alpha"], True)
        ref_b = ref
    scores = O.cosine_scores_np(src, ref_b)
    ok, n_exact, n_tie, bad = O.indices_match_mod_ties(idx, g["indices"].astype(np.int64), scores, 1e-6)
    assert ok, bad
    # features: 1e-5 relative (north_star); rows whose index sets are identical must be bit-exact
    same = np.all(idx == g["indices"], axis=2)                      # [B,T]
    o = np.swapaxes(out, 1, 2)
    r = np.swapaxes(g["out"], 1, 2)
    assert np.array_equal(o[same], r[same]), "gather-mean/blend model is not bit-exact"
    if spec["kind"] not in ("mf_dupes",):
        np.testing.assert_allclose(out, g["out"], rtol=1e-5, atol=1e-6)
    # reference returns a [B,D,T] view of a contiguous [B,T,D] block
    B, D, T = g["out"].shape
    if T > 1:
        assert tuple(g["out_strides"]) == (T * D, 1, D)


@pytest.mark.parametrize("name", ["vl_default", "vl_alpha", "vl_big"])
def test_oracle_gradients_match_reference(name):
    spec = CASES[name]
    g = _load(name)
    src, tokens = make_case_inputs(spec)
    grad_out = np.random.default_rng(spec["seed"] + 1000).standard_normal(g["out"].shape, dtype=np.float32)
    gt, gs = O.voice_library_grads_np(tokens, src, grad_out, spec["k"], spec["alpha"])
    np.testing.assert_allclose(gt, g["grad_tokens"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(gs, g["grad_source"], rtol=1e-6, atol=1e-7)


def test_oracle_error_behaviour():
    errs = dict(l.rstrip("\n").split("\t") for l in open(os.path.join(GOLDEN_DIR, "errors.txt")))
    assert "selected index k out of range" in errs["k_gt_n"]
    with pytest.raises(RuntimeError, match="selected index k out of range"):
        O.match_features_np(np.zeros((1, 768, 3), np.float32), np.ones((1, 768, 2), np.float32), 4)
    with pytest.raises(RuntimeError, match="selected index k out of range"):
        O.match_features_np(np.zeros((1, 768, 3), np.float32), np.ones((1, 768, 0), np.float32), 4)
    with pytest.raises(RuntimeError):
        O.match_features_np(np.zeros((2, 768, 3), np.float32), np.ones((1, 768, 20), np.float32), 4)


def test_oracle_nan_ranks_first():
    # a zero library row has norm 0 -> NaN similarity -> torch.topk ranks it first
    rng = np.random.default_rng(5)
    src = rng.standard_normal((1, 768, 6), dtype=np.float32)
    ref = rng.standard_normal((1, 768, 40), dtype=np.float32)
    ref[:, :, 17] = 0
    _, idx, val = O.match_features_np(src, ref, 4, 0.0, True)
    assert (idx[0, :, 0] == 17).all() and np.isnan(val[0, :, 0]).all()


def test_torch_port_equals_numpy_oracle():
    torch = pytest.importorskip("torch")
    spec = CASES["mf_mid"]
    src, ref = make_case_inputs(spec)
    out_t, idx_t = O.match_features_torch(torch.from_numpy(src), torch.from_numpy(ref), 4, 0.0)
    g = _load("mf_mid")
    assert np.array_equal(idx_t.numpy(), g["indices"])
    assert np.array_equal(out_t.numpy(), g["out"])


# ---- the C restatement (oracle/knn_oracle.c): an independent second oracle ----------------------
@pytest.mark.parametrize("name", sorted(CASES))
def test_c_oracle_matches_reference_golden(name):
    """Same pin as the numpy oracle: every golden vector of the unmodified reference.  Indices modulo ties
    within 1e-6 (the C oracle rounds a double accumulation once, the reference runs a float32 sgemm);
    features bit-exact on every row whose index set agrees."""
    from oracle import c_oracle as C
    spec = CASES[name]
    g = _load(name)
    src, ref = make_case_inputs(spec)
    out, idx, val = C.match_features_c(src, ref, spec["k"], spec["alpha"], True)      # vl cases: ref is [1,D,N]
    ref_b = np.broadcast_to(ref, (src.shape[0],) + ref.shape[1:])
    scores = O.cosine_scores_np(src, ref_b)
    ok, n_exact, n_tie, bad = O.indices_match_mod_ties(idx, g["indices"].astype(np.int64), scores, 1e-6)
    assert ok, bad
    same = np.all(idx == g["indices"], axis=2)
    assert same.mean() > 0.9 or spec["kind"] == "mf_dupes"
    o = np.swapaxes(out, 1, 2)
    r = np.swapaxes(g["out"], 1, 2)
    assert np.array_equal(o[same], r[same], equal_nan=True), "C gather-mean/blend is not bit-exact"
    # its own similarities agree with the float32 restatement within the tie tolerance
    picked = np.take_along_axis(scores, idx, axis=2)
    fin = np.isfinite(val) & np.isfinite(picked)
    assert np.all(np.abs(val[fin] - picked[fin]) <= 1e-6)


def test_c_oracle_equals_numpy_oracle_on_random_cases_and_errors():
    from oracle import c_oracle as C
    rng = np.random.default_rng(12)
    for (B, T, N, k, alpha) in [(1, 50, 3000, 4, 0.0), (2, 17, 600, 8, 0.3), (1, 5, 40, 1, 1.0), (3, 9, 33, 16, 0.0)]:
        src = rng.standard_normal((B, 768, T), dtype=np.float32)
        ref = rng.standard_normal((B, 768, N), dtype=np.float32)
        o_n, i_n, _ = O.match_features_np(src, ref, k, alpha, True)
        o_c, i_c, _ = C.match_features_c(src, ref, k, alpha, True)
        ok, _, _, bad = O.indices_match_mod_ties(i_c, i_n, O.cosine_scores_np(src, ref), 1e-6)
        assert ok, bad
        same = np.all(i_c == i_n, axis=2)
        assert np.array_equal(np.swapaxes(o_c, 1, 2)[same], np.swapaxes(o_n, 1, 2)[same])
    # NaN ranks first, ties to the lowest index
    src = rng.standard_normal((1, 768, 6), dtype=np.float32)
    ref = rng.standard_normal((1, 768, 40), dtype=np.float32)
    ref[:, :, 17] = 0
    ref[:, :, 30] = ref[:, :, 3]
    _, idx, val = C.match_features_c(src, ref, 4, 0.0, True)
    assert (idx[0, :, 0] == 17).all() and np.isnan(val[0, :, 0]).all()
    assert not (idx == 30).any() or ((idx == 3).any(axis=2) >= (idx == 30).any(axis=2)).all()
    with pytest.raises(RuntimeError, match="selected index k out of range"):
        C.match_features_c(np.zeros((1, 768, 3), np.float32), np.ones((1, 768, 2), np.float32), 4)
    with pytest.raises(RuntimeError):
        C.match_features_c(np.zeros((2, 768, 3), np.float32), np.ones((3, 768, 20), np.float32), 4)
