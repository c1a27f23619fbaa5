"""Parity tests proper: the CUDA path (through the C ABI / the drop-in Python API) against the
oracle and against golden vectors produced by the unmodified reference.

Bars (north_star): neighbour indices bit-exact except where fp32 similarities tie within
1e-6; features within 1e-5 relative - and bit-exact wherever the index rows agree, because
the gather/mean/blend arithmetic is modelled exactly.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import alive_vc_b200 as A                                   # noqa: E402
from alive_vc_b200 import _cabi, matching as M               # noqa: E402
from oracle import knn_oracle as O                           # noqa: E402
from oracle.gen_golden import CASES, GOLDEN_DIR, make_case_inputs   # noqa: E402

TIE_TOL = 1e-6       # north_star: indices may differ only where fp32 similarities tie within 1e-6
FEAT_RTOL = 1e-5     # north_star: features within 1e-5 relative


def _cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _assert_parity(out, idx, src, ref_b, k, alpha, want_idx=None, want_out=None):
    scores = O.cosine_scores_np(src, ref_b)
    if want_idx is None:
        want_out, want_idx, _ = O.match_features_np(src, ref_b, k, alpha, True)
    idx_np = idx.cpu().numpy()
    ok, n_exact, n_tie, bad = O.indices_match_mod_ties(idx_np, want_idx.astype(np.int64), scores, TIE_TOL)
    assert ok, f"index parity broken: {bad}"
    o = out.detach().cpu().numpy()
    same = (idx_np == want_idx).all(axis=2)
    assert np.array_equal(np.swapaxes(o, 1, 2)[same], np.swapaxes(want_out, 1, 2)[same]), \
        "features not bit-exact on rows with identical indices"
    # rows excused by a similarity tie: the features must still be the exact mean of OUR indices
    if (~same).any():
        mine = np.swapaxes(O.gather_mean_np(ref_b, idx_np), 1, 2)
        mine = ((mine * np.float32(1 - alpha)).astype(np.float32) + (src * np.float32(alpha)).astype(np.float32))
        np.testing.assert_allclose(o, mine.astype(np.float32), rtol=FEAT_RTOL, atol=1e-6)
    else:
        np.testing.assert_allclose(o, want_out, rtol=FEAT_RTOL, atol=1e-6)
    return n_exact, n_tie


@pytest.fixture
def screen_format(request, monkeypatch):
    """the 16-bit format of the packed planes for this test (matching.SCREEN_FORMAT)"""
    monkeypatch.setattr(M, "SCREEN_FORMAT", request.param)
    M.clear_pack_cache()
    yield request.param
    M.clear_pack_cache()


@pytest.mark.parametrize("screen_format", ["fp16", "bf16"], indirect=True)
@pytest.mark.parametrize("name", sorted(CASES))
def test_golden_vectors_from_reference(name, screen_format):
    """Replays every fixture generated from the unmodified reference (tests/golden), with the planes stored as fp16
    (default) and as bf16."""
    spec = CASES[name]
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    src, ref = make_case_inputs(spec)
    s = _cuda(src)
    if spec["kind"] == "vl":
        vl = A.VoiceLibrary(num_tokens=spec["N"]).cuda()
        with torch.no_grad():
            vl.tokens.copy_(_cuda(ref))
        s.requires_grad_(True)
        out, idx = vl.match(s, k=spec["k"], alpha=spec["alpha"], return_indices=True)
        gout = _cuda(np.random.default_rng(spec["seed"] + 1000).standard_normal(g["out"].shape, dtype=np.float32))
        out.backward(gout)
        np.testing.assert_allclose(vl.tokens.grad.cpu().numpy(), g["grad_tokens"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(s.grad.cpu().numpy(), g["grad_source"], rtol=1e-6, atol=1e-7)
        ref_b = np.broadcast_to(ref, (src.shape[0],) + ref.shape[1:])
    else:
        r = _cuda(ref)
        if spec["kind"] == "mf_strided":           # non-contiguous library view, realtime_inference.py:88
            big = torch.zeros((ref.shape[0], ref.shape[1], ref.shape[2] * 4), device="cuda")
            big[:, :, ::4] = r
            r = big[:, :, ::4]
            assert not r.is_contiguous()
        out, idx = A.match_features(s, r, spec["k"], spec["alpha"], return_indices=True)
        ref_b = ref
    _assert_parity(out, idx, src, ref_b, spec["k"], spec["alpha"], g["indices"], g["out"])
    B, D, T = g["out"].shape
    assert tuple(out.shape) == (B, D, T)
    if T > 1:
        assert tuple(out.stride()) == tuple(g["out_strides"])      # transposed view of [B,T,D]


@pytest.mark.parametrize("screen_format", ["fp16", "bf16"], indirect=True)
@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("B,T,N,k,alpha", [
    (1, 50, 3000, 4, 0.0), (1, 200, 20000, 4, 0.0), (2, 33, 1517, 4, 0.25), (1, 96, 50000, 8, 0.0),
    (1, 1, 2049, 1, 0.0), (1, 129, 4097, 4, 0.0), (1, 300, 257, 4, 0.0), (3, 7, 256, 2, 1.0),
])
def test_screened_path_against_oracle(variant, B, T, N, k, alpha, screen_format):
    """tensor-core screen + certificate + exact rescoring == oracle, both kernel variants"""
    rng = np.random.default_rng(1000 * T + N + k)
    src = rng.standard_normal((B, 768, T), dtype=np.float32)
    ref = rng.standard_normal((B, 768, N), dtype=np.float32)
    out, idx = A.match_features(_cuda(src), _cuda(ref), k, alpha, return_indices=True, mode="screen", variant=variant)
    assert M.last_info.mode == "screen"
    _assert_parity(out, idx, src, ref, k, alpha)


@pytest.mark.parametrize("T,N,k", [(50, 300, 4), (7, 64, 4), (40, 1000, 8), (10, 700, 16), (5, 4, 4), (33, 20000, 4),
                                   (3, 70, 64), (16, 5003, 4), (1, 1001, 1), (9, 302, 32), (13, 40001, 8),
                                   (17, 777, 4)])
def test_exact_scan_against_oracle(T, N, k):
    rng = np.random.default_rng(T * N + k)
    src = rng.standard_normal((1, 768, T), dtype=np.float32)
    ref = rng.standard_normal((1, 768, N), dtype=np.float32)
    out, idx = A.match_features(_cuda(src), _cuda(ref), k, 0.0, return_indices=True, mode="exact")
    _assert_parity(out, idx, src, ref, k, 0.0)


def test_full_size_cfg1_against_oracle():
    """BASELINE configs[0] at full size: T=1000 vs N=100k (the oracle needs ~1 s on CPU)."""
    rng = np.random.default_rng(11)
    src = rng.standard_normal((1, 768, 1000), dtype=np.float32)
    ref = rng.standard_normal((1, 768, 100_000), dtype=np.float32)
    out, idx = A.match_features(_cuda(src), _cuda(ref), 4, 0.0, return_indices=True)
    n_exact, n_tie = _assert_parity(out, idx, src, ref, 4, 0.0)
    assert M.last_info.mode == "screen" and n_exact >= 990
    assert M.last_info.fallback_queries() <= 10


def test_full_size_cfg2_against_oracle():
    """BASELINE configs[1] at full size: streaming chunk T=32 vs N=200k."""
    rng = np.random.default_rng(12)
    src = rng.standard_normal((1, 768, 32), dtype=np.float32)
    ref = rng.standard_normal((1, 768, 200_000), dtype=np.float32)
    out, idx = A.match_features(_cuda(src), _cuda(ref), 4, 0.0, return_indices=True)
    _assert_parity(out, idx, src, ref, 4, 0.0)


@pytest.mark.parametrize("fmt,dtype,err_max,err2_max", [("bf16", torch.bfloat16, 3e-3, 1.2e-5), ("fp16", torch.float16, 4e-4, 2e-6)])
def test_pack_kernel_layout_and_stats(fmt, dtype, err_max, err2_max):
    g = torch.Generator(device="cuda").manual_seed(1)
    # (n, row_major): every kernel of K1 - per-frame, 8-frame, generic 32-frame (n % 4 != 0), the 16-byte
    # cp.async channel-major kernel (ragged and full last CTA) and the register-resident row-major kernel
    for n, row_major in ((1, False), (31, False), (300, False), (4097, False), (8201, False), (8204, False),
                         (12_000 * 32, False), (600, True), (8204, True), (8201, True)):
        x = torch.randn(768, n, device="cuda", generator=g)
        if row_major:
            x = x.t().contiguous().t()
        p = M.pack_frames(x, fmt=fmt)
        assert p.packed.dtype == dtype and p.format == M._FORMATS[fmt]
        assert torch.equal(p.raw, x.t().contiguous())
        nrm = torch.linalg.vector_norm(x.double(), dim=0).float()
        assert torch.allclose(p.norms, nrm, rtol=2e-7, atol=0)
        xn = (x / p.norms[None, :]).t()
        assert torch.equal(p.packed, xn.to(dtype))
        err = (xn.to(dtype).float() - xn).double().norm(dim=1).float()
        assert torch.allclose(p.err, err, rtol=1e-3, atol=1e-8)
        st = p.stats.cpu().numpy().view(np.uint32)
        assert st[1] == 0
        assert abs(np.array([st[0]], np.uint32).view(np.float32)[0] - float(p.err.max())) < 1e-9
        # the second plane: lo = bf16(x^ - hi), err2 = |x^ - hi - lo| (rounded up), stats[2] = its maximum
        assert p.lo is not None and p.err2 is not None
        r1 = xn - xn.to(dtype).float()
        assert torch.equal(p.lo, r1.to(dtype))
        err2 = (r1 - r1.to(dtype).float()).double().norm(dim=1).float()
        assert torch.allclose(p.err2, err2, rtol=1e-3, atol=1e-11) and bool((p.err2 >= err2).all())
        assert abs(np.array([st[2]], np.uint32).view(np.float32)[0] - float(p.err2.max())) < 1e-12
        assert float(p.err2.max()) < err2_max and float(p.err.max()) < err_max
    x = torch.randn(500, 768, device="cuda", generator=g)      # already row-major frames
    assert torch.equal(M.pack_frames(x.t()).raw, x)
    x = torch.randn(768, 40, device="cuda", generator=g)
    x[:, 17] = 0                                               # zero frame -> NaN similarities in the reference
    assert M.pack_frames(x).stats.cpu().numpy().view(np.uint32)[1] == 1


def test_zero_library_frame_ranks_first_like_torch_topk():
    rng = np.random.default_rng(5)
    src = rng.standard_normal((1, 768, 6), dtype=np.float32)
    ref = rng.standard_normal((1, 768, 3000), dtype=np.float32)
    ref[:, :, 17] = 0
    out, idx = A.match_features(_cuda(src), _cuda(ref), 4, 0.0, return_indices=True, mode="screen")
    assert (idx[0, :, 0] == 17).all()
    assert M.last_info.fallback_queries() == 6            # non-finite library rows force the exact scan
    _, idx_o, _ = O.match_features_np(src, ref, 4, 0.0, True)
    assert np.array_equal(idx.cpu().numpy(), idx_o)


def test_duplicate_frames_and_certificate_fallback():
    """exact ties (duplicated frames) cannot be certified by the screen -> exact scan, lowest index wins"""
    rng = np.random.default_rng(15)
    base = rng.standard_normal((1, 768, 700), dtype=np.float32)
    ref = np.concatenate([base] * 4, axis=2)                   # N = 2800, every frame 4 times
    src = base[:, :, :40].copy()
    out, idx = A.match_features(_cuda(src), _cuda(ref), 4, 0.0, return_indices=True, mode="screen")
    idx_np = idx.cpu().numpy()[0]
    want = np.stack([np.arange(40) + 700 * j for j in range(4)], axis=1)
    assert np.array_equal(np.sort(idx_np, axis=1), want)        # the four copies of the query frame itself
    np.testing.assert_allclose(out.cpu().numpy(), src, rtol=1e-6, atol=1e-6)


def test_search_lists_match_bf16_matmul():
    """K2 alone: the screened lists equal a torch matmul of the same bf16 operands, per list."""
    for variant, T, N in [(1, 200, 5000), (2, 200, 5000), (1, 129, 4097), (2, 300, 70000)]:
        g = torch.Generator(device="cuda").manual_seed(T + N)
        q = M.pack_frames(torch.randn(768, T, device="cuda", generator=g))
        lib = M.pack_frames(torch.randn(768, N, device="cuda", generator=g))
        plan = M.make_plan(T, N, 768, q.device, variant)
        cs = torch.full((T, plan.lists, 8), float("nan"), device="cuda")
        ci = torch.full((T, plan.lists, 8), -7, dtype=torch.int32, device="cuda")
        rc = _cabi.load().alive_knn_search(q.packed.data_ptr(), lib.packed.data_ptr(), ctypes.byref(plan),
                                           cs.data_ptr(), ci.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc, "search")
        ref = q.packed.float() @ lib.packed.float().t()
        for seg in range(plan.segments):
            t0, t1 = seg * plan.tiles_per_segment, min((seg + 1) * plan.tiles_per_segment, plan.n_tiles)
            for half in range(2):
                cols = [torch.arange(t * 256 + half * 128, min(t * 256 + half * 128 + 128, N), device="cuda")
                        for t in range(t0, t1) if t * 256 + half * 128 < N]
                lst = seg * 2 + half
                if not cols:
                    assert (ci[:, lst] == -1).all()
                    continue
                cols = torch.cat(cols)
                kk = min(8, cols.numel())
                want_s, _ = ref[:, cols].topk(kk, dim=1)
                assert (cs[:, lst, :kk] - want_s).abs().max().item() < 2e-4
                got_at = ref.gather(1, ci[:, lst, :kk].long())
                assert (got_at - cs[:, lst, :kk]).abs().max().item() < 2e-4
                assert (ci[:, lst, kk:] == -1).all()


@pytest.mark.parametrize("T,N,D", [(32, 5000, 768), (24, 3512, 768), (1, 128, 768), (7, 129, 768), (32, 70001, 768),
                                   (32, 2000, 64), (13, 4000, 1024), (32, 200000, 768)])
def test_skinny_search_lists_match_bf16_matmul(T, N, D):
    """The skinny kernel (t <= 32): list c holds the top-8 of the 128-frame tiles c, c+grid, ... ."""
    g = torch.Generator(device="cuda").manual_seed(T + N)
    q = M.pack_frames(torch.randn(D, T, device="cuda", generator=g))
    lib = M.pack_frames(torch.randn(D, N, device="cuda", generator=g))
    plan = M.make_plan(T, N, D, q.device, 3)
    assert plan.kernel == 1
    cs = torch.full((T, plan.lists, 8), float("nan"), device="cuda")
    ci = torch.full((T, plan.lists, 8), -7, dtype=torch.int32, device="cuda")
    rc = _cabi.load().alive_knn_search(q.packed.data_ptr(), lib.packed.data_ptr(), ctypes.byref(plan),
                                       cs.data_ptr(), ci.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _cabi.check(rc, "search")
    ref = q.packed.float() @ lib.packed.float().t()
    for c in range(plan.grid):
        cols = torch.cat([torch.arange(t * 128, min(t * 128 + 128, N), device="cuda")
                          for t in range(c, plan.n_tiles, plan.grid)])
        kk = min(8, cols.numel())
        want_s, _ = ref[:, cols].topk(kk, dim=1)
        assert (cs[:, c, :kk] - want_s).abs().max().item() < 2e-4
        got_idx = ci[:, c, :kk].long()
        assert torch.isin(got_idx, cols).all()                       # only frames of this CTA's tiles
        assert (ref.gather(1, got_idx) - cs[:, c, :kk]).abs().max().item() < 2e-4
        assert (ci[:, c, kk:] == -1).all()
        assert (cs[:, c, :-1] >= cs[:, c, 1:]).all()                 # descending, -inf padded


def test_pack_cache_invalidation_on_inplace_update():
    """fine_tune.py:170 updates VL.tokens in place every step: the packed copy must follow."""
    vl = A.VoiceLibrary(num_tokens=2000).cuda()            # >= PACK_CACHE_MIN_ELEMENTS: the packed copy is cached
    src = torch.randn(1, 768, 20, device="cuda")
    out1, idx1 = vl.match(src, return_indices=True)
    p1 = vl.packed()
    assert vl.packed() is p1                                   # cached while unchanged
    with torch.no_grad():
        vl.tokens.mul_(-1.0)                                   # in-place: bumps the version counter
    out2, idx2 = vl.match(src, return_indices=True)
    assert vl.packed() is not p1
    ref = vl.tokens.detach().cpu().numpy()
    _assert_parity(out2, idx2, src.cpu().numpy(), ref, 4, 0.0)
    assert not torch.equal(idx1, idx2)


def test_idempotence_and_self_match_property():
    """size-independent properties at a larger size: every library frame's best match is itself
    (similarity 1), and matching twice gives identical bits."""
    g = torch.Generator(device="cuda").manual_seed(9)
    ref = torch.randn(1, 768, 300_000, device="cuda", generator=g)
    src = ref[:, :, 1000:1000 + 2048].contiguous()
    out, idx = A.match_features(src, ref, 4, 0.0, return_indices=True)
    assert torch.equal(idx[0, :, 0], torch.arange(1000, 1000 + 2048, device="cuda"))
    out_b, idx_b = A.match_features(src, ref, 4, 0.0, return_indices=True)
    assert torch.equal(idx, idx_b) and torch.equal(out, out_b)
    # k=1 returns the frames themselves bit-exactly
    out1 = A.match_features(src, ref, 1, 0.0)
    assert torch.equal(out1, src)
    # the two kernel variants agree bit for bit
    o1, i1 = A.match_features(src, ref, 4, 0.0, return_indices=True, variant=1)
    o2, i2 = A.match_features(src, ref, 4, 0.0, return_indices=True, variant=2)
    assert torch.equal(i1, i2) and torch.equal(o1, o2)


def test_dtype_passthrough_and_autograd_of_match_features():
    """dtype rules of the reference [probed on its own code]: same dtype in -> same dtype out; mixed dtypes raise
    in its bmm outside autocast and give promote_types(source, reference) under autocast"""
    src = torch.randn(1, 768, 12, device="cuda", dtype=torch.float16)
    ref = torch.randn(1, 768, 2000, device="cuda")
    with pytest.raises(RuntimeError, match="expected scalar type Float but found Half"):
        A.match_features(src, ref, 4, 0.0)
    out = A.match_features(src, ref.half(), 4, 0.0)
    assert out.dtype == torch.float16 and tuple(out.shape) == (1, 768, 12)
    with torch.autocast("cuda", dtype=torch.float16):
        out = A.match_features(src, ref, 4, 0.0)
    assert out.dtype == torch.float32 and tuple(out.shape) == (1, 768, 12)
    assert torch.equal(out, A.match_features(src.float(), ref, 4, 0.0))
    s = torch.randn(1, 768, 12, device="cuda", requires_grad=True)
    r = torch.randn(1, 768, 2000, device="cuda", requires_grad=True)
    out = A.match_features(s, r, 4, 0.3)
    out.sum().backward()
    assert r.grad is None                                         # common.py:98 no_grad
    torch.testing.assert_close(s.grad, torch.full_like(s, 0.3))


def test_streaming_matcher_graph_replay_equals_functional_api():
    """realtime loop (realtime_inference.py:130-191): graph-replayed fixed-shape matcher"""
    g = torch.Generator(device="cuda").manual_seed(21)
    ref = torch.randn(1, 768, 60_000, device="cuda", generator=g)
    lib = A.pack_library(ref)
    sm = A.StreamingMatcher(lib, T=24, k=4, alpha=0.0)
    assert sm.graph is not None
    for i in range(3):
        chunk = torch.randn(1, 768, 24, device="cuda", generator=g)
        out = sm(chunk).clone()
        want, widx = A.match_features(chunk, ref, 4, 0.0, return_indices=True)
        assert torch.equal(out, want) and torch.equal(sm.top_idx, widx)
        assert tuple(out.shape) == (1, 768, 24)
    sm2 = A.StreamingMatcher(lib, T=24, k=4, alpha=0.5, use_graph=False)
    chunk = torch.randn(1, 768, 24, device="cuda", generator=g)
    assert torch.equal(sm2(chunk), A.match_features(chunk, ref, 4, 0.5))


@pytest.mark.parametrize("T,N", [(24, 3000), (32, 600_000), (300, 70_000)])
def test_streaming_matcher_fallback_inside_the_graph(T, N):
    """The fallback chain of a captured pipeline (collect pass / exhaustive scan): clean chunks and chunks with
    uncertifiable queries (a cluster of 400 near-identical library frames - more survivors than r_max) alternate
    on the SAME graph and equal the eager path; the device-side fallback counter is reset by every replay."""
    g = torch.Generator(device="cuda").manual_seed(33)
    ref = torch.randn(1, 768, N, device="cuda", generator=g)
    centre = torch.randn(768, 1, device="cuda", generator=g)
    ref[0, :, 1000:1400] = centre + 0.01 * torch.randn(768, 400, device="cuda", generator=g)
    lib = A.pack_library(ref)
    sm = A.StreamingMatcher(lib, T=T, k=4, alpha=0.0)
    assert sm.graph is not None
    info = M.last_info                                         # views of the matcher's own workspace counters
    clean = torch.randn(1, 768, T, device="cuda", generator=g)
    dirty = clean.clone()
    dirty[0, :, :10] = centre + 0.01 * torch.randn(768, 10, device="cuda", generator=g)   # 400 frames inside the band
    for i, (chunk, expect_fb) in enumerate([(clean, False), (dirty, True), (clean, False), (dirty, True), (dirty, True),
                                            (clean, False)]):
        out = sm(chunk).clone()
        idx = sm.top_idx.clone()
        fb = info.fallback_queries()
        assert (fb > 0) == expect_fb, (i, fb)
        want, widx = A.match_features(chunk, ref, 4, 0.0, return_indices=True)
        assert torch.equal(idx, widx), i
        assert torch.equal(out, want), i


def test_one_call_pipeline_equals_step_by_step_kernels():
    """alive_knn_match == pack + search + prune + rescore + exact + gather called one by one"""
    g = torch.Generator(device="cuda").manual_seed(22)
    src = torch.randn(2, 768, 77, device="cuda", generator=g)
    ref = torch.randn(1, 768, 30_000, device="cuda", generator=g)
    lib = A.pack_library(ref)
    out, idx, sc = M.run_match(src, lib, 4, 0.25, mode="screen")
    q = M.pack_queries(src)
    sc2, idx2 = M.search_topk(q, lib, 4, mode="screen")
    out2 = torch.empty((2, 77, 768), device="cuda")
    M.gather_mean(lib, idx2, q, 0.25, out2)
    assert torch.equal(idx.view(-1, 4), idx2) and torch.equal(sc.view(-1, 4), sc2) and torch.equal(out, out2)


def test_packed_library_save_load_roundtrip(tmp_path):
    """library lifecycle (generate_voice_library.py:42 / inference.py:78-82): the packed layout is
    stored next to the legacy `tokens` key; a reference-format checkpoint loads (and is packed) too"""
    g = torch.Generator(device="cuda").manual_seed(31)
    ref = torch.randn(1, 768, 5000, device="cuda", generator=g)
    src = torch.randn(1, 768, 40, device="cuda", generator=g)
    lib = A.pack_library(ref)
    p = str(tmp_path / "voice_library_packed.pt")
    A.save_packed_library(lib, p)
    lib2 = A.load_packed_library(p)
    for a, b in ((lib.raw, lib2.raw), (lib.norms, lib2.norms), (lib.packed, lib2.packed), (lib.stats, lib2.stats),
                 (lib.err, lib2.err)):
        assert torch.equal(a, b)
    want = A.match_features(src, ref)
    out, _, _ = A.match_packed(src, lib2)
    assert torch.equal(out.transpose(1, 2), want)
    blob = torch.load(p, weights_only=True)
    assert list(blob.keys()) == ["tokens"] and torch.equal(blob["tokens"], ref.cpu())     # the reference's own file
    strict = A.VoiceLibrary(num_tokens=5000)
    strict.load_state_dict(torch.load(p, weights_only=True))      # what inference.py:81 does, strict
    assert torch.equal(strict.tokens.detach(), ref.cpu())
    # a sidecar that no longer belongs to the tokens (they were edited) is ignored: the tokens are packed again
    edited = ref.cpu().clone()
    edited[0, :, 7] += 1.0
    torch.save({"tokens": edited}, p)
    lib_e = A.load_packed_library(p)
    assert torch.equal(lib_e.raw, edited[0].t().contiguous().cuda()) and torch.equal(lib_e.packed, A.pack_library(edited.cuda()).packed)
    # a set of per-speaker libraries keeps its item structure
    many = A.pack_libraries(torch.randn(3, 768, 700, device="cuda", generator=g))
    pm = str(tmp_path / "speakers.pt")
    A.save_packed_library(many, pm)
    many2 = A.load_packed_library(pm)
    assert many2.items == 3 and many2.n_item == 700 and torch.equal(many2.packed, many.packed) and torch.equal(many2.raw, many.raw)
    legacy = str(tmp_path / "voice_library.pt")
    vl = A.VoiceLibrary(num_tokens=5000)
    with torch.no_grad():
        vl.tokens.copy_(ref.cpu())
    torch.save(vl.state_dict(), legacy)                           # what generate_voice_library.py:42 writes
    lib3 = A.load_packed_library(legacy)
    assert torch.equal(lib3.packed, lib.packed) and torch.equal(lib3.raw, lib.raw)


def test_planted_neighbours_at_large_n():
    """size-independent property at a BASELINE-scale library (N = 2M frames, T = 4096): for every
    query four noisy copies with strictly decreasing similarity are planted at random library
    positions; the match must return exactly those positions in that order, and the features
    must be the sequential mean of those raw rows."""
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(77)
    N, T, D = 2_000_000, 4096, 768
    lib = M.alloc_packed(N, D, dev)
    src = torch.randn(1, D, T, device=dev, generator=g)
    pos = torch.randperm(N, device=dev, generator=g)[: 4 * T].view(T, 4)
    raw_rows = torch.empty((T, 4, D), device=dev)
    for c0 in range(0, N, 250_000):
        c1 = min(N, c0 + 250_000)
        x = torch.randn(D, c1 - c0, device=dev, generator=g)
        # plant: copy j of query t = q_t * scale_j + noise_j  (noise grows with j -> similarity drops)
        inside = (pos >= c0) & (pos < c1)
        tt, jj = inside.nonzero(as_tuple=True)
        if tt.numel():
            # relative noise 0.05 / 0.15 / 0.3 / 0.5 -> cosines ~0.999 / 0.989 / 0.958 / 0.894 (> 10 sigma apart)
            rel = torch.tensor([0.05, 0.15, 0.3, 0.5], device=dev)[jj]
            noise = torch.randn(D, tt.numel(), device=dev, generator=g) * rel[None, :]
            planted = (src[0][:, tt] + noise) * (1.0 + 0.5 * jj.float())[None, :]
            x[:, pos[tt, jj] - c0] = planted
            raw_rows[tt, jj] = planted.t()
        M.pack_into(lib, c0, x)
        del x
    out, idx, score = A.match_packed(src, lib, 4, 0.0)
    assert M.last_info.mode == "screen" and M.last_info.fallback_queries() == 0
    assert torch.equal(idx[0], pos), "planted neighbours not recovered exactly"
    assert (score[0, :, :-1] > score[0, :, 1:]).all()
    want = ((raw_rows[:, 0] + raw_rows[:, 1]) + raw_rows[:, 2] + raw_rows[:, 3]) / 4.0
    assert torch.equal(out[0], want)
    # sharded-by-rows view of the same library gives the same answer (2 shards on one device)
    half = N // 2
    tops = []
    for lo, hi in ((0, half), (half, N)):
        shard = lib.rows(lo, hi)
        _, i, s = M.run_match(src, shard, 4, 0.0, want_out=False)
        tops.append((s.view(T, 4), i.view(T, 4)))
    sc = torch.stack([t[0] for t in tops]).contiguous()
    ix = torch.stack([t[1] for t in tops]).contiguous()
    top_s = torch.empty((T, 4), device=dev)
    top_i = torch.empty((T, 4), dtype=torch.int64, device=dev)
    rc = _cabi.load().alive_knn_merge(sc.data_ptr(), ix.data_ptr(), 2, T, 4, top_s.data_ptr(), top_i.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream)
    _cabi.check(rc, "merge")
    assert torch.equal(top_i, pos) and torch.equal(top_s, score[0])


def test_functional_api_packs_the_library_once_and_follows_inplace_edits():
    """realtime_inference.py:165 passes the same `tgt` tensor every chunk: it must be packed once,
    re-packed when it is modified in place, and never confused with another tensor."""
    M.clear_pack_cache()
    g = torch.Generator(device="cuda").manual_seed(41)
    tgt = torch.randn(1, 768, 16000, device="cuda", generator=g)
    view = tgt[:, :, ::4]                                      # the [:, :, ::4] view of realtime_inference.py:88
    chunk = torch.randn(1, 768, 24, device="cuda", generator=g)
    a = A.match_features(chunk, view)
    n_entries = len(M._pack_cache)
    packed_first = next(iter(M._pack_cache.values()))[2]
    b = A.match_features(chunk, view)
    assert len(M._pack_cache) == n_entries and next(iter(M._pack_cache.values()))[2] is packed_first
    assert torch.equal(a, b)
    tgt[:, :, 0] += 1.0                                        # in-place edit through the base tensor
    c = A.match_features(chunk, view)
    assert next(iter(M._pack_cache.values()))[2] is not packed_first
    want = A.match_features(chunk, view.contiguous())
    assert torch.equal(c, want)


def test_batched_libraries_equal_per_item_matches():
    """B different libraries packed back to back and matched in ONE launch (BASELINE cfg5,
    train_decoder.py:134-135) == B separate single-library matches, bit for bit; also vs the oracle."""
    rng = np.random.default_rng(91)
    for (B, T, N, k, alpha, mode) in [(3, 130, 1500, 4, 0.0, "auto"), (5, 7, 300, 2, 0.25, "auto"),
                                      (2, 40, 900, 16, 0.0, "auto"), (4, 300, 5000, 4, 0.0, "screen"),
                                      (3, 3000, 2000, 4, 0.0, "screen")]:      # 9000 query frames: 32-frame pack CTAs across items
        src = rng.standard_normal((B, 768, T), dtype=np.float32)
        ref = rng.standard_normal((B, 768, N), dtype=np.float32)
        if B == 3:
            ref[1, :, 100:140] = ref[1, :, 60:100]                 # duplicated frames inside item 1
            src[1, :, :10] = ref[1, :, 60:70]                      # ... queried exactly: exact ties
        s, r = _cuda(src), _cuda(ref)
        out, idx = A.match_features(s, r, k, alpha, return_indices=True, mode=mode)
        # one pipeline, one query pack (+ the two collect-pass launches once an item is large enough to have one)
        assert M.last_info.launches == 1 + ((6 if M.last_info.collect else 4) if M.last_info.mode == "screen" else 2)
        assert M.last_info.collect == (mode == "screen" and B * T * N >= 2 ** 24)
        for b in range(B):
            o1, i1 = A.match_features(s[b:b + 1], r[b:b + 1], k, alpha, return_indices=True, mode=mode)
            assert torch.equal(idx[b:b + 1], i1) and torch.equal(out[b:b + 1], o1)
        _assert_parity(out, idx, src, ref, k, alpha)
    # explicit handle
    lib = A.pack_libraries(r)
    assert lib.items == B and lib.n_item == N
    o2, i2, _ = A.match_packed(s, lib, k, alpha)
    assert torch.equal(o2.transpose(1, 2), out) and torch.equal(i2, idx)


@pytest.mark.parametrize("D,mode", [(64, "screen"), (256, "screen"), (1024, "screen"), (100, "auto"), (1536, "screen")])
def test_other_feature_dims(D, mode):
    """VoiceLibrary(num_tokens, hubert_dim) takes any feature dim (voice_library.py:7): multiples of
    64 run on the tensor-core screen, everything else (multiple of 4) on the exact scan."""
    rng = np.random.default_rng(D)
    src = rng.standard_normal((2, D, 37), dtype=np.float32)
    tok = rng.standard_normal((1, D, 1300), dtype=np.float32)
    vl = A.VoiceLibrary(num_tokens=1300, hubert_dim=D).cuda()
    with torch.no_grad():
        vl.tokens.copy_(_cuda(tok))
    out, idx = vl.match(_cuda(src), k=4, alpha=0.1, return_indices=True, mode=mode)
    assert M.last_info.mode == ("exact" if D % 64 else "screen")
    want_out, want_idx, _ = O.voice_library_match_np(tok, src, 4, 0.1, True)
    ref_b = np.broadcast_to(tok, (2,) + tok.shape[1:])
    scores = O.cosine_scores_np(src, ref_b)
    ok, n_exact, n_tie, bad = O.indices_match_mod_ties(idx.cpu().numpy(), want_idx, scores, TIE_TOL)
    assert ok, bad
    same = (idx.cpu().numpy() == want_idx).all(axis=2)
    o = out.detach().cpu().numpy()
    assert np.array_equal(np.swapaxes(o, 1, 2)[same], np.swapaxes(want_out, 1, 2)[same])


def _clustered(T, N, nclus, noise, seed, dev="cuda"):
    g = torch.Generator(device=dev).manual_seed(seed)
    cent = torch.randn(768, nclus, device=dev, generator=g)
    ref = (cent[:, torch.randint(0, nclus, (N,), device=dev, generator=g)] +
           noise * torch.randn(768, N, device=dev, generator=g))[None]
    src = (cent[:, torch.randint(0, nclus, (T,), device=dev, generator=g)] +
           noise * torch.randn(768, T, device=dev, generator=g))[None]
    return src, ref


@pytest.mark.parametrize("fmt", ["bf16", "fp16"])
@pytest.mark.parametrize("refine", [True, False])
@pytest.mark.parametrize("T,N,nclus,noise,expect", [
    (300, 60_000, 60, 0.2, "collect"),        # ~1000 frames inside every query's first-pass band
    (300, 60_000, 12, 0.2, "overflow"),       # ~5000 per cluster: more than a candidate buffer holds WITHOUT refinement
    (2500, 8_000, 8, 0.2, "many"),            # (nearly) every query of a large batch uncertified
    (400, 50_000, 50, 0.5, "mixed"),
    (32, 600_000, 600, 0.2, "collect"),       # a realtime chunk (skinny first screen) against a clustered library
    (200, 100_000, 20, 0.05, "tight"),        # 5000 frames within ~3e-4 of each other per cluster
])
def test_collect_pass_on_clustered_libraries(T, N, nclus, noise, expect, refine, fmt, monkeypatch):
    """Tight clusters defeat the bf16 certificate; the second (collecting) tensor-core pass must give exactly what the
    exhaustive scan gives, whichever of its exits a query takes.  With the library's second bf16 plane (refine) the
    pass runs hi.hi + hi.lo + lo.hi against a cut tightened by a few exact rescorings, and NOTHING is left for the
    exhaustive scan on any of these libraries; without it the candidate buffers of the densest ones overflow.
    With fp16 planes (the default) the first-pass band is 8x narrower: most of these libraries certify at once."""
    monkeypatch.setattr(M, "SCREEN_FORMAT", fmt)
    src, ref = _clustered(T, N, nclus, noise, seed=T + N)
    lib = A.pack_library(ref, refine=refine)
    assert (lib.lo is not None) == refine and lib.format == M._FORMATS[fmt]
    out_s, idx_s, sc_s = M.run_match(src, lib, 4, 0.25, mode="screen")
    info = M.last_info
    fb, ex = info.fallback_queries(), info.exact_scan_queries()
    assert info.collect and info.launches == 1 + (7 if refine else 6)
    out_e, idx_e, sc_e = M.run_match(src, lib, 4, 0.25, mode="exact")
    assert torch.equal(idx_s, idx_e)
    assert torch.equal(out_s, out_e)
    assert torch.equal(sc_s, sc_e)
    if fmt == "bf16" and expect in ("collect", "overflow", "many", "tight"):
        assert fb > T // 2
    if fmt == "fp16" and expect == "tight":
        assert fb > T // 2                     # ~5000 frames within 3e-4: beyond any 16-bit screen
    if refine:
        assert ex == 0, (fb, ex)
    elif expect == "collect":
        assert ex == 0
    elif expect == "overflow" and fmt == "bf16":
        assert ex > T // 2
    # and against the oracle on a slice (the whole batch would take the CPU too long)
    sl = slice(0, 40)
    _assert_parity(out_s[:, sl].transpose(1, 2), idx_s[:, sl], src[:, :, sl].cpu().numpy(), ref.cpu().numpy(), 4, 0.25)


def test_collect_pass_for_batched_libraries():
    """per-speaker libraries (items > 1, BASELINE cfg5 / train_decoder.py:134-135) get the collect pass per item:
    clustered items are resolved by it, clean items are untouched, everything equals the exhaustive scan"""
    B, T, N = 3, 300, 30_000                       # B * T * N = 2.7e7 >= 2^24
    srcs, refs = [], []
    for b in range(B):
        s, r = _clustered(T, N, 30, 0.2, seed=100 + b) if b != 1 else \
            (torch.randn(1, 768, T, device="cuda"), torch.randn(1, 768, N, device="cuda"))
        srcs.append(s)
        refs.append(r)
    src, ref = torch.cat(srcs), torch.cat(refs)
    for refine in (False, True):
        lib = A.pack_libraries(ref, refine=refine, fmt="bf16")      # (bf16: the clustered items need the collect pass)
        assert lib.items == B and (lib.lo is not None) == refine
        out_s, idx_s, sc_s = M.run_match(src, lib, 4, 0.0, mode="screen")
        info = M.last_info
        assert info.collect
        fbs = info.fb_count.cpu().tolist()
        assert fbs[0] > T // 2 and fbs[1] == 0 and fbs[2] > T // 2
        assert info.exact_scan_queries() == 0
        out_e, idx_e, sc_e = M.run_match(src, lib, 4, 0.0, mode="exact")
        assert torch.equal(idx_s, idx_e) and torch.equal(out_s, out_e) and torch.equal(sc_s, sc_e)
    for b in (0, 2):
        sl = slice(0, 16)
        rel = idx_s[b:b + 1, sl] - b * N
        _assert_parity(out_s[b:b + 1, sl].transpose(1, 2), rel, src[b:b + 1, :, sl].cpu().numpy(), ref[b:b + 1].cpu().numpy(), 4, 0.0)


def test_tensor_core_accumulation_error_model():
    """The certificate's accumulation slack (select.cu accum_slack: TWICE one float32 ulp of the accumulator per
    tcgen05.mma) against the hardware: |screened score - float64 dot of the same bf16 operands| over every pair the
    fused kernel kept, on i.i.d. frames and on near-duplicates (scores up to 1: the worst case for a truncating
    accumulator).  The observed maximum must stay below HALF the slack, and the error must be one-sided (truncation)."""
    g = torch.Generator(device="cuda").manual_seed(1)
    for D, fmt in ((768, "bf16"), (768, "fp16"), (1536, "bf16"), (1536, "fp16")):
        slack = (D // 16) * 2.0 ** -22 + 2e-7
        T, N = 256, 40_000
        cases = [(torch.randn(D, T, device="cuda", generator=g), torch.randn(D, N, device="cuda", generator=g))]
        for noise in (0.2, 0.01):
            cent = torch.randn(D, 20, device="cuda", generator=g)
            cases.append((cent[:, torch.randint(0, 20, (T,), device="cuda", generator=g)] + noise * torch.randn(D, T, device="cuda", generator=g),
                          cent[:, torch.randint(0, 20, (N,), device="cuda", generator=g)] + noise * torch.randn(D, N, device="cuda", generator=g)))
        cases.append((torch.rand(D, T, device="cuda", generator=g) + 0.5, torch.rand(D, N, device="cuda", generator=g) + 0.5))
        for src, ref in cases:
            q, lib = M.pack_frames(src, fmt=fmt), M.pack_frames(ref, fmt=fmt)
            for variant in (1, 2):
                plan = M.make_plan(T, N, D, q.device, variant, fmt)
                cs = torch.empty((T, plan.lists * 8), device="cuda")
                ci = torch.empty((T, plan.lists * 8), dtype=torch.int32, device="cuda")
                _cabi.check(_cabi.load().alive_knn_search(q.packed.data_ptr(), lib.packed.data_ptr(), ctypes.byref(plan),
                                                          cs.data_ptr(), ci.data_ptr(), torch.cuda.current_stream().cuda_stream), "search")
                ok = ci >= 0
                worst, most_positive = 0.0, 0.0
                for t0 in range(0, T, 64):
                    rows = lib.packed[ci[t0:t0 + 64].clamp(min=0).long()].double()
                    exact = (rows * q.packed[t0:t0 + 64].double()[:, None, :]).sum(dim=2)
                    err = torch.where(ok[t0:t0 + 64], cs[t0:t0 + 64].double() - exact, torch.zeros_like(exact))
                    worst = max(worst, float(err.abs().max()))
                    most_positive = max(most_positive, float(err.max()))
                assert worst <= 0.5 * slack, (D, variant, worst, slack)
                assert most_positive <= 2.0 ** -23, (D, variant, most_positive)     # truncation: never above by more than a rounding


def test_collect_pass_with_unusable_cut():
    """Non-finite library rows void the certificate AND the cut: every query must reach the exhaustive scan."""
    g = torch.Generator(device="cuda").manual_seed(3)
    ref = torch.randn(1, 768, 90_000, device="cuda", generator=g)
    ref[0, :, 777] = 0.0                                   # zero row -> NaN similarity, ranks first (torch.topk)
    src = torch.randn(1, 768, 200, device="cuda", generator=g)
    lib = A.pack_library(ref)
    out_s, idx_s, _ = M.run_match(src, lib, 4, 0.0, mode="screen")
    info = M.last_info
    assert info.collect and info.fallback_queries() == 200 and info.exact_scan_queries() == 200
    out_e, idx_e, _ = M.run_match(src, lib, 4, 0.0, mode="exact")
    assert torch.equal(idx_s, idx_e)
    assert (idx_s[..., 0] == 777).all()


def test_collect_pass_on_a_row_shard():
    """The collect pass inside a row shard (multi-GPU local top-k: row_base != 0, no gather): global
    indices, and the two shards merged must equal the unsharded answer."""
    T, N = 500, 80_000                     # per shard T * N/2 = 2e7 >= 2^24: the collect pass is on
    src, ref = _clustered(T, N, 40, 0.2, seed=17)
    lib = A.pack_library(ref, fmt="bf16")  # (bf16 planes: at this cluster width nothing certifies in the first pass)
    _, want_idx, want_sc = M.run_match(src, lib, 4, 0.0, mode="exact")
    half = N // 2
    tops, uncertified = [], 0
    for lo, hi in ((0, half), (half, N)):
        shard = lib.rows(lo, hi)
        assert shard.lo is not None
        _, i, s = M.run_match(src, shard, 4, 0.0, want_out=False)
        assert M.last_info.collect
        uncertified += M.last_info.fallback_queries()
        assert M.last_info.exact_scan_queries() == 0
        assert int(i.min()) >= lo and int(i.max()) < hi
        tops.append((s.view(T, 4), i.view(T, 4)))
    assert uncertified > T
    sc = torch.stack([t[0] for t in tops]).contiguous()
    ix = torch.stack([t[1] for t in tops]).contiguous()
    top_s = torch.empty((T, 4), device="cuda")
    top_i = torch.empty((T, 4), dtype=torch.int64, device="cuda")
    rc = _cabi.load().alive_knn_merge(sc.data_ptr(), ix.data_ptr(), 2, T, 4, top_s.data_ptr(), top_i.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream)
    _cabi.check(rc, "merge")
    assert torch.equal(top_i, want_idx[0]) and torch.equal(top_s, want_sc[0])
