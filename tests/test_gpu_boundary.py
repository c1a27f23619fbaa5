"""The drop-in boundary under the reference's real call contexts: the registered torch custom op
(fake kernel, autograd formula, torch.compile), `torch.inference_mode()` + fp16 autocast exactly as
realtime_inference.py:143-165 calls the match, the pack cache's blind spots (`.data` writes, inference tensors),
several devices in one process, and the library-builder / file-format contracts."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import alive_vc_b200 as A                                          # noqa: E402
from alive_vc_b200 import matching as M, ops                        # noqa: E402
from alive_vc_b200.lifecycle import LibraryBuilder                  # noqa: E402
from oracle import knn_oracle as O                                  # noqa: E402


def _cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _oracle_check(out, idx, src, ref, k, alpha):
    o_out, o_idx, _ = O.match_features_np(src, ref, k, alpha, True)
    scores = O.cosine_scores_np(src, ref)
    ok, _, _, bad = O.indices_match_mod_ties(idx.cpu().numpy(), o_idx, scores, 1e-6)
    assert ok, bad
    same = (idx.cpu().numpy() == o_idx).all(axis=2)
    o = out.float().cpu().numpy()
    assert np.array_equal(np.swapaxes(o, 1, 2)[same], np.swapaxes(o_out, 1, 2)[same])


@pytest.mark.parametrize("lib_inside_inference_mode", [False, True])
@pytest.mark.parametrize("fp16", [False, True])
def test_realtime_call_context(lib_inside_inference_mode, fp16):
    """realtime_inference.py:79-96 builds `tgt` (cat of an encoder output and VL.tokens: it REQUIRES GRAD) at module
    level, then every chunk calls match_features inside inference_mode + autocast(enabled=fp16) with the encoder's
    output (fp16 under autocast) - :143-165.  The library may also be created inside inference mode (an inference
    tensor: no version counter)."""
    rng = np.random.default_rng(3)
    N, T = 3512, 24
    tok = torch.nn.Parameter(_cuda(rng.standard_normal((1, 768, 512), dtype=np.float32)))
    utt = _cuda(rng.standard_normal((1, 768, 4 * (N - 512)), dtype=np.float32))

    def build():
        return torch.cat([utt.detach()[:, :, ::4], tok], dim=2)          # :88, :94
    if lib_inside_inference_mode:
        with torch.inference_mode():
            tgt = build()
        assert tgt.is_inference()
    else:
        tgt = build()
        assert tgt.requires_grad
    ref_np = tgt.detach().cpu().numpy()
    for chunk_no in range(3):
        content32 = _cuda(rng.standard_normal((1, 768, T), dtype=np.float32))
        with torch.inference_mode():
            with torch.autocast("cuda", enabled=fp16):
                content = content32.half() if fp16 else content32      # what CE(spec) returns under autocast
                out = A.match_features(content, tgt, k=4, alpha=0.0)
                out2, idx2 = A.match_features(content, tgt, k=4, alpha=0.25, return_indices=True)
        assert out.dtype == torch.float32 and tuple(out.shape) == (1, 768, T)     # promote(fp16, fp32) like the reference
        assert tuple(out.stride()) == (T * 768, 1, 768)
        src_np = content.float().cpu().numpy()                                  # the values the match really saw
        _oracle_check(out, idx2, src_np, ref_np, 4, 0.0)
        o2 = O.match_features_np(src_np, ref_np, 4, 0.25)
        np.testing.assert_allclose(out2.cpu().numpy(), o2, rtol=1e-5, atol=1e-6)
    if lib_inside_inference_mode:
        assert not any(k[0] == id(tgt) for k in M._pack_cache)       # never cached: nothing to validate a copy against


def test_custom_op_is_registered_with_fake_and_autograd():
    assert torch.ops.alive_vc_b200.knn_match.default is not None
    g = torch.Generator(device="cuda").manual_seed(1)
    src = torch.randn(2, 768, 17, device="cuda", generator=g, requires_grad=True)
    tok = torch.randn(1, 768, 900, device="cuda", generator=g, requires_grad=True)
    torch.library.opcheck(torch.ops.alive_vc_b200.knn_match.default, (src, tok, 4, 0.25, "auto", 0, True),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    rows = torch.randn(34, 768, device="cuda", generator=g)
    idx = torch.randint(0, 900, (34, 4), device="cuda", generator=g)
    torch.library.opcheck(torch.ops.alive_vc_b200.knn_scatter_grad.default, (rows, idx, 900, 0.25),
                          test_utils=("test_schema", "test_faketensor"))
    # the op's autograd formula == the reference's gradient (golden vl_* fixtures check the values; here: both paths agree)
    out_btd, idx, _ = ops.knn_match(src, tok, 4, 0.25, "auto", 0, True)
    gout = torch.randn_like(out_btd)
    out_btd.backward(gout)
    want_tok = torch.zeros(900, 768, device="cuda")
    want_tok.index_add_(0, idx.reshape(-1), (gout.reshape(-1, 1, 768).expand(-1, 4, -1) * (0.75 / 4)).reshape(-1, 768))
    torch.testing.assert_close(tok.grad, want_tok.t()[None], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(src.grad, gout.transpose(1, 2) * 0.25)
    # reference_grad=False (match_features): the library gets nothing
    tok.grad = None
    ops.knn_match(src, tok, 4, 0.25, "auto", 0, False)[0].sum().backward()
    assert tok.grad is None


def test_torch_compile_traces_through_the_op():
    """a compiled caller (aot_eager: dynamo + AOT autograd with the op's fake kernel and autograd formula; no
    inductor codegen needed) gives the eager bits, forward and backward"""
    g = torch.Generator(device="cuda").manual_seed(2)
    vl = A.VoiceLibrary(num_tokens=1500).cuda()
    ref = torch.randn(1, 768, 3000, device="cuda", generator=g)
    src = torch.randn(2, 768, 40, device="cuda", generator=g)

    def caller(x, r):
        y = A.match_features(x * 1.0, r.expand(x.shape[0], -1, -1), k=4, alpha=0.25)
        z = vl.match(y, k=4, alpha=0.5)
        return z * 2.0

    eager = caller(src, ref)
    eager.sum().backward()
    g_eager = vl.tokens.grad.clone()
    vl.tokens.grad = None
    compiled = torch.compile(caller, backend="aot_eager", fullgraph=True)
    out = compiled(src, ref)
    assert torch.equal(out, eager)
    out.sum().backward()
    torch.testing.assert_close(vl.tokens.grad, g_eager)


def test_data_writes_and_explicit_invalidation():
    """`VL.tokens.data[:, :, n] = t` (generate_voice_library.py:38) does not bump the version counter: a default-sized
    VoiceLibrary is re-packed every call (always current); a large one needs VoiceLibrary.invalidate()."""
    g = torch.Generator(device="cuda").manual_seed(4)
    src = torch.randn(1, 768, 8, device="cuda", generator=g)
    for n_tok, cached in ((512, False), (4000, True)):
        vl = A.VoiceLibrary(num_tokens=n_tok).cuda()
        assert (vl.tokens.numel() >= M.PACK_CACHE_MIN_ELEMENTS) == cached
        _, idx0 = vl.match(src, return_indices=True)
        j = next(i for i in range(n_tok) if i not in idx0[0, 0].tolist())   # a frame query 0 did NOT select
        vl.tokens.data[:, :, j] = src[0, :, 0] * 3.0                     # frame j := a copy of query 0
        _, idx1 = vl.match(src, return_indices=True)
        if cached:
            assert torch.equal(idx1, idx0)                               # stale by design: the write was invisible ...
            vl.invalidate()                                              # ... until told
            _, idx1 = vl.match(src, return_indices=True)
        assert int(idx1[0, 0, 0]) == j


def test_d_limit_is_checked_before_any_launch():
    src = torch.randn(1, 2048, 8, device="cuda")
    ref = torch.randn(1, 2048, 300, device="cuda")
    before = M.launch_count
    with pytest.raises(RuntimeError, match="1536"):
        A.match_features(src, ref)
    # only the library pack ran (it has the same limit and fails first), nothing of the match chain
    assert M.launch_count == before


def test_library_builder_unwritten_slots_and_randn_init():
    """generate_voice_library.py:30-42 starts from VoiceLibrary()'s random-normal tokens; a builder that starts empty
    must not silently turn untouched slots into zero frames (NaN similarity, rank first in every match)."""
    g = torch.Generator(device="cuda").manual_seed(6)
    frames = torch.randn(768, 512, device="cuda", generator=g)
    slots = torch.randint(0, 512, (512,), device="cuda", generator=g)
    b = LibraryBuilder(d=768, capacity=512)
    b.put(slots, frames)
    missing = b.unwritten_slots()
    assert 100 < missing.numel() < 300                           # ~1/e of the slots are never drawn
    for fn in (b.tokens, b.packed, lambda: b.save("/tmp/should_not_exist.pt")):
        with pytest.raises(RuntimeError, match="never written"):
            fn()
    b2 = LibraryBuilder(d=768, capacity=512, init="randn", generator=torch.Generator(device="cuda").manual_seed(7))
    assert len(b2) == 512 and b2.unwritten_slots().numel() == 0
    start = b2.tokens().clone()
    b2.put(slots, frames)
    got = b2.tokens()
    want = start.clone()
    for i in range(512):                                          # the reference loop, literally
        want[:, :, int(slots[i])] = frames[:, i]
    assert torch.equal(got, want)
    lib = b2.packed()
    assert int(lib.stats.cpu().numpy().view(np.uint32)[1]) == 0   # no zero rows
    src = torch.randn(1, 768, 16, device="cuda", generator=g)
    out, idx, _ = A.match_packed(src, lib, 4, 0.0)
    assert M.last_info.fallback_queries() == 0
    _oracle_check(out.transpose(1, 2), idx, src.cpu().numpy(), want.cpu().numpy(), 4, 0.0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process():
    """per-device state in the C library (function attributes, pacing counters) and device guards in the host code:
    the current device stays cuda:0 while the tensors live on cuda:1"""
    rng = np.random.default_rng(8)
    src = rng.standard_normal((1, 768, 300), dtype=np.float32)
    ref = rng.standard_normal((1, 768, 70_000), dtype=np.float32)
    outs = []
    torch.cuda.set_device(0)
    for dev in ("cuda:0", "cuda:1", "cuda:0", "cuda:1"):
        s, r = torch.from_numpy(src).to(dev), torch.from_numpy(ref).to(dev)
        out, idx = A.match_features(s, r, 4, 0.0, return_indices=True)
        assert out.device == torch.device(dev) and torch.cuda.current_device() == 0
        outs.append((out.cpu(), idx.cpu()))
        chunk = s[:, :, :24].contiguous()
        sm = A.StreamingMatcher(A.pack_library(r), T=24)
        assert torch.equal(sm(chunk).cpu(), out[:, :, :24].cpu())
    for o, i in outs[1:]:
        assert torch.equal(o, outs[0][0]) and torch.equal(i, outs[0][1])
    _oracle_check(outs[0][0], outs[0][1], src, ref, 4, 0.0)


def test_auto_format_probe_picks_fp16_for_clustered_libraries():
    """pack_library's default ("auto"): bf16 planes unless a probe with 256 of the library's own frames finds it
    clustered, then IEEE fp16 (8x finer rounding -> the certificate clears in the first pass).  Results are the
    exhaustive scan's either way."""
    assert M.SCREEN_FORMAT == "auto"
    g = torch.Generator(device="cuda").manual_seed(12)
    spread = torch.randn(1, 768, 50_000, device="cuda", generator=g)
    lib = A.pack_library(spread)
    assert lib.format == 0 and lib.packed.dtype == torch.bfloat16 and lib.lo is not None
    assert M.probe_uncertified_fraction(lib) == 0.0
    cent = torch.randn(768, 50, device="cuda", generator=g)
    clustered = (cent[:, torch.randint(0, 50, (50_000,), device="cuda", generator=g)] +
                 0.2 * torch.randn(768, 50_000, device="cuda", generator=g))[None]
    src = (cent[:, torch.randint(0, 50, (500,), device="cuda", generator=g)] + 0.2 * torch.randn(768, 500, device="cuda", generator=g))[None]
    forced = A.pack_library(clustered, fmt="bf16")
    assert M.probe_uncertified_fraction(forced) > 0.5
    lib = A.pack_library(clustered)
    assert lib.format == 1 and lib.packed.dtype == torch.float16
    out, idx, sc = A.match_packed(src, lib, 4, 0.0)
    assert M.last_info.mode == "screen" and M.last_info.fallback_queries() < 50      # certified in the first pass
    out_b, idx_b, sc_b = A.match_packed(src, forced, 4, 0.0)
    assert M.last_info.fallback_queries() > 250 and M.last_info.exact_scan_queries() == 0
    out_e, idx_e, sc_e = M.run_match(src, lib, 4, 0.0, mode="exact")
    for o, i, s_ in ((out, idx, sc), (out_b, idx_b, sc_b)):
        assert torch.equal(i, idx_e) and torch.equal(o, out_e) and torch.equal(s_, sc_e)
    # the functional API inherits the choice through its pack cache
    want = A.match_features(src, clustered, 4, 0.0)
    assert torch.equal(want, out_e.transpose(1, 2))
    # small libraries are never probed; libraries with non-finite rows keep bf16
    assert A.pack_library(clustered[:, :, :4000]).format == 0
    broken = clustered.clone()
    broken[0, :, 5] = 0.0
    assert A.pack_library(broken).format == 0
