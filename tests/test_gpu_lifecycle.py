"""GPU tests of the callers either side of the match: LibraryBuilder
(generate_voice_library.py:30-42), match_windows (inference.py:96-134) and HostStreamingMatcher
(realtime_inference.py:158-176), each against the oracle / the plain per-call path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import alive_vc_b200 as A                                   # noqa: E402
from alive_vc_b200 import matching as M                      # noqa: E402
from alive_vc_b200.lifecycle import HostStreamingMatcher, LibraryBuilder, match_windows   # noqa: E402
from oracle import knn_oracle as O                           # noqa: E402


def _cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def test_builder_put_follows_the_reference_generation_loop():
    """generate_voice_library.py:36-38: sequential writes into random slots, the last write wins,
    untouched slots keep the initial tokens."""
    rng = np.random.default_rng(5)
    D, N, writes = 768, 512, 513
    tokens = rng.standard_normal((1, D, N), dtype=np.float32)           # VoiceLibrary() initial randn
    slots = rng.integers(0, N, size=writes)
    frames = rng.standard_normal((D, writes), dtype=np.float32)
    want = tokens.copy()
    for i in range(writes):                                             # the reference loop, literally
        want[:, :, slots[i]] = frames[:, i]
    b = LibraryBuilder(d=D, capacity=N, tokens=_cuda(tokens))
    assert len(b) == N
    b.put(slots, _cuda(frames))
    got = b.tokens()
    assert tuple(got.shape) == (1, D, N)
    assert np.array_equal(got.cpu().numpy(), want)
    # the packed layout is what pack_library makes of the same tokens, bit for bit
    p, q = b.packed(), A.pack_library(_cuda(want))
    for name in ("raw", "norms", "packed", "err"):
        assert torch.equal(getattr(p, name), getattr(q, name)), name
    # and it matches like the reference library does
    src = rng.standard_normal((1, D, 40), dtype=np.float32)
    out, idx, _ = A.match_packed(_cuda(src), p, 4, 0.0)
    w_out, w_idx, _ = O.match_features_np(src, want, 4, 0.0, True)
    assert np.array_equal(idx.cpu().numpy(), w_idx)
    assert np.array_equal(out.cpu().numpy(), np.swapaxes(w_out, 1, 2))


def test_builder_append_grows_and_saves(tmp_path):
    """inference.py:67-84: library = cat([CE(target), VL.tokens], dim=2), any N."""
    rng = np.random.default_rng(6)
    D = 768
    parts = [rng.standard_normal((1, D, n), dtype=np.float32) for n in (130, 1, 700, 0, 2049)]
    b = LibraryBuilder(d=D, capacity=16)
    for p in parts:
        b.append(_cuda(p))
    want = np.concatenate(parts, axis=2)
    assert len(b) == want.shape[2]
    assert np.array_equal(b.tokens().cpu().numpy(), want)
    path = str(tmp_path / "voice_library.pt")
    b.save(path)
    blob = torch.load(path, map_location="cpu", weights_only=True)
    assert np.array_equal(blob["tokens"].numpy(), want)                 # the reference's own key
    lib = A.load_packed_library(path)
    src = rng.standard_normal((1, D, 33), dtype=np.float32)
    _, idx, _ = A.match_packed(_cuda(src), lib, 4, 0.0)
    _, w_idx, _ = O.match_features_np(src, want, 4, 0.0, True)
    assert np.array_equal(idx.cpu().numpy(), w_idx)
    # a strided view (realtime_inference.py:88 subsamples [:, :, ::4]) appends like its copy
    b2 = LibraryBuilder(d=D, capacity=4)
    big = _cuda(parts[2])
    b2.append(big[:, :, ::4])
    assert np.array_equal(b2.tokens().cpu().numpy(), parts[2][:, :, ::4])


def test_builder_errors():
    b = LibraryBuilder(d=8, capacity=4)
    with pytest.raises(RuntimeError):
        b.packed()                                                      # empty library
    with pytest.raises(RuntimeError):
        b.append(torch.zeros(1, 7, 3, device="cuda"))                   # wrong channel count
    with pytest.raises(RuntimeError):
        b.put([0, 1], torch.zeros(8, 3, device="cuda"))                 # slots / frames mismatch
    with pytest.raises(IndexError):
        b.put([-1], torch.zeros(8, 1, device="cuda"))


@pytest.mark.parametrize("lens", [[150, 150, 150, 150], [150, 150, 37], [1], [0, 5, 0]])
def test_match_windows_equals_one_call_per_window(lens):
    """inference.py:129 runs the match once per overlapped window; one batched launch must give the
    same bits."""
    rng = np.random.default_rng(7)
    D, N, k, alpha = 768, 3512, 4, 0.25
    ref = rng.standard_normal((1, D, N), dtype=np.float32)
    lib = A.pack_library(_cuda(ref))
    wins = [_cuda(rng.standard_normal((1, D, n), dtype=np.float32)) for n in lens]
    got = match_windows(wins, lib, k, alpha)
    assert len(got) == len(wins)
    for w, g in zip(wins, got):
        assert tuple(g.shape) == tuple(w.shape)
        if w.shape[2] == 0:
            continue
        want = A.match_features(w, _cuda(ref), k, alpha)
        assert torch.equal(g, want)
        w_out = O.match_features_np(w.cpu().numpy(), ref, k, alpha)
        np.testing.assert_allclose(g.cpu().numpy(), w_out, rtol=1e-5, atol=1e-6)
    if len(set(lens)) == 1:                                             # the [W, D, Tw] tensor form
        stacked = torch.cat(wins, dim=0)
        got2 = match_windows(stacked, lib, k, alpha)
        for a, b_ in zip(got, got2):
            assert torch.equal(a, b_)


@pytest.mark.parametrize("T,N,batch", [(24, 3512, 1), (32, 20000, 1), (16, 1000, 3)])
def test_host_streaming_matcher(T, N, batch):
    """realtime_inference.py:158-176 with host buffers: graph(H2D + pipeline + D2H) == plain call."""
    rng = np.random.default_rng(8)
    D, k, alpha = 768, 4, 0.0
    ref = rng.standard_normal((1, D, N), dtype=np.float32)
    lib = A.pack_library(_cuda(ref))
    hm = HostStreamingMatcher(lib, T, k, alpha, batch=batch)
    for it in range(3):
        chunk = torch.from_numpy(rng.standard_normal((batch, D, T), dtype=np.float32))
        out = hm(chunk)
        assert not out.is_cuda and tuple(out.shape) == (batch, D, T)
        want, _, _ = A.match_packed(chunk.cuda(), lib, k, alpha)
        assert torch.equal(out, want.transpose(1, 2).cpu())
        w_out = O.match_features_np(chunk.numpy(), np.broadcast_to(ref, (batch, D, N)), k, alpha)
        np.testing.assert_allclose(out.numpy(), w_out, rtol=1e-5, atol=1e-6)
    # submit / result split: the host may do other work between the two
    chunk = torch.from_numpy(rng.standard_normal((batch, D, T), dtype=np.float32))
    hm.submit(chunk)
    want, _, _ = A.match_packed(chunk.cuda(), lib, k, alpha)
    assert torch.equal(hm.result(), want.transpose(1, 2).cpu())


# ---- row-major producer format ------------------------------------------------
@pytest.mark.parametrize("T,N", [(24, 3512), (450, 3512), (1000, 20_000), (9000, 40_000)])
def test_match_rows_equals_match_features(T, N):
    """[T, D] row-major frames in, [T, D] out: bit-identical to the channel-major drop-in (and to the
    oracle) - the per-call transposes of common.py:100/108 simply do not happen."""
    from alive_vc_b200.lifecycle import match_rows, pack_rows
    rng = np.random.default_rng(T + N)
    D = 768
    q_rows = rng.standard_normal((T, D), dtype=np.float32)
    l_rows = rng.standard_normal((N, D), dtype=np.float32)
    lib = pack_rows(_cuda(l_rows))
    ref_lib = A.pack_library(_cuda(l_rows.T.copy()[None]))               # the reference's [1, D, N] layout
    for name in ("raw", "norms", "packed"):
        assert torch.equal(getattr(lib, name), getattr(ref_lib, name)), name
    # the screening-error norm is summed in a layout-specific order (it only feeds the certificate's bound)
    assert torch.allclose(lib.err, ref_lib.err, rtol=1e-5, atol=0)
    out, idx = match_rows(_cuda(q_rows), lib, 4, 0.25, return_indices=True)
    assert tuple(out.shape) == (T, D) and out.is_contiguous()
    w_out, w_idx = A.match_features(_cuda(q_rows.T.copy()[None]), _cuda(l_rows.T.copy()[None]), 4, 0.25,
                                    return_indices=True)
    assert torch.equal(idx, w_idx[0])
    assert torch.equal(out, w_out[0].t())
    if T * N <= 2_000_000:
        o_out, o_idx, _ = O.match_features_np(q_rows.T[None], l_rows.T[None], 4, 0.25, True)
        scores = O.cosine_scores_np(q_rows.T[None], l_rows.T[None])
        ok, _, _, bad = O.indices_match_mod_ties(idx[None].cpu().numpy(), o_idx, scores, 1e-6)
        assert ok, bad
        np.testing.assert_allclose(out.cpu().numpy(), o_out[0].T, rtol=1e-5, atol=1e-6)


def test_match_rows_batched_strided_and_raw_library():
    """[B, T, D] with a non-contiguous frame stride, library handed over as a plain [N, D] tensor."""
    from alive_vc_b200.lifecycle import match_rows
    rng = np.random.default_rng(77)
    D, B, T, N = 768, 3, 50, 5000
    wide = _cuda(rng.standard_normal((B, T, 2 * D), dtype=np.float32))
    frames = wide[:, :, :D]                                               # row stride 2*D
    l_rows = _cuda(rng.standard_normal((N, D), dtype=np.float32))
    out, idx = match_rows(frames, l_rows, 4, 0.0, return_indices=True)
    assert tuple(out.shape) == (B, T, D)
    w_out, w_idx = A.match_features(frames.transpose(1, 2).contiguous(),
                                    l_rows.t().contiguous()[None].expand(B, D, N), 4, 0.0, return_indices=True)
    assert torch.equal(idx, w_idx)
    assert torch.equal(out, w_out.transpose(1, 2))
    # batch items that are NOT uniformly strided (a [B, 2T, D] buffer, first T frames of each item), enough
    # frames for the register-resident row-major pack kernel: one pack launch with a per-item frame map
    T2 = 200
    tall = _cuda(rng.standard_normal((B, 2 * T2, D), dtype=np.float32))
    fr2 = tall[:, :T2]
    out2, idx2 = match_rows(fr2, l_rows, 4, 0.0, return_indices=True)
    w2, wi2 = A.match_features(fr2.transpose(1, 2).contiguous(), l_rows.t().contiguous()[None].expand(B, D, N), 4, 0.0,
                               return_indices=True)
    assert torch.equal(idx2, wi2) and torch.equal(out2, w2.transpose(1, 2))
    # autograd: the frames receive alpha * grad, like `source` of match_features (common.py:109)
    leaf = frames.clone().requires_grad_(True)
    match_rows(leaf, l_rows, 4, 0.25).sum().backward()
    assert torch.equal(leaf.grad, torch.full_like(leaf, 0.25))
    # empty chunk and dtype passthrough
    e = match_rows(frames[:, :0], l_rows)
    assert tuple(e.shape) == (B, 0, D)
    h = match_rows(frames.half(), l_rows)
    assert h.dtype == torch.float16


def test_fast_pack_kernels_equal_the_generic_kernel(tmp_path):
    """K1's layout-specific kernels (16-byte cp.async tile / register-resident rows, reciprocal-based
    correctly rounded division) write exactly the frames, norms and bf16 rows the generic kernel writes
    (full IEEE division per element); `err` is summed in fp32 instead of fp64 and may differ by 1e-5
    relative.  Channel-major and row-major inputs, ragged last CTA, unaligned fallback, zero / huge /
    tiny / non-finite rows, denormal and negative-zero elements."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, torch, hashlib, numpy as np\n"
        "from alive_vc_b200 import matching as M\n"
        "g = torch.Generator(device='cuda').manual_seed(3)\n"
        "res = {}\n"
        "for n in (20000, 20012, 20011, 9001):\n"
        "    x = torch.randn(768, n, device='cuda', generator=g)\n"
        "    x[:, 5] = 0.0\n"
        "    x[:, 6] *= 1e30\n"
        "    x[:, 7] *= 1e-30\n"
        "    x[:, 8] *= 3e-13\n"
        "    x[:, 9] *= 2e12\n"
        "    x[3, 10] = float('inf')\n"
        "    x[4, 11] = float('nan')\n"
        "    x[0:64, 12] = -0.0\n"
        "    x[64:128, 12] = 1e-42\n"
        "    x[128:192, 12] *= 1e-20\n"
        "    x[:, 13] *= torch.logspace(-30, 8, 768, device='cuda')\n"
        "    for tag, v in (('cm', x), ('rm', x.t().contiguous().t())):\n"
        "        p = M.pack_frames(v, fmt='bf16')\n"
        "        h = hashlib.sha256()\n"
        "        for f in ('raw', 'norms', 'packed', 'lo'):\n"
        "            h.update(getattr(p, f).view(torch.uint8).cpu().numpy().tobytes())\n"
        "        res[f'{tag}{n}_hash'] = np.frombuffer(h.digest(), dtype=np.uint8)\n"
        "        res[f'{tag}{n}_err'] = p.err.cpu().numpy()\n"
        "        res[f'{tag}{n}_err2_err'] = p.err2.cpu().numpy()\n"
        "        res[f'{tag}{n}_stats'] = p.stats.cpu().numpy()\n"
        "        p1 = M.pack_frames(v, refine=False, fmt='bf16')       # one plane\n"
        "        assert p1.lo is None\n"
        "        for f in ('raw', 'norms', 'packed', 'err'):              # bit patterns (the rows hold NaNs)\n"
        "            assert torch.equal(getattr(p1, f).view(torch.uint8), getattr(p, f).view(torch.uint8)), (tag, n, f)\n"
        "        assert torch.equal(p1.stats[:2], p.stats[:2]), (tag, n)\n"
        "np.savez(sys.argv[1], **res)\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for flag in ("1", "0"):
        path = str(tmp_path / f"pack{flag}.npz")
        env = dict(os.environ, ALIVE_KNN_PACK_FAST=flag)
        r = subprocess.run([sys.executable, "-c", code, path], capture_output=True, text=True, env=env, cwd=root,
                           timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(np.load(path))
    fast, gen = outs
    for key in gen.files:
        if key.endswith("_hash"):
            assert np.array_equal(fast[key], gen[key]), key
        elif key.endswith("_err"):
            np.testing.assert_allclose(fast[key], gen[key], rtol=1e-5, atol=0, err_msg=key)
        else:
            assert fast[key][1] == gen[key][1], key                      # count of non-finite rows
            a, b = fast[key][:1].view(np.float32)[0], gen[key][:1].view(np.float32)[0]
            assert abs(a - b) <= 1e-5 * b, key


# ---- (f)3: caller modules inside the chunk graph; (f)4: the producer-side pack ---------------------------------
class _TinyEncoder(torch.nn.Module):
    """stand-in with the structure of module/content_encoder.py:8-25 (the reference's modules cannot travel to the
    GPU box): 1x1 conv in, a depthwise + pointwise block, 1x1 conv out to 768 channels"""

    def __init__(self, c_in=65, c_mid=96, d=768):
        super().__init__()
        self.input_layer = torch.nn.Conv1d(c_in, c_mid, 1)
        self.mid_layers = torch.nn.Sequential(torch.nn.Conv1d(c_mid, c_mid, 7, padding=3, groups=c_mid), torch.nn.GELU(),
                                              torch.nn.Conv1d(c_mid, c_mid, 1))
        self.output_layer = torch.nn.Conv1d(c_mid, d, 1)

    def forward(self, x):
        return self.output_layer(self.mid_layers(self.input_layer(x)))


@pytest.mark.parametrize("zero_copy", [False, "out", "both"])
def test_host_streaming_matcher_copy_modes(zero_copy):
    rng = np.random.default_rng(18)
    D, T, N = 768, 32, 30_000
    ref = rng.standard_normal((1, D, N), dtype=np.float32)
    lib = A.pack_library(_cuda(ref))
    hm = HostStreamingMatcher(lib, T, 4, 0.25, zero_copy=zero_copy)
    for _ in range(3):
        chunk = torch.from_numpy(rng.standard_normal((1, D, T), dtype=np.float32))
        out = hm(chunk)
        want, _, _ = A.match_packed(chunk.cuda(), lib, 4, 0.25)
        assert torch.equal(out, want.transpose(1, 2).cpu())
    hm.src_host.copy_(chunk * 2.0)                       # fill the pinned buffer in place, then run()
    hm.run()
    want, _, _ = A.match_packed((chunk * 2.0).cuda(), lib, 4, 0.25)
    assert torch.equal(hm.result(), want.transpose(1, 2).cpu())


@pytest.mark.parametrize("early", [True, False])
def test_host_streaming_matcher_early_wait_and_fallback_chunks(early):
    """zero_copy="out": `result()` returns on the flag the notify kernel raises behind finish_kernel (early=True) -
    also when a chunk holds queries the screen cannot certify (duplicated library frames, a zero query): the flag then
    carries the "uncertified" bit and result() waits for the whole graph, so the fallback chain's rows are there."""
    rng = np.random.default_rng(28)
    D, T, N = 768, 32, 30_000
    ref = rng.standard_normal((1, D, N), dtype=np.float32)
    ref[0, :, 1000:1040] = ref[0, :, 999:1000]            # 41 identical frames: the top-k of their query ties exactly
    lib = A.pack_library(_cuda(ref))
    hm = HostStreamingMatcher(lib, T, 4, 0.0, early=early)
    assert hm.early == early
    for it in range(6):
        chunk = torch.from_numpy(rng.standard_normal((1, D, T), dtype=np.float32))
        if it % 2 == 1:
            chunk[0, :, 5] = torch.from_numpy(ref[0, :, 1010])      # uncertifiable: 41 exact ties
            chunk[0, :, 9] = 0.0                                    # zero query: NaN similarities, exhaustive scan
        out = hm(chunk)
        want, w_idx, _ = A.match_packed(chunk.cuda(), lib, 4, 0.0)
        if it % 2 == 1:
            assert M.last_info.fallback_queries() >= 2
            assert w_idx[0, 5].tolist() == [999, 1000, 1001, 1002]  # ties -> lowest indices (torch.topk on CPU, common.py:105)
        got, ref_out = out.contiguous(), want.transpose(1, 2).cpu()
        assert torch.equal(torch.nan_to_num(got, nan=123.0), torch.nan_to_num(ref_out, nan=123.0))
    # many chunks back to back: every flag value is consumed exactly once
    for it in range(50):
        chunk = torch.from_numpy(rng.standard_normal((1, D, T), dtype=np.float32))
        hm.submit(chunk)
        want, _, _ = A.match_packed(chunk.cuda(), lib, 4, 0.0)
        assert torch.equal(hm.result(), want.transpose(1, 2).cpu())


def test_chunk_graph_with_encoder_and_decoder_modules():
    """realtime_inference.py:143-167 as ONE graph: H2D -> encoder (pre) -> match -> decoder (post) -> D2H, under
    fp16 autocast like the reference's `-fp16`; equals the same modules run eagerly around match_features."""
    torch.manual_seed(5)
    enc = _TinyEncoder().cuda().eval()
    dec = torch.nn.Sequential(torch.nn.Conv1d(768, 32, 3, padding=1), torch.nn.Tanh()).cuda().eval()
    T, N = 24, 3512
    tgt = torch.randn(1, 768, N, device="cuda")
    lib = A.pack_library(tgt)

    def pre(spec):
        with torch.autocast("cuda", dtype=torch.float16):
            return enc(spec)                             # fp16 features, [1, 768, T]

    def post(feat):
        with torch.autocast("cuda", dtype=torch.float16):
            return dec(feat).float()
    hm = HostStreamingMatcher(lib, T, 4, 0.0, pre=pre, post=post, in_shape=(1, 65, T))
    for _ in range(3):
        spec = torch.randn(1, 65, T)
        out = hm(spec)
        with torch.inference_mode(), torch.autocast("cuda", dtype=torch.float16):
            content = enc(spec.cuda())
            matched = A.match_features(content, tgt, k=4, alpha=0.0)
            want = dec(matched).float()
        assert tuple(out.shape) == (1, 32, T)
        assert torch.equal(out, want.cpu())
    # encoder only (the decoder stays outside): the matched frames come back as [B, D, T]
    hm2 = HostStreamingMatcher(lib, T, 4, 0.0, pre=pre, in_shape=(1, 65, T))
    out2 = hm2(spec)
    assert tuple(out2.shape) == (1, 768, T) and torch.equal(out2, matched.cpu())


def test_producer_side_pack_is_bit_identical():
    """RowsContentEncoder: the encoder's last 1x1 conv evaluated channels-last + K1 as its epilogue; the match on the
    packed queries (no K1 inside) equals the match on the same row-major frames, against one and several libraries."""
    from alive_vc_b200.lifecycle import RowsContentEncoder, match_rows, pack_rows
    torch.manual_seed(6)
    enc = _TinyEncoder().cuda().eval()
    prod = RowsContentEncoder(enc)
    spec = torch.randn(2, 65, 150, device="cuda")
    with torch.no_grad():
        rows = prod.rows(spec)                                        # [2, 150, 768] row-major
        assert tuple(rows.shape) == (2, 150, 768) and rows.is_contiguous()
        torch.testing.assert_close(rows, enc(spec).transpose(1, 2), rtol=1e-4, atol=1e-4)   # same layer, other GEMM layout
    q = prod(spec)
    assert q.n == 300 and torch.equal(q.raw.view(2, 150, 768), rows)
    for seed in (1, 2):                                               # one utterance, two speakers' libraries
        lib = pack_rows(torch.randn(20_000, 768, device="cuda", generator=torch.Generator(device="cuda").manual_seed(seed)))
        launches0 = M.launch_count
        out, idx, _ = A.match_packed_queries(q, lib, 4, 0.25, batch=2)
        assert M.launch_count - launches0 == 4                        # search, finish, exact_partial/rows/final - no pack
        w_out, w_idx = match_rows(rows, lib, 4, 0.25, return_indices=True)
        assert torch.equal(idx, w_idx) and torch.equal(out, w_out)
    # per-speaker libraries (items == batch) take packed queries too
    libs = A.pack_libraries(torch.randn(2, 768, 3000, device="cuda"))
    out, idx, _ = A.match_packed_queries(q, libs, 4, 0.0, batch=2)
    w_out, w_idx, _ = M.run_match(rows.transpose(1, 2), libs, 4, 0.0)
    assert torch.equal(idx, w_idx) and torch.equal(out, w_out)


@pytest.mark.parametrize("T,N", [(300, 70_000), (40, 9_000)])
def test_deferred_fallback_equals_one_call(T, N):
    """ALIVE_KNN_MODE_DEFER_FALLBACK + alive_knn_match_fallback (the two halves HostStreamingMatcher uses) give exactly
    what the one-call pipeline gives - with the collect pass (T*N >= 2^24) and without - on a clustered library where
    every query is uncertified, plus a zero query (no usable cut: exhaustive scan)."""
    g = torch.Generator(device="cuda").manual_seed(11)
    cent = torch.randn(768, 30, device="cuda", generator=g)
    ref = (cent[:, torch.randint(0, 30, (N,), device="cuda", generator=g)] +
           0.2 * torch.randn(768, N, device="cuda", generator=g))[None]
    src = (cent[:, torch.randint(0, 30, (T,), device="cuda", generator=g)] +
           0.2 * torch.randn(768, T, device="cuda", generator=g))[None].contiguous()
    src[0, :, 3] = 0.0
    lib = A.pack_library(ref, fmt="bf16")
    out1, idx1, sc1 = M.run_match(src, lib, 4, 0.25, mode="screen")
    n_fb = M.last_info.fallback_queries()
    assert n_fb >= T // 2
    info = {}
    out2, idx2, sc2 = M.run_match(src, lib, 4, 0.25, mode="screen", defer_fallback=True, info_sink=info)
    assert M.last_info.launches == 3 and M.last_info.fallback_queries() == n_fb
    M.run_match_fallback(1, T, lib, 4, 0.25, "screen", 0, M.DEFAULT_R_MAX, info["workspace"], out2, idx2, sc2)
    torch.cuda.synchronize()
    assert torch.equal(idx1, idx2) and torch.equal(sc1.view(torch.int32), sc2.view(torch.int32))
    assert torch.equal(out1.view(torch.int32), out2.view(torch.int32))


def test_host_streaming_matcher_deferred_collect_pass():
    """The deferred fallback chain of the realtime loop on a library large enough for the collect pass (T*N >= 2^24):
    clustered bf16 planes leave every query of a chunk uncertified, result() enqueues refine-prep -> collect search ->
    collect-rescore -> exact scan behind the graph and returns what the one-call pipeline returns."""
    g = torch.Generator(device="cuda").manual_seed(12)
    N, T = 600_000, 32
    cent = torch.randn(768, 200, device="cuda", generator=g)
    ref = (cent[:, torch.randint(0, 200, (N,), device="cuda", generator=g)] +
           0.2 * torch.randn(768, N, device="cuda", generator=g))[None]
    lib = A.pack_library(ref, fmt="bf16")
    del ref
    hm = HostStreamingMatcher(lib, T, 4, 0.0)
    assert hm.early
    for it in range(4):
        if it % 2 == 0:        # a clustered chunk: nothing certifies
            chunk = (cent[:, torch.randint(0, 200, (T,), device="cuda", generator=g)] +
                     0.2 * torch.randn(768, T, device="cuda", generator=g))[None].cpu()
        else:                  # an ordinary chunk: everything certifies, no fallback is enqueued
            chunk = torch.randn(1, 768, T, generator=torch.Generator().manual_seed(it))
        out = hm(chunk)
        want, _, _ = A.match_packed(chunk.cuda(), lib, 4, 0.0)
        fb = M.last_info.fallback_queries()
        assert (fb > 0) == (it % 2 == 0) and M.last_info.collect
        assert torch.equal(out.contiguous(), want.transpose(1, 2).cpu())


@pytest.mark.parametrize("B,T,N", [(1, 3000, 50_000), (3, 500, 8_000)])
def test_host_pipeline_overlapped_copies(B, T, N):
    """lifecycle.HostPipeline: H2D / match / D2H of successive batches overlap on three streams through double-buffered
    device staging; every batch's rows equal match_packed on the same frames (different data per step, so a buffer
    recycled too early would show)."""
    from alive_vc_b200.lifecycle import HostPipeline
    g = torch.Generator(device="cuda").manual_seed(40)
    lib = A.pack_library(torch.randn(1, 768, N, device="cuda", generator=g))
    hp = HostPipeline(lib, B, T, 4, 0.25)
    srcs = [torch.randn(B, 768, T, generator=torch.Generator().manual_seed(s)).pin_memory() for s in range(5)]
    outs = [torch.empty((B, T, 768)).pin_memory() for _ in range(5)]
    for s_h, o_h in zip(srcs, outs):
        hp.step(s_h, o_h)
    hp.drain()
    torch.cuda.synchronize()
    for s_h, o_h in zip(srcs, outs):
        want, _, _ = A.match_packed(s_h.cuda(), lib, 4, 0.25)
        assert torch.equal(o_h, want.cpu())
    with pytest.raises(RuntimeError, match="pinned"):
        hp.step(torch.randn(B, 768, T), outs[0])
    # a caller-supplied match (what the sharded e2e path passes: its scattered match) in the same pipeline
    hp2 = HostPipeline(lib, B, T, match_fn=lambda sd: A.match_packed(sd, lib, 4, 0.25)[0])
    outs2 = [torch.empty((B, T, 768)).pin_memory() for _ in range(5)]
    for s_h, o_h in zip(srcs, outs2):
        hp2.step(s_h, o_h)
    hp2.drain()
    torch.cuda.synchronize()
    for a, b_ in zip(outs, outs2):
        assert torch.equal(a, b_)
