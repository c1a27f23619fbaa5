"""Oracle parity at the FULL sizes of BASELINE.json configs[2..4] (cfg3, cfg4, cfg5).

The numpy oracle cannot run these sizes; the independent C restatement (oracle/knn_oracle.c, OpenMP, pinned to the
golden vectors by tests/test_oracle.py) can, on a slice of the query rows: the GPU matches the WHOLE batch in one
call exactly as bench.py does, then a random sample of rows is recomputed on the host against the whole library
(module/common.py:96-109 step by step) and compared - indices bit-exact except where fp32 similarities tie within
1e-6, features bit-exact on rows with equal indices, within 1e-5 relative otherwise.  The libraries are the bench's
own synthetic ones (bench.build_library); the host copy is the raw fp32 block the pack kernel stored, which equals
the generated frames bit for bit (test_pack_kernel_layout_and_stats)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import alive_vc_b200 as A                                          # noqa: E402
from alive_vc_b200 import matching as M                             # noqa: E402
from oracle import c_oracle                                         # noqa: E402

TIE_TOL = 1e-6
FEAT_RTOL = 1e-5


def _check_rows(idx_gpu, out_gpu, q_rows, lib_rows_np, k, alpha, o_idx=None, o_val=None, o_out=None):
    """idx_gpu [R,k], out_gpu [R,D] (torch, cuda) against the C oracle on q_rows [R,D] (numpy) x lib_rows_np [N,D]."""
    if o_idx is None:
        o_out, o_idx, o_val = c_oracle.match_features_c(q_rows[None], lib_rows_np[None], k, alpha, True, rows=True)
        o_out, o_idx, o_val = o_out[0], o_idx[0], o_val[0]
    g_idx = idx_gpu.cpu().numpy()
    same = (g_idx == o_idx).all(axis=1)
    n_tie = 0
    for r in np.nonzero(~same)[0]:
        # an index difference is only excused when the similarities at every position tie within 1e-6: rescore our
        # frames exactly like the oracle scores its own (float32-normalised frames, double accumulation)
        qn = (q_rows[r] / np.float32(np.sqrt(np.sum(q_rows[r].astype(np.float64) ** 2)))).astype(np.float32)
        ours = []
        for j in g_idx[r]:
            f = lib_rows_np[j]
            fn = (f / np.float32(np.sqrt(np.sum(f.astype(np.float64) ** 2)))).astype(np.float32)
            ours.append(np.float32(np.dot(qn.astype(np.float64), fn.astype(np.float64))))
        assert np.all(np.abs(np.array(ours, np.float32) - o_val[r]) <= TIE_TOL), (r, g_idx[r], o_idx[r], ours, o_val[r])
        n_tie += 1
    o = out_gpu.cpu().numpy()
    assert np.array_equal(o[same], o_out[same]), "features not bit-exact on rows with identical indices"
    for r in np.nonzero(~same)[0]:
        # rows excused by a tie: the features must still be the sequential mean of OUR frames (common.py:107)
        acc = lib_rows_np[g_idx[r][0]].copy()
        for j in g_idx[r][1:]:
            acc = (acc + lib_rows_np[j]).astype(np.float32)
        mine = ((acc / np.float32(k)).astype(np.float32) * np.float32(1 - alpha)).astype(np.float32) + \
            (q_rows[r] * np.float32(alpha)).astype(np.float32)
        np.testing.assert_allclose(o[r], mine.astype(np.float32), rtol=FEAT_RTOL, atol=1e-6)
    return int(same.sum()), n_tie


def test_cfg3_full_size_against_c_oracle():
    """BASELINE configs[2]: T = 100,000 query frames vs N = 1,000,000 library frames, one launch; 512 rows checked."""
    import bench
    dev = torch.device("cuda", 0)
    torch.cuda.empty_cache()
    T, N, k = 100_000, 1_000_000, 4
    lib = bench.build_library(0, N, 31, dev)
    g = torch.Generator(device=dev).manual_seed(32)
    src = torch.randn(1, 768, T, device=dev, generator=g)
    out, idx, _ = A.match_packed(src, lib, k, 0.0, mode="screen")
    assert M.last_info.mode == "screen" and M.last_info.exact_scan_queries() == 0
    rows = torch.randperm(T, device=dev, generator=g)[:512].sort().values
    q_rows = src[0][:, rows].t().contiguous().cpu().numpy()
    n_exact, n_tie = _check_rows(idx[0][rows], out[0][rows], q_rows, lib.raw.cpu().numpy(), k, 0.0)
    assert n_exact + n_tie == 512 and n_exact >= 500
    del lib, src, out, idx
    torch.cuda.empty_cache()


def test_cfg4_full_size_against_c_oracle_and_exhaustive_scan():
    """BASELINE configs[3] library size on one GPU: T = 10,000 vs N = 10,000,000, one launch; 64 rows checked against
    the C oracle (run on four 2.5M-frame slices of the library and merged, to bound host memory), and 192 rows
    against the exhaustive fp64 scan on the GPU."""
    import bench
    dev = torch.device("cuda", 0)
    torch.cuda.empty_cache()
    T, N, k = 10_000, 10_000_000, 4
    lib = bench.build_library(0, N, 3, dev)
    g = torch.Generator(device=dev).manual_seed(5)
    src = torch.randn(1, 768, T, device=dev, generator=g)
    out, idx, sc = A.match_packed(src, lib, k, 0.0, mode="screen")
    assert M.last_info.fallback_queries() == 0
    # (a) screen == exhaustive scan, bit for bit
    sub = src[:, :, :192].contiguous()
    out_e, idx_e, sc_e = M.run_match(sub, lib, k, 0.0, mode="exact")
    assert torch.equal(idx[:, :192], idx_e) and torch.equal(sc[:, :192], sc_e) and torch.equal(out[:, :192], out_e)
    # (b) C oracle: per-slice top-k (global indices), merged under the same total order
    rows = torch.randperm(T, device=dev, generator=g)[:64].sort().values
    q_rows = src[0][:, rows].t().contiguous().cpu().numpy()
    cand_i, cand_v = [], []
    for lo in range(0, N, 2_500_000):
        part = lib.raw[lo:lo + 2_500_000].cpu().numpy()
        _, p_idx, p_val = c_oracle.match_features_c(q_rows[None], part[None], k, 0.0, True, rows=True)
        cand_i.append(p_idx[0] + lo)
        cand_v.append(p_val[0])
        del part
    cand_i, cand_v = np.concatenate(cand_i, axis=1), np.concatenate(cand_v, axis=1)       # [64, 4*k]
    order = np.lexsort((cand_i, -cand_v), axis=1)[:, :k]                                   # score desc, index asc
    o_idx = np.take_along_axis(cand_i, order, axis=1)
    o_val = np.take_along_axis(cand_v, order, axis=1)
    picked = lib.raw[torch.from_numpy(o_idx).to(dev)]                                      # [64, k, D] raw rows
    acc = picked[:, 0].clone()
    for j in range(1, k):
        acc = acc + picked[:, j]
    o_out = (acc / k).cpu().numpy()                                                        # common.py:107 (alpha = 0)

    class _Rows:                                                                           # rows fetched on demand
        def __getitem__(self, j):
            return lib.raw[int(j)].cpu().numpy()
    n_exact, n_tie = _check_rows(idx[0][rows], out[0][rows], q_rows, _Rows(), k, 0.0, o_idx, o_val, o_out)
    assert n_exact + n_tie == 64 and n_exact >= 60
    del lib, src, out, idx
    torch.cuda.empty_cache()


def test_cfg5_full_size_against_c_oracle():
    """BASELINE configs[4]: 64 utterances x 1000 frames vs 64 per-speaker 500,000-frame libraries, ONE batched
    pipeline launch (147 GB of packed libraries); 64 rows of 8 of the speakers checked against the C oracle."""
    dev = torch.device("cuda", 0)
    torch.cuda.empty_cache()
    free, _total = torch.cuda.mem_get_info(dev)
    B, T, N, k = 64, 1000, 500_000, 4
    need = B * N * (768 * 6 + 16) + (8 << 30)
    if free < need:
        pytest.skip(f"cfg5 needs {need / 1e9:.0f} GB of free HBM, {free / 1e9:.0f} GB available")
    lib = M.alloc_packed(B * N, 768, dev, items=B)
    for b in range(B):
        for c0 in range(0, N, 250_000):
            gg = torch.Generator(device=dev).manual_seed((91 + 17 * (b + 1)) * 1_000_003 + c0 // 250_000)
            x = torch.randn(768, 250_000, device=dev, generator=gg)
            M.pack_into(lib, b * N + c0, x)
            del x
    g = torch.Generator(device=dev).manual_seed(92)
    src = torch.randn(B, 768, T, device=dev, generator=g)
    out, idx, _ = A.match_packed(src, lib, k, 0.0, mode="screen")            # idx relative to the item's own library
    assert M.last_info.mode == "screen" and M.last_info.launches == 7     # ONE pipeline: pack, search, finish, collect pass (2, idle), exact (3, idle)
    total_exact = total_tie = 0
    for b in (0, 9, 18, 27, 36, 45, 54, 63):
        rows = torch.randperm(T, device=dev, generator=g)[:64].sort().values
        q_rows = src[b][:, rows].t().contiguous().cpu().numpy()
        lib_rows = lib.raw[b * N:(b + 1) * N].cpu().numpy()
        n_exact, n_tie = _check_rows(idx[b][rows], out[b][rows], q_rows, lib_rows, k, 0.0)
        total_exact += n_exact
        total_tie += n_tie
        del lib_rows
    assert total_exact + total_tie == 8 * 64 and total_exact >= 8 * 60
    del lib, src, out, idx
    torch.cuda.empty_cache()
