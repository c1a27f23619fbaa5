"""Small end-to-end invocations for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  python tests/gpu_tools/gpu_sanitize.py
    compute-sanitizer --tool racecheck python tests/gpu_tools/gpu_sanitize.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import alive_vc_b200 as A                       # noqa: E402
from alive_vc_b200 import matching as M        # noqa: E402
from oracle import knn_oracle as O             # noqa: E402


def main():
    rng = np.random.default_rng(3)
    for (B, T, N, k, alpha, mode, variant) in [(1, 40, 700, 4, 0.0, "screen", 1), (1, 300, 2100, 4, 0.25, "screen", 2),
                                               (2, 9, 257, 2, 0.0, "screen", 0), (1, 12, 90, 16, 0.0, "exact", 0),
                                               (1, 5, 600, 4, 0.0, "exact", 0)]:
        src = rng.standard_normal((B, 768, T), dtype=np.float32)
        ref = rng.standard_normal((B, 768, N), dtype=np.float32)
        out, idx = A.match_features(torch.from_numpy(src).cuda(), torch.from_numpy(ref).cuda(), k, alpha,
                                    return_indices=True, mode=mode, variant=variant)
        torch.cuda.synchronize()
        out_o, idx_o, _ = O.match_features_np(src, ref, k, alpha, True)
        assert np.array_equal(idx.cpu().numpy(), idx_o), (T, N)
        assert np.array_equal(out.cpu().numpy(), out_o), (T, N)
        print("ok", B, T, N, k, alpha, mode, variant, flush=True)
    # duplicates -> certificate fails -> exact scan with gather; autograd scatter
    base = rng.standard_normal((1, 768, 200), dtype=np.float32)
    ref = np.concatenate([base] * 3, axis=2)
    ref[:, :, 77] = 0          # a zero frame: every query goes to the exact scan (NaN ranks first)
    out, idx = A.match_features(torch.from_numpy(base[:, :, :10].copy()).cuda(), torch.from_numpy(ref).cuda(), 3, 0.0,
                                return_indices=True, mode="screen")
    torch.cuda.synchronize()
    # one realtime chunk: the skinny kernel (operands swapped, resident query chunk)
    src = rng.standard_normal((1, 768, 24), dtype=np.float32)
    ref = rng.standard_normal((1, 768, 3512), dtype=np.float32)
    out, idx = A.match_features(torch.from_numpy(src).cuda(), torch.from_numpy(ref).cuda(), 4, 0.0, return_indices=True)
    torch.cuda.synchronize()
    assert np.array_equal(idx.cpu().numpy(), O.match_features_np(src, ref, 4, 0.0, True)[1])
    print("ok skinny chunk", flush=True)
    # tight clusters on a library big enough for the collect pass (second tensor-core pass + rescoring)
    g = torch.Generator(device="cuda").manual_seed(1)
    cent = torch.randn(768, 40, device="cuda", generator=g)
    refc = (cent[:, torch.randint(0, 40, (70_000,), device="cuda", generator=g)] +
            0.2 * torch.randn(768, 70_000, device="cuda", generator=g))[None]
    srcc = (cent[:, torch.randint(0, 40, (260,), device="cuda", generator=g)] +
            0.2 * torch.randn(768, 260, device="cuda", generator=g))[None]
    libc = A.pack_library(refc, fmt="bf16")            # (the automatic format would certify it in fp16)
    _, idx_s, _ = M.run_match(srcc, libc, 4, 0.0, mode="screen")
    assert M.last_info.collect and M.last_info.fallback_queries() > 0
    torch.cuda.synchronize()
    print("ok collect pass", M.last_info.fallback_queries(), M.last_info.exact_scan_queries(), flush=True)
    # every K1 variant: generic 8/32-frame tiles (n % 4 != 0), the 16-byte cp.async channel-major kernel with a
    # ragged last tile, the register-resident row-major kernel with a ragged last CTA, and one launch for a
    # batch of query items (uniform and non-uniform item strides)
    from alive_vc_b200.lifecycle import match_rows
    for n, row_major in ((8201, False), (8300, False), (9001, False), (8203, True), (700, True)):
        x = torch.randn(768, n, device="cuda", generator=g)
        if row_major:
            x = x.t().contiguous().t()
        p = M.pack_frames(x)
        torch.cuda.synchronize()
        assert torch.equal(p.raw, x.t().contiguous())
        assert torch.equal(p.packed, (x / p.norms[None, :]).t().bfloat16())
    tall = torch.randn(3, 400, 768, device="cuda", generator=g)
    lrows = torch.randn(900, 768, device="cuda", generator=g)
    o1 = match_rows(tall[:, :200], lrows)                     # non-uniform item stride, 600 frames: pack_rm_kernel
    o2 = match_rows(tall[:, :200].contiguous(), lrows)        # uniform
    torch.cuda.synchronize()
    assert torch.equal(o1, o2)
    srcb = torch.randn(3, 768, 3000, device="cuda", generator=g)   # 9000 channel-major query frames in one launch
    ob, ib = A.match_features(srcb, lrows.t().contiguous()[None].expand(3, 768, 900), 4, 0.0, return_indices=True)
    o0, i0 = A.match_features(srcb[1:2], lrows.t().contiguous()[None], 4, 0.0, return_indices=True)
    torch.cuda.synchronize()
    assert torch.equal(ib[1:2], i0) and torch.equal(ob[1:2], o0)
    print("ok pack variants", flush=True)
    # round 2: the resident-query CTA-pair kernel (d = 768, >= 2 tiles per unit), both plane formats
    for fmt in ("bf16", "fp16"):
        refr = torch.randn(1, 768, 40_000, device="cuda", generator=g)
        srcr = torch.randn(1, 768, 300, device="cuda", generator=g)
        libr = A.pack_library(refr, fmt=fmt)
        _, idx_s, _ = M.run_match(srcr, libr, 4, 0.0, mode="screen")
        _, idx_e, _ = M.run_match(srcr, libr, 4, 0.0, mode="exact")
        torch.cuda.synchronize()
        assert torch.equal(idx_s, idx_e)
    print("ok resident kernel", flush=True)
    # one-plane row-major pack (the pipelined kernel), K4 warp kernel with and without the blend row
    xr = torch.randn(9000, 768, device="cuda", generator=g)
    p1 = M.pack_frames(xr.t(), refine=False, fmt="bf16")
    p2 = M.pack_frames(xr.t(), refine=True, fmt="bf16")
    torch.cuda.synchronize()
    assert torch.equal(p1.packed, p2.packed) and torch.equal(p1.norms, p2.norms) and p1.lo is None
    for al in (0.0, 0.3):
        o_a, i_a, _ = A.match_packed(srcr, p2, 4, al)
        q_pf = M.pack_queries(srcr)
        o_b = torch.empty((300, 768), device="cuda")
        M.gather_mean(p2, i_a.view(300, 4), q_pf, al, o_b)
        torch.cuda.synchronize()
        assert torch.equal(o_b, o_a[0])
    print("ok one-plane pack + gather", flush=True)
    # sharded: two in-process ranks, records all-gathered through shared memory, fused merge + peer gather
    import threading
    from alive_vc_b200.sharded import ShardedLibrary, ThreadComm, shard_bounds
    comms = ThreadComm.make(2)
    refs = torch.randn(1, 768, 20_000, device="cuda", generator=g)
    srcs = torch.randn(1, 768, 150, device="cuda", generator=g)
    want, widx = A.match_features(srcs, refs, 4, 0.0, return_indices=True)
    res = [None, None]

    def rank_fn(r):
        torch.cuda.set_device(0)
        sh = ShardedLibrary.from_full(refs, comm=comms[r])
        o, i = sh.match(srcs, 4, 0.0, return_indices=True)
        res[r] = (o.clone(), i.clone())
        torch.cuda.synchronize()
        comms[r].barrier()          # a rank's shard must outlive every peer's gather
    ths = [threading.Thread(target=rank_fn, args=(r,)) for r in range(2)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    for r in range(2):
        assert torch.equal(res[r][1], widx) and torch.equal(res[r][0], want)
    print("ok sharded thread ranks", flush=True)
    # realtime loop with host buffers: zero-copy result + early notification
    from alive_vc_b200.lifecycle import HostStreamingMatcher
    hm = HostStreamingMatcher(A.pack_library(refs), 24)
    for _ in range(3):
        ch = torch.randn(1, 768, 24)
        got = hm(ch)
        wantc, _, _ = A.match_packed(ch.cuda(), hm.lib, 4, 0.0)
        assert torch.equal(got, wantc.transpose(1, 2).cpu())
    print("ok host streaming (early notify)", flush=True)
    vl = A.VoiceLibrary(num_tokens=300).cuda()
    s = torch.randn(2, 768, 11, device="cuda", requires_grad=True)
    vl.match(s, alpha=0.5).sum().backward()
    torch.cuda.synchronize()
    print("ok fallback + autograd", flush=True)


if __name__ == "__main__":
    main()
