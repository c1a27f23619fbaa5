"""Stage-by-stage GPU diagnostics for the alive_knn kernels (run on the B200 box).

    python tests/gpu_tools/gpu_diag.py            # runs every stage in its own subprocess
    python tests/gpu_tools/gpu_diag.py pack       # one stage in-process

Each stage runs in a fresh process so that a trapped kernel (sticky CUDA error)
cannot poison the following stages.  Output goes to stdout; the driver also writes
gpurun_out/diag_<stage>.log.
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

STAGES = ["pack", "exact", "search1", "search2", "pipeline1", "pipeline2", "golden", "perf1", "perf2"]


def _torch():
    import torch
    return torch


def stage_pack():
    torch = _torch()
    from alive_vc_b200 import matching as M
    g = torch.Generator(device="cuda").manual_seed(1)
    for n in (1, 31, 300, 4097):
        x = torch.randn(768, n, device="cuda", generator=g)
        p = M.pack_frames(x)
        torch.cuda.synchronize()
        assert torch.equal(p.raw, x.t().contiguous()), "raw copy differs"
        nrm = torch.linalg.vector_norm(x.double(), dim=0).float()
        assert torch.allclose(p.norms, nrm, rtol=2e-7, atol=0), (p.norms - nrm).abs().max()
        xn = (x / p.norms[None, :]).t()
        want = xn.bfloat16()
        assert torch.equal(p.packed, want), "packed bf16 differs"
        err = (want.float() - xn).double().norm(dim=1).float()
        assert torch.allclose(p.err, err, rtol=1e-3, atol=1e-8), (p.err - err).abs().max()
        st = p.stats.cpu().numpy().view("uint32")
        import numpy as np
        mx = np.array([st[0]], dtype=np.uint32).view(np.float32)[0]
        assert abs(mx - float(p.err.max())) < 1e-9 and st[1] == 0, (mx, float(p.err.max()), st)
        print(f"pack n={n}: ok  max err-norm {mx:.3e}")
    # row-major input and strided input
    x = torch.randn(500, 768, device="cuda", generator=g)
    p = M.pack_frames(x.t())
    assert torch.equal(p.raw, x)
    big = torch.randn(768, 2000, device="cuda", generator=g)
    p = M.pack_frames(big[:, ::4])
    assert torch.equal(p.raw, big[:, ::4].t().contiguous())
    # zero row -> non-finite count
    x = torch.randn(768, 40, device="cuda", generator=g)
    x[:, 17] = 0
    p = M.pack_frames(x)
    st = p.stats.cpu().numpy().view("uint32")
    assert st[1] == 1, st
    print("pack: strided / row-major / zero-row ok")


def _oracle_case(T, N, k, seed, B=1):
    import numpy as np
    rng = np.random.default_rng(seed)
    src = rng.standard_normal((B, 768, T), dtype=np.float32)
    ref = rng.standard_normal((B, 768, N), dtype=np.float32)
    return src, ref


def _check_against_oracle(idx_gpu, src, ref, k, label):
    import numpy as np
    from oracle import knn_oracle as O
    scores = O.cosine_scores_np(src, ref)
    _, idx_o = O.topk_desc_np(scores, k)
    ok, n_exact, n_tie, bad = O.indices_match_mod_ties(idx_gpu, idx_o, scores, 1e-6)
    print(f"{label}: rows exact {n_exact}, tie-excused {n_tie}, ok={ok}")
    if not ok:
        print("  first bad:", bad)
    return ok


def stage_exact():
    torch = _torch()
    from alive_vc_b200 import matching as M
    for (T, N, k) in [(50, 300, 4), (7, 64, 4), (40, 1000, 8), (10, 700, 16), (5, 4, 4), (33, 20000, 4)]:
        src, ref = _oracle_case(T, N, k, 100 + T)
        q = M.pack_queries(torch.from_numpy(src).cuda())
        lib = M.pack_library(torch.from_numpy(ref).cuda())
        sc, idx = M.search_topk(q, lib, k, mode="exact")
        torch.cuda.synchronize()
        assert _check_against_oracle(idx.cpu().numpy()[None], src, ref, k, f"exact T={T} N={N} k={k}")


def _check_lists(variant, T, N, seed):
    """K2 in isolation: compare the screened lists with a torch matmul of the same bf16 operands."""
    import ctypes
    torch = _torch()
    from alive_vc_b200 import _cabi, matching as M
    g = torch.Generator(device="cuda").manual_seed(seed)
    q = M.pack_frames(torch.randn(768, T, device="cuda", generator=g))
    lib = M.pack_frames(torch.randn(768, N, device="cuda", generator=g))
    plan = M.make_plan(T, N, 768, q.device, variant)
    print(f"variant {variant} T={T} N={N} plan={plan.as_dict()}")
    cs = torch.full((T, plan.lists, 8), float("nan"), device="cuda")
    ci = torch.full((T, plan.lists, 8), -7, dtype=torch.int32, device="cuda")
    rc = _cabi.load().alive_knn_search(q.packed.data_ptr(), lib.packed.data_ptr(), ctypes.byref(plan),
                                       cs.data_ptr(), ci.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _cabi.check(rc, "search")
    torch.cuda.synchronize()
    ref = q.packed.float() @ lib.packed.float().t()          # [T,N] fp32 of the same bf16 operands
    tile_n = 256
    bad = 0
    worst = 0.0
    for seg in range(plan.segments):
        t0 = seg * plan.tiles_per_segment
        t1 = min(t0 + plan.tiles_per_segment, plan.n_tiles)
        for half in range(2):
            cols = torch.cat([torch.arange(t * tile_n + half * 128, min(t * tile_n + half * 128 + 128, N), device="cuda")
                              for t in range(t0, t1) if t * tile_n + half * 128 < N] or
                             [torch.empty(0, dtype=torch.long, device="cuda")])
            lst = seg * 2 + half
            s_k, i_k = cs[:, lst, :], ci[:, lst, :].long()
            if cols.numel() == 0:
                assert (i_k == -1).all()
                continue
            kk = min(8, cols.numel())
            want_s, want_pos = ref[:, cols].topk(kk, dim=1)
            got_s = s_k[:, :kk]
            d = (got_s - want_s).abs().max().item()
            worst = max(worst, d)
            if d > 2e-4:
                bad += 1
                if bad <= 3:
                    print(f"  list {lst}: score mismatch {d:.3e}\n   got {got_s[0].tolist()}\n   want {want_s[0].tolist()}")
            # indices must point at entries with those scores
            valid = i_k[:, :kk] >= 0
            if not valid.all():
                bad += 1
                if bad <= 3:
                    print(f"  list {lst}: invalid idx", i_k[0].tolist())
                continue
            at = ref.gather(1, i_k[:, :kk])
            d2 = (at - got_s).abs().max().item()
            if d2 > 2e-4:
                bad += 1
                if bad <= 3:
                    print(f"  list {lst}: idx/score inconsistent {d2:.3e}")
            if kk < 8:
                assert (i_k[:, kk:] == -1).all() and torch.isinf(s_k[:, kk:]).all()
    print(f"variant {variant} T={T} N={N}: lists checked, bad={bad}, worst |score diff|={worst:.3e}")
    return bad == 0


def stage_search1():
    ok = True
    for (T, N) in [(128, 256), (200, 5000), (1, 300), (129, 4097), (300, 70000)]:
        ok &= _check_lists(1, T, N, T + N)
    assert ok


def stage_search2():
    ok = True
    for (T, N) in [(256, 256), (200, 5000), (1, 300), (129, 4097), (300, 70000)]:
        ok &= _check_lists(2, T, N, T + N)
    assert ok


def _pipeline(variant):
    torch = _torch()
    import numpy as np
    from alive_vc_b200 import matching as M
    from oracle import knn_oracle as O
    ok = True
    for (B, T, N, k, alpha) in [(1, 50, 3000, 4, 0.0), (1, 200, 20000, 4, 0.0), (2, 33, 1517, 4, 0.25),
                                (1, 96, 50000, 8, 0.0), (1, 1, 2049, 1, 0.0)]:
        src, ref = _oracle_case(T, N, k, 7 * T + N, B)
        out, idx = M.match_features(torch.from_numpy(src).cuda(), torch.from_numpy(ref).cuda(), k, alpha,
                                    return_indices=True, mode="screen", variant=variant)
        torch.cuda.synchronize()
        info = M.last_info
        fb = info.fallback_queries()
        seln = info.sel_n.cpu().numpy()
        ok &= _check_against_oracle(idx.cpu().numpy(), src, ref, k, f"pipeline v{variant} B={B} T={T} N={N} k={k}")
        out_o, idx_o, _ = O.match_features_np(src, ref, k, alpha, True)
        same = (idx.cpu().numpy() == idx_o).all(axis=2)
        o = out.cpu().numpy()
        exact_rows = np.array_equal(np.swapaxes(o, 1, 2)[same], np.swapaxes(out_o, 1, 2)[same])
        close = np.allclose(o, out_o, rtol=1e-5, atol=1e-6)
        print(f"   features: bit-exact on index-identical rows={exact_rows}, allclose={close}; "
              f"fallback queries={fb}, survivors mean={seln[seln >= 0].mean() if (seln >= 0).any() else -1:.1f} "
              f"max={seln.max()}, out strides={tuple(out.stride())}")
        ok &= exact_rows and close
    assert ok


def stage_pipeline1():
    _pipeline(1)


def stage_pipeline2():
    _pipeline(2)


def stage_golden():
    torch = _torch()
    import numpy as np
    import alive_vc_b200 as A
    from oracle import knn_oracle as O
    from oracle.gen_golden import CASES, GOLDEN_DIR, make_case_inputs
    ok = True
    for name, spec in sorted(CASES.items()):
        g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        src, ref = make_case_inputs(spec)
        s = torch.from_numpy(np.ascontiguousarray(src)).cuda()
        if spec["kind"] == "vl":
            vl = A.VoiceLibrary(num_tokens=spec["N"]).cuda()
            with torch.no_grad():
                vl.tokens.copy_(torch.from_numpy(ref).cuda())
            s.requires_grad_(True)
            out, idx = vl.match(s, k=spec["k"], alpha=spec["alpha"], return_indices=True)
            gout = torch.from_numpy(np.random.default_rng(spec["seed"] + 1000)
                                    .standard_normal(g["out"].shape, dtype=np.float32)).cuda()
            out.backward(gout)
            gt_ok = np.allclose(vl.tokens.grad.cpu().numpy(), g["grad_tokens"], rtol=1e-5, atol=1e-5)
            gs_ok = np.allclose(s.grad.cpu().numpy(), g["grad_source"], rtol=1e-6, atol=1e-7)
            ref_b = np.broadcast_to(ref, (src.shape[0],) + ref.shape[1:])
        else:
            r = torch.from_numpy(np.ascontiguousarray(ref)).cuda()
            if spec["kind"] == "mf_strided":
                big = torch.zeros((ref.shape[0], ref.shape[1], ref.shape[2] * 4), device="cuda")
                big[:, :, ::4] = r
                r = big[:, :, ::4]
            out, idx = A.match_features(s, r, spec["k"], spec["alpha"], return_indices=True)
            gt_ok = gs_ok = True
            ref_b = ref
        scores = O.cosine_scores_np(src, ref_b)
        i_ok, n_exact, n_tie, bad = O.indices_match_mod_ties(idx.cpu().numpy(), g["indices"].astype(np.int64), scores, 1e-6)
        o = out.detach().cpu().numpy()
        same = (idx.cpu().numpy() == g["indices"]).all(axis=2)
        bit = np.array_equal(np.swapaxes(o, 1, 2)[same], np.swapaxes(g["out"], 1, 2)[same])
        close = name in ("mf_dupes", "mf_clustered") or np.allclose(o, g["out"], rtol=1e-5, atol=1e-6)   # tie rows differ legitimately
        st_ok = tuple(out.stride()) == tuple(g["out_strides"]) or o.shape[2] == 1
        print(f"{name}: idx ok={i_ok} (exact rows {n_exact}, ties {n_tie}) bit-exact={bit} close={close} "
              f"strides ok={st_ok} grads ok={gt_ok and gs_ok}")
        if not i_ok:
            print("   ", bad)
        ok &= i_ok and bit and close and st_ok and gt_ok and gs_ok
    assert ok


def _perf(variant):
    import ctypes
    torch = _torch()
    from alive_vc_b200 import _cabi, matching as M
    g = torch.Generator(device="cuda").manual_seed(3)
    for (T, N) in [(1000, 100000), (32, 200000), (8192, 400000), (16384, 1000000)]:
        lib = M.alloc_packed(N, 768, "cuda")
        step = 100000
        for r0 in range(0, N, step):
            n = min(step, N - r0)
            M.pack_into(lib, r0, torch.randn(768, n, device="cuda", generator=g))
        q = M.pack_frames(torch.randn(768, T, device="cuda", generator=g))
        plan = M.make_plan(T, N, 768, q.device, variant)
        cs = torch.empty((T, plan.lists, 8), device="cuda")
        ci = torch.empty((T, plan.lists, 8), dtype=torch.int32, device="cuda")
        c = _cabi.load()
        st = torch.cuda.current_stream().cuda_stream

        def run():
            rc = c.alive_knn_search(q.packed.data_ptr(), lib.packed.data_ptr(), ctypes.byref(plan),
                                    cs.data_ptr(), ci.data_ptr(), st)
            _cabi.check(rc, "search")
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10 if T * N < 5e9 else 3
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tf = 2.0 * T * N * 768 / (ms * 1e-3) / 1e12
        print(f"perf v{variant} T={T} N={N}: search {ms:.3f} ms  {tf:.1f} TFLOP/s  plan seg={plan.segments} tps={plan.tiles_per_segment} grid={plan.grid}")
        # whole pipeline
        src = torch.randn(1, 768, T, device="cuda", generator=g)
        for _ in range(2):
            M.match_packed(src, lib, 4, 0.0, "screen", variant)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            M.match_packed(src, lib, 4, 0.0, "screen", variant)
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / reps
        info = M.last_info
        seln = info.sel_n
        print(f"      full match {ms2:.3f} ms ({T / (ms2 * 1e-3):.3e} qframes/s), fallback={info.fallback_queries()}, "
              f"survivors mean={seln[seln >= 0].float().mean().item():.1f} max={seln.max().item()}")
        del lib, q, cs, ci
        torch.cuda.empty_cache()


def stage_perf1():
    _perf(1)


def stage_perf2():
    _perf(2)


def main():
    if len(sys.argv) > 1 and sys.argv[1] != "all":
        for s in sys.argv[1:]:
            globals()["stage_" + s]()
            print(f"[stage {s}] PASS")
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    summary = []
    for s in STAGES:
        t0 = time.time()
        try:
            res = subprocess.run([sys.executable, os.path.abspath(__file__), s], capture_output=True, text=True,
                                 timeout=420, cwd=ROOT)
            out, rc = res.stdout + res.stderr, res.returncode
        except subprocess.TimeoutExpired as e:
            out, rc = (e.stdout or b"").decode() + (e.stderr or b"").decode() + "\nTIMEOUT", -9
        with open(os.path.join(ROOT, "gpurun_out", f"diag_{s}.log"), "w") as f:
            f.write(out)
        tail = "\n".join(out.strip().splitlines()[-40:])
        print(f"===== stage {s}: rc={rc} ({time.time() - t0:.1f}s) =====\n{tail}\n", flush=True)
        summary.append((s, rc))
    print("SUMMARY", summary)


if __name__ == "__main__":
    main()
