"""How the screened path behaves away from i.i.d. random data (run on the B200 box):
fallback counts and timings for clustered / low-rank libraries, and the exact-scan speed."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import alive_vc_b200 as A                       # noqa: E402
from alive_vc_b200 import matching as M        # noqa: E402


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(5)
    print("== exact scan speed (mode=exact)")
    for (T, N) in [(24, 3512), (450, 3512), (1000, 100_000), (32, 200_000)]:
        src = torch.randn(1, 768, T, device=dev, generator=g)
        lib = A.pack_library(torch.randn(1, 768, N, device=dev, generator=g))
        ms = timed(lambda: M.run_match(src, lib, 4, 0.0, mode="exact"))
        ms2 = timed(lambda: M.run_match(src, lib, 4, 0.0, mode="screen")) if N >= 256 else float("nan")
        print(f"T={T} N={N}: exact {ms:.3f} ms ({2*T*N*768/ms/1e9:.2f} TFLOP/s-equivalent), screen {ms2:.3f} ms")
    print("== structured libraries (T=1000, N=100k, k=4)")
    T, N = 1000, 100_000
    for name, noise, nclus in [("iid", None, 0), ("clusters noise=1.0", 1.0, 100), ("clusters noise=0.5", 0.5, 100),
                               ("clusters noise=0.2", 0.2, 100), ("clusters noise=0.05", 0.05, 100),
                               ("lowrank32", -1, 0)]:
        if noise is None:
            ref = torch.randn(1, 768, N, device=dev, generator=g)
            src = torch.randn(1, 768, T, device=dev, generator=g)
        elif noise < 0:
            W = torch.randn(768, 32, device=dev, generator=g)
            ref = (W @ torch.randn(32, N, device=dev, generator=g) + 0.3 * torch.randn(768, N, device=dev, generator=g))[None]
            src = (W @ torch.randn(32, T, device=dev, generator=g) + 0.3 * torch.randn(768, T, device=dev, generator=g))[None]
        else:
            cent = torch.randn(768, nclus, device=dev, generator=g)
            ref = (cent[:, torch.randint(0, nclus, (N,), device=dev, generator=g)] +
                   noise * torch.randn(768, N, device=dev, generator=g))[None]
            src = (cent[:, torch.randint(0, nclus, (T,), device=dev, generator=g)] +
                   noise * torch.randn(768, T, device=dev, generator=g))[None]
        lib = A.pack_library(ref)
        ms = timed(lambda: M.run_match(src, lib, 4, 0.0, mode="screen"))
        info = M.last_info
        seln = info.sel_n
        fb = info.fallback_queries()
        ok = seln >= 0
        # cross-check against the exhaustive scan
        _, idx_s, _ = M.run_match(src, lib, 4, 0.0, mode="screen")
        _, idx_e, _ = M.run_match(src, lib, 4, 0.0, mode="exact")
        same = bool(torch.equal(idx_s, idx_e))
        print(f"{name:22s}: {ms:8.3f} ms  uncertified {fb:5d}/{T} (exhaustive scan {info.exact_scan_queries():4d})  survivors mean "
              f"{seln[ok].float().mean().item() if ok.any() else float('nan'):6.1f} max {seln.max().item():4d}  "
              f"screen==exact: {same}")


def sparse_fallback():
    """a few uncertifiable queries inside a large, otherwise well-behaved batch"""
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(11)
    print("== sparse fallbacks (T=1000, N=1M iid + one tight cluster of 300 frames; 5 queries sit in the cluster)")
    T, N = 1000, 1_000_000
    ref = torch.randn(768, N, device=dev, generator=g)
    c = torch.randn(768, 1, device=dev, generator=g)
    where = torch.randperm(N, device=dev, generator=g)[:300]
    ref[:, where] = c + 0.01 * torch.randn(768, 300, device=dev, generator=g)
    src = torch.randn(1, 768, T, device=dev, generator=g)
    lib = A.pack_library(ref[None])
    ms0 = timed(lambda: M.run_match(src, lib, 4, 0.0, mode="screen"))
    fb0 = M.last_info.fallback_queries()
    src[0, :, :5] = c + 0.01 * torch.randn(768, 5, device=dev, generator=g)
    ms1 = timed(lambda: M.run_match(src, lib, 4, 0.0, mode="screen"))
    fb1 = M.last_info.fallback_queries()
    ex1 = M.last_info.exact_scan_queries()
    _, idx_s, _ = M.run_match(src, lib, 4, 0.0, mode="screen")
    _, idx_e, _ = M.run_match(src[:, :, :16].contiguous(), lib, 4, 0.0, mode="exact")
    print(f"no cluster queries: {ms0:.3f} ms (fallback {fb0});  5 cluster queries: {ms1:.3f} ms (uncertified {fb1}, exhaustive scan {ex1}); "
          f"first 16 queries screen==exact: {bool(torch.equal(idx_s[:, :16], idx_e))}")


if __name__ == "__main__":
    main()
    sparse_fallback()
