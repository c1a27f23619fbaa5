"""K4 standalone: the warp-per-query gather kernel against the CTA-per-query one, the query-row skip, grid sizes."""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run():
    from alive_vc_b200 import _cabi, matching as M
    c = _cabi.load()
    dev = "cuda"
    D, K = 768, 4
    N = int(os.environ.get("K4_N", 10_000_000))
    raw = torch.randn(N, D, device=dev)
    peak = 6548.5
    for T in (10_000, 100_000):
        g = torch.Generator(device=dev).manual_seed(T)
        idx = torch.randint(0, N, (T, K), device=dev, generator=g)
        q_raw = torch.randn(T, D, device=dev, generator=g)
        q_norm = torch.linalg.vector_norm(q_raw, dim=1)
        out = torch.empty((T, D), device=dev)
        for label, qn, alpha in (("skip-q", q_norm, 0.0), ("read-q", None, 0.0)):
            def call():
                rc = c.alive_knn_gather_mean(raw.data_ptr(), N, D, idx.data_ptr(), T, K, q_raw.data_ptr(),
                                             qn.data_ptr() if qn is not None else None, alpha, out.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream)
                _cabi.check(rc, "gather")
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                call()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            alg = T * (K * D * 4 + D * 4)
            print(f"T={T:6d} {label}: {ms * 1e3:8.1f} us  {alg / ms / 1e6:7.1f} GB/s algorithmic = {alg / ms / 1e6 / peak:.3f} of peak "
                  f"(env {os.environ.get('ALIVE_KNN_GATHER_CTA', '0')}/{os.environ.get('ALIVE_KNN_GATHER_CTAS_PER_SM', 'dflt')})", flush=True)
        sel = torch.empty((T * K, D), device=dev)
        flat = idx.reshape(-1)
        for _ in range(3):
            torch.index_select(raw, 0, flat, out=sel)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            torch.index_select(raw, 0, flat, out=sel)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"T={T:6d} torch.index_select (read+write {2 * T * K * D * 4 / 1e6:.0f} MB): {ms * 1e3:.1f} us = {2 * T * K * D * 4 / ms / 1e6:.1f} GB/s", flush=True)
        del sel


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        run()
    else:
        for env in ({"ALIVE_KNN_GATHER_CTA": "1"}, {}, {"ALIVE_KNN_GATHER_CTAS_PER_SM": "2"}, {"ALIVE_KNN_GATHER_CTAS_PER_SM": "6"},
                    {"ALIVE_KNN_GATHER_CTAS_PER_SM": "16"}):
            print("==", env or "default", flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=dict(os.environ, **env), check=False)
