"""A/B of the finish-kernel CTA shape (ALIVE_KNN_FINISH_THREADS) on three batch sizes: whole-step time of the
screened match by CUDA events, and the step minus the search kernel's own duration (= everything after the search)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench                                               # noqa: E402
from alive_vc_b200 import matching as M                    # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    for T, N in ((1000, 100_000), (10_000, 1_250_000), (64_000, 500_000)):
        lib = bench.build_library(0, N, 5, dev)
        g = torch.Generator(device=dev).manual_seed(2)
        src = torch.randn(1, 768, T, device=dev, generator=g)
        for _ in range(3):
            M.run_match(src, lib, 4, 0.0, mode="screen")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20 if T <= 10_000 else 5
        e0.record()
        for _ in range(reps):
            M.run_match(src, lib, 4, 0.0, mode="screen")
        e1.record()
        torch.cuda.synchronize()
        print(f"FINISH_THREADS={os.environ.get('ALIVE_KNN_FINISH_THREADS', 'auto'):>5s}  T={T:6d} N={N:8d}: "
              f"{e0.elapsed_time(e1) / reps * 1e3:10.1f} us per step, fallback {M.last_info.fallback_queries()}", flush=True)
        del lib, src
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
