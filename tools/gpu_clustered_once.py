"""A few matches of cfg1's shape against a clustered library (100 clusters, noise 0.2) with bf16 planes - the case
that needs the refined collect pass - for profiling (ncu -k regex:knn_search_kernel picks the collect launches)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import alive_vc_b200 as A                     # noqa: E402
from alive_vc_b200 import matching as M      # noqa: E402
import bench                                  # noqa: E402


def main():
    fmt = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    src, ref = bench.build_clustered(1000, 100_000, 100, 0.2, 7, torch.device("cuda"))
    lib = A.pack_library(ref, fmt=fmt)
    for _ in range(4):
        M.run_match(src, lib, 4, 0.0, mode="screen")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        M.run_match(src, lib, 4, 0.0, mode="screen")
    e1.record()
    torch.cuda.synchronize()
    print(f"planes {fmt}: {e0.elapsed_time(e1) / 10:.4f} ms per call, uncertified {M.last_info.fallback_queries()}, "
          f"exhaustive scan {M.last_info.exact_scan_queries()}")


if __name__ == "__main__":
    main()
