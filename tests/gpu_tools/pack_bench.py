"""K1 (alive_knn_pack) alone: CUDA-event timing of the library pack on a chunk far larger than L2, for both
input layouts (the reference's channel-major [D, n] and a row-major producer's [n, D]).

    python tests/gpu_tools/pack_bench.py [n_frames] [reps]

Algorithmic bytes per frame at D=768: D*(4 read + 4 raw + 2 packed) + 8 = 7,688; with the second bf16 plane
(the default for a single library) D*(4 + 4 + 2 + 2) + 12 = 9,228.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from alive_vc_b200 import matching as M  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    D = 768
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    dev = torch.device("cuda", 0)
    x_dn = torch.randn(D, n, device=dev)
    x_nd = x_dn.t().contiguous()
    for refine in (True, False):
        # with the second bf16 plane (the default for a single library): D*(4 + 4 + 2 + 2) + 12 = 9,228 B per frame
        dst = M.alloc_packed(n, D, dev, refine=refine)
        per_frame = D * (12 if refine else 10) + (12 if refine else 8)
        for name, view in (("channel-major [D,n]", x_dn), ("row-major [n,D]", x_nd.t())):
            for _ in range(2):
                M.pack_into(dst, 0, view)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                M.pack_into(dst, 0, view)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            gbs = n * per_frame / (ms * 1e-3) / 1e9
            print(f"pack {name:22s} n={n} second plane {'on ' if refine else 'off'} ({per_frame} B/frame): {ms:.4f} ms  "
                  f"{gbs:7.1f} GB/s  = {100 * gbs / peak:.1f}% of {peak} GB/s", flush=True)
        del dst


if __name__ == "__main__":
    main()
