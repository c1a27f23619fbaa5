"""Host-to-host chunk latency of HostStreamingMatcher (realtime_inference.py:130-191 with pinned buffers) for the
three copy modes, at cfg2 (T=32 x N=200k) and at the reference's realtime defaults (T=24 x N=3512)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import alive_vc_b200 as A                                      # noqa: E402
from alive_vc_b200.lifecycle import HostStreamingMatcher       # noqa: E402


def main():
    g = torch.Generator(device="cuda").manual_seed(1)
    for T, N in ((32, 200_000), (24, 3512)):
        lib = A.pack_library(torch.randn(1, 768, N, device="cuda", generator=g))
        chunk = torch.randn(1, 768, T)
        sm = A.StreamingMatcher(lib, T)
        dchunk = chunk.cuda()
        for _ in range(20):
            sm(dchunk)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(500):
            sm(dchunk)
        e1.record()
        torch.cuda.synchronize()
        print(f"T={T} N={N}: device-resident graph replay {e0.elapsed_time(e1) / 500 * 1e3:.1f} us per chunk", flush=True)
        for mode in (False, "out", "both"):
            hm = HostStreamingMatcher(lib, T, zero_copy=mode)
            for _ in range(50):
                hm(chunk)
            lat, lat_run = [], []
            for _ in range(2000):
                t0 = time.perf_counter()
                hm(chunk)
                lat.append((time.perf_counter() - t0) * 1e6)
            for _ in range(2000):                      # chunk already in the pinned buffer (producer wrote it there)
                t0 = time.perf_counter()
                hm.run()
                hm.result()
                lat_run.append((time.perf_counter() - t0) * 1e6)
            lat.sort()
            lat_run.sort()
            print(f"   zero_copy={mode!s:5}: call p50 {lat[1000]:.1f} us p99 {lat[1980]:.1f} us min {lat[0]:.1f} | "
                  f"run()+result() p50 {lat_run[1000]:.1f} us min {lat_run[0]:.1f}", flush=True)
            del hm


if __name__ == "__main__":
    main()
