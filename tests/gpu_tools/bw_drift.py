"""Does the achievable HBM bandwidth of a fresh box drift with time?  Every second: a device copy
(torch, 2 GiB of traffic), the pack kernel on 250k and 1M frames in both layouts, and nvidia-smi's
temperature / power / clocks.   python tests/gpu_tools/bw_drift.py [seconds]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from alive_vc_b200 import matching as M  # noqa: E402


def timed(fn, reps):
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
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 40
    dev = torch.device("cuda", 0)
    D = 768
    a = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    b = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    xs = {n: torch.randn(D, n, device=dev) for n in (250_000, 1_000_000)}
    xr = {n: xs[n].t().contiguous() for n in xs}
    dst = M.alloc_packed(1_000_000, D, dev)
    t0 = time.time()
    print("t_s copy_GB/s cm250k rm250k cm1M rm1M | temp,power,sm,mem", flush=True)
    while time.time() - t0 < secs:
        ms = timed(lambda: b.copy_(a), 3)
        row = [f"{time.time() - t0:5.1f}", f"{2 * (1 << 30) / (ms * 1e-3) / 1e9:7.0f}"]
        for n in xs:
            for v in (xs[n], xr[n].t()):
                ms = timed(lambda: M.pack_into(dst, 0, v), 3)
                row.append(f"{n * (D * 10 + 8) / (ms * 1e-3) / 1e9:7.0f}")
        row[3], row[4] = row[4], row[3]
        q = subprocess.run(["nvidia-smi", "--query-gpu=temperature.gpu,power.draw,clocks.sm,clocks.mem,temperature.memory",
                            "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
        print(" ".join(row), "|", q, flush=True)
        time.sleep(0.5)


if __name__ == "__main__":
    main()
