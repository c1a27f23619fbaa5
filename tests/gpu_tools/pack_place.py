"""Is the pack kernel's speed a property of the kernel or of where its buffers live?
   python tests/gpu_tools/pack_place.py <dummy_gb_before> <dummy_gb_after>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from alive_vc_b200 import matching as M  # noqa: E402


def timed(fn, reps=10):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    before, after = float(sys.argv[1]), float(sys.argv[2])
    n, D = 250_000, 768
    dev = torch.device("cuda", 0)
    d0 = torch.empty(int(before * (1 << 30)), dtype=torch.uint8, device=dev) if before > 0 else None
    x_dn = torch.randn(D, n, device=dev)
    x_nd = x_dn.t().contiguous()
    dst = M.alloc_packed(n, D, dev)
    d1 = torch.empty(int(after * (1 << 30)), dtype=torch.uint8, device=dev) if after > 0 else None
    for t in (d0, d1):
        if t is not None:
            t.zero_()
    ptrs = [hex(t.data_ptr()) for t in (x_dn, x_nd, dst.raw, dst.packed)]
    cm = timed(lambda: M.pack_into(dst, 0, x_dn))
    rm = timed(lambda: M.pack_into(dst, 0, x_nd.t()))
    cp = timed(lambda: dst.raw.copy_(x_nd))
    b = n * (D * 10 + 8)
    print(f"dummy {before:.0f}+{after:.0f} GB: cm {b / cm / 1e6:6.0f} GB/s  rm {b / rm / 1e6:6.0f} GB/s  "
          f"torch copy x_nd->raw {2 * n * D * 4 / cp / 1e6:6.0f} GB/s  ptrs {ptrs}", flush=True)


if __name__ == "__main__":
    main()
