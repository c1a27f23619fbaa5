"""How far is the tensor core's fp32 accumulation from the exact dot product of the SAME bf16 operands?
(the certificate's `kAccumSlack`, select.cu).  For every (query, frame) pair the fused kernel kept in its lists:
|screened score - float64 dot of the two packed bf16 rows|, on i.i.d. data (scores ~0.1) and on near-duplicate
data (scores 0.9 .. 1.0: the accumulator sits near 1 for most of the K loop - the worst case for truncation)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from alive_vc_b200 import _cabi, matching as M        # noqa: E402


def lists(q, lib, variant):
    plan = M.make_plan(q.n, lib.n, lib.d, q.device, variant, lib.format)
    cs = torch.empty((q.n, plan.lists, 8), device="cuda")
    ci = torch.empty((q.n, plan.lists, 8), dtype=torch.int32, device="cuda")
    rc = _cabi.load().alive_knn_search(q.packed.data_ptr(), lib.packed.data_ptr(), ctypes.byref(plan), cs.data_ptr(),
                                       ci.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _cabi.check(rc, "search")
    return cs, ci


def report(name, q, lib, variant):
    cs, ci = lists(q, lib, variant)
    T = q.n
    cs, ci = cs.view(T, -1), ci.view(T, -1).long()
    ok = ci >= 0
    worst, worst_rel, signed = 0.0, 0.0, []
    for t0 in range(0, T, 64):
        t1 = min(T, t0 + 64)
        rows = lib.packed[ci[t0:t1].clamp(min=0)].double()                   # [b, L, D]
        exact = (rows * q.packed[t0:t1].double()[:, None, :]).sum(dim=2)
        err = (cs[t0:t1].double() - exact)
        err = torch.where(ok[t0:t1], err, torch.zeros_like(err))
        worst = max(worst, float(err.abs().max()))
        signed.append(err[ok[t0:t1]])
    signed = torch.cat(signed)
    top = cs[ok]
    print(f"{name:34s} variant {variant}: pairs {signed.numel():8d}  scores [{float(top.min()):.3f}, {float(top.max()):.3f}]  "
          f"max |err| {worst:.3e}  mean err {float(signed.mean()):+.3e}  rms {float(signed.pow(2).mean().sqrt()):.3e}  "
          f"min {float(signed.min()):+.3e} max {float(signed.max()):+.3e}")


def main():
    g = torch.Generator(device="cuda").manual_seed(1)
    M.SCREEN_FORMAT = os.environ.get("PLANES", "bf16")        # explicit: no pack-time probe, both operands alike
    print("planes:", M.SCREEN_FORMAT)
    for D in (768, 1536):
        T, N = 512, 60_000
        q = M.pack_frames(torch.randn(D, T, device="cuda", generator=g))
        lib = M.pack_frames(torch.randn(D, N, device="cuda", generator=g))
        for v in (1, 2):
            report(f"iid D={D}", q, lib, v)
        for noise in (0.5, 0.2, 0.05, 0.01):
            cent = torch.randn(D, 30, device="cuda", generator=g)
            ref = cent[:, torch.randint(0, 30, (N,), device="cuda", generator=g)] + noise * torch.randn(D, N, device="cuda", generator=g)
            src = cent[:, torch.randint(0, 30, (T,), device="cuda", generator=g)] + noise * torch.randn(D, T, device="cuda", generator=g)
            q, lib = M.pack_frames(src), M.pack_frames(ref)
            for v in (1, 2):
                report(f"clusters noise={noise} D={D}", q, lib, v)
        # all-positive frames: every product is positive, the accumulator grows monotonically to ~1
        ref = torch.rand(D, N, device="cuda", generator=g) + 0.5
        src = torch.rand(D, T, device="cuda", generator=g) + 0.5
        q, lib = M.pack_frames(src), M.pack_frames(ref)
        for v in (1, 2):
            report(f"all-positive D={D}", q, lib, v)
        # skinny kernel (t <= 32)
        q32 = M.pack_frames(src[:, :32].contiguous())
        report(f"all-positive skinny D={D}", q32, lib, 3)


if __name__ == "__main__":
    main()
