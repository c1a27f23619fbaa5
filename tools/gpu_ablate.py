"""Ablation / profiling driver for the fused search kernel (run on the B200 box).

    python tools/gpu_ablate.py ablate          # epilogue ablations, both variants
    python tools/gpu_ablate.py prof <variant> [T N]   # a few launches for ncu
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def setup(T, N, variant):
    import torch
    from alive_vc_b200 import _cabi, matching as M
    g = torch.Generator(device="cuda").manual_seed(3)
    lib = M.alloc_packed(N, 768, "cuda")
    for r0 in range(0, N, 100000):
        n = min(100000, N - r0)
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
    return run, plan, (lib, q, cs, ci)


def time_it(run, reps):
    import torch
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def one(variant, debug, T, N):
    os.environ["ALIVE_KNN_DEBUG_EPILOGUE"] = str(debug)
    run, plan, keep = setup(T, N, variant)
    ms = time_it(run, 5)
    tf = 2.0 * T * N * 768 / (ms * 1e-3) / 1e12
    print(f"variant {variant} debug {debug} T={T} N={N}: {ms:.3f} ms {tf:.1f} TFLOP/s "
          f"(seg={plan.segments} tps={plan.tiles_per_segment} grid={plan.grid})", flush=True)


def main():
    what = sys.argv[1]
    if what == "ablate":
        for (T, N) in [(8192, 400000), (1024, 100000)]:
            for variant in (1, 2):
                for debug in (0, 1, 2):
                    subprocess.run([sys.executable, os.path.abspath(__file__), "one", str(variant), str(debug),
                                    str(T), str(N)], cwd=ROOT)
    elif what == "one":
        one(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
    elif what == "sustain":
        # the search kernel alone, back to back for seconds (power-capped regime), at a bench shape
        import torch
        import bench
        from alive_vc_b200 import _cabi, matching as M
        T, N, variant = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
        lib = bench.build_library(0, N, 7, torch.device("cuda", 0))
        q = M.pack_frames(torch.randn(768, T, device="cuda"))
        plan = M.make_plan(T, N, 768, q.device, variant)
        cs = torch.empty((T, plan.lists, 8), device="cuda")
        ci = torch.empty((T, plan.lists, 8), dtype=torch.int32, device="cuda")
        c = _cabi.load()
        st = torch.cuda.current_stream().cuda_stream

        def run():
            _cabi.check(c.alive_knn_search(q.packed.data_ptr(), lib.packed.data_ptr(), ctypes.byref(plan),
                                           cs.data_ptr(), ci.data_ptr(), st), "search")
        for _ in range(5):
            run()
        ms = time_it(run, 15)
        print(f"sustain T={T} N={N} variant={variant} debug={os.environ.get('ALIVE_KNN_DEBUG_EPILOGUE', '0')}: "
              f"{ms:.3f} ms  {2.0 * T * N * 768 / (ms * 1e-3) / 1e12:.1f} TFLOP/s", flush=True)
    elif what == "pipeline":
        # the whole one-call pipeline a few times (for an ncu launch list / per-kernel breakdown)
        import torch
        import bench
        from alive_vc_b200 import matching as M
        T, N = int(sys.argv[2]), int(sys.argv[3])
        lib = bench.build_library(0, N, 7, torch.device("cuda", 0))
        src = torch.randn(1, 768, T, device="cuda")
        for _ in range(3):
            M.run_match(src, lib, 4, 0.0, mode="screen")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            M.run_match(src, lib, 4, 0.0, mode="screen")
        e1.record()
        torch.cuda.synchronize()
        print(f"pipeline T={T} N={N}: {e0.elapsed_time(e1) / 5:.3f} ms per call")
    elif what == "prof":
        import torch
        variant = int(sys.argv[2])
        T = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
        N = int(sys.argv[4]) if len(sys.argv) > 4 else 200000
        run, plan, keep = setup(T, N, variant)
        for _ in range(4):
            run()
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
