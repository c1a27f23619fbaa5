"""Phase timing of finish_kernel (block 0), needs a library built with -DALIVE_FINISH_TIMING:
    nvcc <flags of alive_vc_b200/_cabi.py> -DALIVE_FINISH_TIMING -o dbg/libalive_knn_timing.so alive_vc_b200/csrc/*.cu
    ALIVE_KNN_LIB=dbg/libalive_knn_timing.so python tests/gpu_tools/finish_phases.py T N
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from alive_vc_b200 import _cabi, matching as M      # noqa: E402


def main():
    T, N = int(sys.argv[1]), int(sys.argv[2])
    lib = M.pack_frames(torch.randn(768, N, device="cuda"))
    src = torch.randn(1, 768, T, device="cuda")
    c = _cabi.load()
    for _ in range(5):
        M.run_match(src, lib, 4, 0.0, mode="screen")
    torch.cuda.synchronize()
    buf = (ctypes.c_uint64 * 16)()
    fn = c.alive_knn_debug_finish_times
    fn.argtypes = [ctypes.POINTER(ctypes.c_uint64)]
    fn(buf)
    t = list(buf)
    names = {1: "stage lists + normalise query", 2: "tau + per-warp top-k", 3: "S_k + certificate (warp 0)",
             4: "survivor compaction", 5: "exact rescoring", 7: "select top-k + stage", 8: "gather-mean-blend"}
    prev = t[0]
    for i in (1, 2, 3, 4, 5, 7, 8):
        print(f"T={T} N={N} phase {i} {names[i]:32s}: {t[i] - prev:7d} cycles")
        prev = t[i]
    print(f"total {t[8] - t[0]} cycles; survivors (query 0) = {int(M.last_info.sel_n[0])}")


if __name__ == "__main__":
    main()
