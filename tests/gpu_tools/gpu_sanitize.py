"""Small end-to-end invocations for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  python tests/gpu_tools/gpu_sanitize.py
    compute-sanitizer --tool racecheck python tests/gpu_tools/gpu_sanitize.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import alive_vc_b200 as A                       # noqa: E402
from alive_vc_b200 import matching as M        # noqa: E402
from oracle import knn_oracle as O             # noqa: E402


def main():
    rng = np.random.default_rng(3)
    for (B, T, N, k, alpha, mode, variant) in [(1, 40, 700, 4, 0.0, "screen", 1), (1, 300, 2100, 4, 0.25, "screen", 2),
                                               (2, 9, 257, 2, 0.0, "screen", 0), (1, 12, 90, 16, 0.0, "exact", 0),
                                               (1, 5, 600, 4, 0.0, "exact", 0)]:
        src = rng.standard_normal((B, 768, T), dtype=np.float32)
        ref = rng.standard_normal((B, 768, N), dtype=np.float32)
        out, idx = A.match_features(torch.from_numpy(src).cuda(), torch.from_numpy(ref).cuda(), k, alpha,
                                    return_indices=True, mode=mode, variant=variant)
        torch.cuda.synchronize()
        out_o, idx_o, _ = O.match_features_np(src, ref, k, alpha, True)
        assert np.array_equal(idx.cpu().numpy(), idx_o), (T, N)
        assert np.array_equal(out.cpu().numpy(), out_o), (T, N)
        print("ok", B, T, N, k, alpha, mode, variant, flush=True)
    # duplicates -> certificate fails -> exact scan with gather; autograd scatter
    base = rng.standard_normal((1, 768, 200), dtype=np.float32)
    ref = np.concatenate([base] * 3, axis=2)
    ref[:, :, 77] = 0          # a zero frame: every query goes to the exact scan (NaN ranks first)
    out, idx = A.match_features(torch.from_numpy(base[:, :, :10].copy()).cuda(), torch.from_numpy(ref).cuda(), 3, 0.0,
                                return_indices=True, mode="screen")
    torch.cuda.synchronize()
    # one realtime chunk: the skinny kernel (operands swapped, resident query chunk)
    src = rng.standard_normal((1, 768, 24), dtype=np.float32)
    ref = rng.standard_normal((1, 768, 3512), dtype=np.float32)
    out, idx = A.match_features(torch.from_numpy(src).cuda(), torch.from_numpy(ref).cuda(), 4, 0.0, return_indices=True)
    torch.cuda.synchronize()
    assert np.array_equal(idx.cpu().numpy(), O.match_features_np(src, ref, 4, 0.0, True)[1])
    print("ok skinny chunk", flush=True)
    # tight clusters on a library big enough for the collect pass (second tensor-core pass + rescoring)
    g = torch.Generator(device="cuda").manual_seed(1)
    cent = torch.randn(768, 40, device="cuda", generator=g)
    refc = (cent[:, torch.randint(0, 40, (70_000,), device="cuda", generator=g)] +
            0.2 * torch.randn(768, 70_000, device="cuda", generator=g))[None]
    srcc = (cent[:, torch.randint(0, 40, (260,), device="cuda", generator=g)] +
            0.2 * torch.randn(768, 260, device="cuda", generator=g))[None]
    libc = A.pack_library(refc)
    _, idx_s, _ = M.run_match(srcc, libc, 4, 0.0, mode="screen")
    assert M.last_info.collect and M.last_info.fallback_queries() > 0
    torch.cuda.synchronize()
    print("ok collect pass", M.last_info.fallback_queries(), M.last_info.exact_scan_queries(), flush=True)
    vl = A.VoiceLibrary(num_tokens=300).cuda()
    s = torch.randn(2, 768, 11, device="cuda", requires_grad=True)
    vl.match(s, alpha=0.5).sum().backward()
    torch.cuda.synchronize()
    print("ok fallback + autograd", flush=True)


if __name__ == "__main__":
    main()
