"""Generate golden vectors from the UNMODIFIED reference implementation.

TEST INFRASTRUCTURE ONLY.  Run in the build container (where the read-only
reference checkout lives at /root/reference); the GPU box does not have it, so
the outputs are committed under tests/golden/ and replayed by tests.

    python oracle/gen_golden.py            # writes tests/golden/*.npz

The reference functions (`module/common.py:96-109` match_features,
`module/voice_library.py:6-33` VoiceLibrary) are imported as they are; the only
shim is an empty `pyworld` module, because `module/common.py:5` imports pyworld
at top level for the (unrelated) F0 code and pyworld is not installed here.
The reference never returns the top-k indices, so they are captured by wrapping
`torch.topk` while the unmodified function runs (no restatement involved).

Inputs are regenerated from numpy seeds (`make_case_inputs`), so the fixtures
hold only the reference's outputs and stay small.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")
D = 768

# name -> dict(B, T, N, k, alpha, seed, kind)
CASES = {
    # kind "mf": match_features(source[B,D,T], reference[B,D,N])
    "mf_small":        dict(kind="mf", B=1, T=50,  N=300,  k=4,  alpha=0.0,  seed=1),
    "mf_batch2":       dict(kind="mf", B=2, T=33,  N=517,  k=4,  alpha=0.0,  seed=2),
    "mf_alpha":        dict(kind="mf", B=1, T=24,  N=3512, k=4,  alpha=0.25, seed=3),
    "mf_alpha1":       dict(kind="mf", B=1, T=7,   N=64,   k=4,  alpha=1.0,  seed=4),
    "mf_k1":           dict(kind="mf", B=1, T=40,  N=1000, k=1,  alpha=0.0,  seed=5),
    "mf_k8":           dict(kind="mf", B=1, T=40,  N=1000, k=8,  alpha=0.0,  seed=6),
    "mf_k16":          dict(kind="mf", B=1, T=10,  N=700,  k=16, alpha=0.0,  seed=7),
    "mf_k_eq_n":       dict(kind="mf", B=1, T=5,   N=4,    k=4,  alpha=0.0,  seed=8),
    "mf_t1":           dict(kind="mf", B=1, T=1,   N=2049, k=4,  alpha=0.0,  seed=9),
    "mf_realistic":    dict(kind="mf", B=1, T=450, N=3512, k=4,  alpha=0.0,  seed=10),
    "mf_ragged_tile":  dict(kind="mf", B=1, T=129, N=4097, k=4,  alpha=0.0,  seed=11),
    "mf_mid":          dict(kind="mf", B=1, T=96,  N=20000, k=4, alpha=0.0,  seed=12),
    # self-match / cross-utterance match as train_decoder.py:134-135 uses it
    "mf_self":         dict(kind="mf_self", B=3, T=120, N=120, k=4, alpha=0.0, seed=13),
    # strided (non-contiguous) library as realtime_inference.py:88 builds it ([:, :, ::4])
    "mf_strided":      dict(kind="mf_strided", B=1, T=24, N=500, k=4, alpha=0.0, seed=14),
    # duplicated library rows: exact ties in similarity
    "mf_dupes":        dict(kind="mf_dupes", B=1, T=16, N=256, k=4, alpha=0.0, seed=15),
    # clustered library: many near neighbours (stresses the candidate screen)
    "mf_clustered":    dict(kind="mf_clustered", B=1, T=64, N=6000, k=4, alpha=0.0, seed=16),
    # kind "vl": VoiceLibrary(num_tokens=N).match(source[B,D,T]) incl. gradients
    "vl_default":      dict(kind="vl", B=2, T=120, N=512,  k=4, alpha=0.0, seed=21),
    "vl_alpha":        dict(kind="vl", B=3, T=40,  N=512,  k=4, alpha=0.5, seed=22),
    "vl_big":          dict(kind="vl", B=1, T=64,  N=5000, k=4, alpha=0.0, seed=23),
}


def make_case_inputs(spec: dict):
    """Deterministic float32 inputs for a case (numpy PCG64, so identical on
    every machine).  Returns (source[B,D,T], reference[B,D,N]) - for kind "vl"
    reference is tokens[1,D,N]."""
    rng = np.random.default_rng(spec["seed"])
    B, T, N = spec["B"], spec["T"], spec["N"]
    kind = spec["kind"]
    src = rng.standard_normal((B, D, T), dtype=np.float32)
    if kind == "vl":
        ref = rng.standard_normal((1, D, N), dtype=np.float32)
    elif kind == "mf_self":
        ref = np.roll(src, 1, axis=0).copy() if spec.get("roll", True) else src.copy()
    elif kind == "mf_strided":
        big = rng.standard_normal((B, D, N * 4), dtype=np.float32)
        ref = big[:, :, ::4]                      # non-contiguous view
    elif kind == "mf_dupes":
        base = rng.standard_normal((B, D, N // 4), dtype=np.float32)
        ref = np.concatenate([base, base, base, base], axis=2)
    elif kind == "mf_clustered":
        cent = rng.standard_normal((B, D, 12), dtype=np.float32)
        which = rng.integers(0, 12, size=N)
        ref = cent[:, :, which] + 0.05 * rng.standard_normal((B, D, N), dtype=np.float32)
        ref = ref.astype(np.float32)
        qwhich = rng.integers(0, 12, size=T)
        src = (cent[:, :, qwhich] + 0.05 * rng.standard_normal((B, D, T), dtype=np.float32)).astype(np.float32)
    else:
        ref = rng.standard_normal((B, D, N), dtype=np.float32)
    return src, ref


def _import_reference():
    sys.modules.setdefault("pyworld", types.ModuleType("pyworld"))
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    from module.common import match_features          # noqa: E402
    from module.voice_library import VoiceLibrary     # noqa: E402
    return match_features, VoiceLibrary


def main():
    import torch

    match_features, VoiceLibrary = _import_reference()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    captured = {}
    real_topk = torch.topk

    def spy_topk(*a, **kw):
        out = real_topk(*a, **kw)
        captured["values"] = out.values.detach().clone()
        captured["indices"] = out.indices.detach().clone()
        return out

    for name, spec in CASES.items():
        src_np, ref_np = make_case_inputs(spec)
        src = torch.from_numpy(np.ascontiguousarray(src_np))
        ref = torch.from_numpy(ref_np)          # keeps numpy strides (mf_strided stays non-contiguous)
        k, alpha = spec["k"], spec["alpha"]
        torch.topk = spy_topk
        try:
            if spec["kind"] == "vl":
                vl = VoiceLibrary(num_tokens=spec["N"])
                with torch.no_grad():
                    vl.tokens.copy_(ref)
                src_g = src.clone().requires_grad_(True)
                out = vl.match(src_g, k=k, alpha=alpha)
                g = torch.from_numpy(np.random.default_rng(spec["seed"] + 1000)
                                     .standard_normal(tuple(out.shape), dtype=np.float32))
                out.backward(g)
                extra = dict(grad_tokens=vl.tokens.grad.numpy().copy(),
                             grad_source=src_g.grad.numpy().copy())
                out = out.detach()
            else:
                out = match_features(src, ref, k=k, alpha=alpha)
                extra = {}
        finally:
            torch.topk = real_topk
        assert tuple(out.shape) == (spec["B"], D, spec["T"])
        np.savez_compressed(
            os.path.join(GOLDEN_DIR, name + ".npz"),
            out=out.numpy().copy(),
            out_strides=np.array(out.stride(), dtype=np.int64),
            indices=captured["indices"].numpy().astype(np.int32),
            values=captured["values"].numpy().astype(np.float32),
            **extra,
        )
        print(f"{name}: out{tuple(out.shape)} strides{tuple(out.stride())} idx{tuple(captured['indices'].shape)}")

    # error behaviour of the reference: k > N and batch mismatch
    errs = {}
    try:
        match_features(torch.randn(1, D, 3), torch.randn(1, D, 2), k=4)
    except RuntimeError as e:
        errs["k_gt_n"] = str(e).splitlines()[0]
    try:
        match_features(torch.randn(2, D, 3), torch.randn(1, D, 20), k=4)
    except RuntimeError as e:
        errs["batch_mismatch"] = str(e).splitlines()[0]
    try:
        match_features(torch.randn(1, D, 3), torch.zeros(1, D, 0), k=4)
    except RuntimeError as e:
        errs["empty_library"] = str(e).splitlines()[0]
    with open(os.path.join(GOLDEN_DIR, "errors.txt"), "w") as f:
        for kk, v in errs.items():
            f.write(f"{kk}\t{v}\n")
    print(errs)


if __name__ == "__main__":
    main()
