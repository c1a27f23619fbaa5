"""CPU oracle for the ALiVE-VC kNN voice-library matching path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`alive_vc_b200/`)
imports this module; it is used by `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` as the checker / the
CPU arm, never as the thing shipped.

What it restates (reference = uthree/ALiVE-VC, paths relative to its root):

* `module/common.py:96-109`  `match_features(source, reference, k, alpha)`
* `module/voice_library.py:15-33`  `VoiceLibrary.match(source, k, alpha)`
  (same arithmetic, the library is `tokens.expand(B, D, N)`)

The arithmetic itself lives in PyTorch (third party, unpinned: the reference's
`requirements.txt:1` is a bare `torch`; this image has torch 2.11.0).  The
reference ships NO tests, golden vectors or fixtures for this path
(SURVEY.md §4, §8c), so the oracle is pinned the only way available: against
outputs of the unmodified reference function imported in the build container
(`oracle/gen_golden.py` -> `tests/golden/*.npz`; `tests/test_oracle.py`
replays them without the reference being present).

A third, independent restatement in plain C lives in `oracle/knn_oracle.c` (loader: `oracle/c_oracle.py`);
`tests/test_oracle.py` pins it to the same golden vectors and to this module.

Two restatements are provided here:

* `match_features_np`  - numpy, float32 end to end, used by the parity tests.
  It also returns the top-k indices and similarities, which the reference
  computes (`common.py:105`) but never returns.
* `match_features_torch` - the same steps spelled with torch CPU ops, used only
  to time the CPU baseline (the reference's own speed is torch/MKL's speed).
"""
from __future__ import annotations

import numpy as np


def _as_f32(x) -> np.ndarray:
    a = np.asarray(x)
    if a.dtype != np.float32:
        a = a.astype(np.float32)
    return a


def cosine_scores_np(source, reference) -> np.ndarray:
    """`common.py:100-104`: transpose to [B,L,D], L2-normalise each frame
    (divide by `norm`, no epsilon), batched matmul -> [B,T,N] float32."""
    src = np.swapaxes(_as_f32(source), 1, 2)          # common.py:100
    ref = np.swapaxes(_as_f32(reference), 1, 2)       # common.py:101
    if src.shape[0] != ref.shape[0]:
        # torch.bmm raises on a batch mismatch (no broadcasting), SURVEY §8(a)
        raise RuntimeError("batch1 and batch2 must have same batch size")
    src_norm = np.sqrt(np.sum(src * src, axis=2, keepdims=True, dtype=np.float32))  # :102
    ref_norm = np.sqrt(np.sum(ref * ref, axis=2, keepdims=True, dtype=np.float32))  # :103
    with np.errstate(divide="ignore", invalid="ignore"):
        sn = (src / src_norm).astype(np.float32)
        rn = (ref / ref_norm).astype(np.float32)
    return np.matmul(sn, np.swapaxes(rn, 1, 2)).astype(np.float32)   # :104


def topk_desc_np(scores: np.ndarray, k: int):
    """`common.py:105` `torch.topk(cos_sims, k, dim=2)`: largest first, NaN
    ranks above everything; exact ties are implementation-defined in torch,
    here they resolve to the lowest index (stable sort)."""
    n = scores.shape[-1]
    if k > n or n == 0:
        raise RuntimeError("selected index k out of range")
    key = np.where(np.isnan(scores), np.float32(np.inf), scores)
    # NaN must beat +inf: give NaN a strictly larger key via a two-level sort
    nan_first = np.isnan(scores)
    order = np.lexsort((np.arange(n)[None, None, :].repeat(scores.shape[1], 1).repeat(scores.shape[0], 0),
                        -key, ~nan_first), axis=-1)
    idx = order[..., :k].astype(np.int64)
    val = np.take_along_axis(scores, idx, axis=-1)
    return val, idx


def gather_mean_np(reference, idx: np.ndarray) -> np.ndarray:
    """`common.py:107`: gather the RAW (un-normalised) library rows of the k
    winners and average them.  Bit-exact model of torch's CPU `.mean(dim=2)`
    on `[B,T,k,D]`: sequential float32 sum in descending-score
    order, then a true division by k."""
    ref = np.swapaxes(_as_f32(reference), 1, 2)        # [B,N,D]
    B, T, k = idx.shape
    out = np.empty((B, T, ref.shape[2]), dtype=np.float32)
    for b in range(B):
        rows = ref[b][idx[b]]                          # [T,k,D]
        acc = rows[:, 0, :].copy()
        for j in range(1, k):
            acc = (acc + rows[:, j, :]).astype(np.float32)
        out[b] = (acc / np.float32(k)).astype(np.float32)
    return out


def match_features_np(source, reference, k: int = 4, alpha: float = 0.0,
                      return_indices: bool = False):
    """Full restatement of `module/common.py:96-109`.

    source [B,D,T], reference [B,D,N] float32 -> [B,D,T] float32 whose memory
    is a contiguous [B,T,D] block (the reference returns that transposed view,
    `common.py:108`), blended as `result*(1-alpha) + input*alpha` (`:109`).
    """
    source = _as_f32(source)
    reference = _as_f32(reference)
    scores = cosine_scores_np(source, reference)
    val, idx = topk_desc_np(scores, k)
    res = gather_mean_np(reference, idx)                         # [B,T,D]
    res = np.swapaxes(res, 1, 2)                                 # :108 (view)
    a1 = np.float32(1 - alpha)
    a0 = np.float32(alpha)
    with np.errstate(invalid="ignore"):
        out = ((res * a1).astype(np.float32) + (source * a0).astype(np.float32)).astype(np.float32)
    if return_indices:
        return out, idx, val
    return out


def voice_library_match_np(tokens, source, k: int = 4, alpha: float = 0.0,
                           return_indices: bool = False):
    """`module/voice_library.py:15-33`: `tokens` [1,D,N] expanded over the
    batch of `source` [B,D,T], then the same steps as `match_features`."""
    tokens = _as_f32(tokens)
    source = _as_f32(source)
    ref = np.broadcast_to(tokens, (source.shape[0],) + tokens.shape[1:])  # :16-19
    return match_features_np(source, ref, k, alpha, return_indices)


def voice_library_grads_np(tokens, source, grad_out, k: int = 4, alpha: float = 0.0):
    """Gradients of `VoiceLibrary.match` (probed):
    tokens.grad[0,:,j] = (1-alpha)/k * sum over (b,t) with j in topk(b,t) of
    g[b,:,t];  source.grad = alpha * g  (the similarity path contributes 0)."""
    tokens = _as_f32(tokens)
    g = _as_f32(grad_out)
    _, idx, _ = voice_library_match_np(tokens, source, k, alpha, True)
    gt = np.zeros_like(tokens, dtype=np.float64)
    B, T, _ = idx.shape
    scale = (1.0 - alpha) / k
    for b in range(B):
        for t in range(T):
            for j in idx[b, t]:
                gt[0, :, j] += scale * g[b, :, t]
    return gt.astype(np.float32), (np.float32(alpha) * g).astype(np.float32)


def match_features_torch(source, reference, k: int = 4, alpha: float = 0.0):
    """The same steps with torch CPU ops - used ONLY to time the CPU baseline
    (`bench.py` `cpu_baseline` / `--impl reference`): the reference's speed on
    host cores is the speed of these torch/MKL calls (`common.py:100-109`)."""
    import torch

    with torch.no_grad():
        s = source.transpose(1, 2)
        r = reference.transpose(1, 2)
        sims = torch.bmm(s / torch.norm(s, dim=2, keepdim=True),
                         (r / torch.norm(r, dim=2, keepdim=True)).transpose(1, 2))
        best = torch.topk(sims, k, dim=2)
        picked = [r[b][best.indices[b]] for b in range(s.shape[0])]
        res = torch.stack(picked, dim=0).mean(dim=2).transpose(1, 2)
        return res * (1 - alpha) + source * alpha, best.indices


def indices_match_mod_ties(idx_a: np.ndarray, idx_b: np.ndarray, scores: np.ndarray,
                           tol: float = 1e-6):
    """north_star parity rule: neighbour indices bit-exact except where fp32
    similarities tie within `tol`.  `scores` [B,T,N] are the oracle's
    similarities.  Returns (ok, n_exact_rows, n_tie_rows, first_bad)."""
    B, T, k = idx_a.shape
    exact = 0
    ties = 0
    for b in range(B):
        for t in range(T):
            a = idx_a[b, t]
            c = idx_b[b, t]
            if np.array_equal(a, c):
                exact += 1
                continue
            sa = scores[b, t, a]
            sc = scores[b, t, c]
            # position-wise: the score sequences must agree within tol, i.e. any
            # index difference is a swap/replacement among near-tied entries
            if np.all(np.abs(sa - sc) <= tol) or (np.isnan(sa) == np.isnan(sc)).all() and \
                    np.all((np.abs(sa - sc) <= tol) | (np.isnan(sa) & np.isnan(sc))):
                ties += 1
                continue
            return False, exact, ties, (b, t, a.tolist(), c.tolist(), sa.tolist(), sc.tolist())
    return True, exact, ties, None
