"""The torch custom ops of the kNN matching path (north_star: "host code stays Python/PyTorch and calls a
thin C-ABI torch custom op").

    torch.ops.alive_vc_b200.knn_match(source, reference, k, alpha, mode, variant, reference_grad)
        -> (out [B, T, D] float32, idx [B, T, k] int64, score [B, T, k] float32)
    torch.ops.alive_vc_b200.knn_scatter_grad(grad_rows, idx, n, scale) -> [n, D] float32

Both are registered with `torch.library.custom_op`: a CUDA implementation (ctypes into libalive_knn.so - there
is no CPU kernel and no fallback), a fake (meta) kernel so FakeTensor / torch.compile / export can trace through
callers, and an autograd formula that is the reference's own gradient:

    module/common.py:96-109          `reference` gets no gradient (scored under no_grad), source gets alpha * g
    module/voice_library.py:15-33    tokens.grad[:, :, j] = (1 - alpha) / k * sum of g over the (b, t) that picked j
                                     (only the gather at :31 is differentiable; the similarity path contributes 0)

`match_features` / `VoiceLibrary.match` are thin wrappers around `knn_match` (transpose view + dtype promotion).
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor

from . import _cabi
from . import matching as M


@torch.library.custom_op("alive_vc_b200::knn_match", mutates_args=(), device_types="cuda")
def knn_match(source: Tensor, reference: Tensor, k: int, alpha: float, mode: str, variant: int,
              reference_grad: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """common.py:100-109 in one pipeline.  `reference` is [B, D, N] (library b for batch item b) or [1, D, N] /
    a stride-0 expand (one library for every item).  `reference_grad` only selects the autograd formula."""
    return M._match_impl(source, reference, k, alpha, mode, variant)


@knn_match.register_fake
def _knn_match_fake(source, reference, k, alpha, mode, variant, reference_grad):
    B, D, T = source.shape
    return (source.new_empty((B, T, D), dtype=torch.float32), source.new_empty((B, T, k), dtype=torch.int64),
            source.new_empty((B, T, k), dtype=torch.float32))


@torch.library.custom_op("alive_vc_b200::knn_scatter_grad", mutates_args=(), device_types="cuda")
def knn_scatter_grad(grad_rows: Tensor, idx: Tensor, n: int, scale: float) -> Tensor:
    """Backward of the gather at voice_library.py:31: out[idx[r, j]] += scale * grad_rows[r] for every query row r
    and neighbour j.  grad_rows [R, D] float32 contiguous, idx [R, k] int64 -> [n, D] float32."""
    rows, d = grad_rows.shape
    k = idx.shape[1]
    g = grad_rows.contiguous().float()
    ix = idx.contiguous()
    out = torch.zeros((n, d), dtype=torch.float32, device=g.device)
    if rows > 0 and n > 0:
        with M._on(g.device):
            rc = _cabi.load().alive_knn_scatter_grad(g.data_ptr(), ix.data_ptr(), rows, k, d, float(scale),
                                                     out.data_ptr(), n, M._stream_ptr(g.device))
        _cabi.check(rc, "alive_knn_scatter_grad")
        M._count(1)
    return out


@knn_scatter_grad.register_fake
def _knn_scatter_grad_fake(grad_rows, idx, n, scale):
    return grad_rows.new_empty((n, grad_rows.shape[1]), dtype=torch.float32)


def _setup_context(ctx, inputs, output):
    source, reference, k, alpha, _mode, _variant, reference_grad = inputs
    _out, idx, _score = output
    ctx.k, ctx.alpha, ctx.reference_grad = k, alpha, reference_grad
    ctx.src_dtype, ctx.ref_dtype, ctx.ref_shape = source.dtype, reference.dtype, tuple(reference.shape)
    ctx.save_for_backward(idx)
    ctx.set_materialize_grads(False)


def _backward(ctx, g_out, _g_idx, _g_score):
    (idx,) = ctx.saved_tensors
    grad_source = grad_reference = None
    if g_out is not None:
        B, T, D = g_out.shape
        if ctx.needs_input_grad[0]:
            grad_source = (g_out.transpose(1, 2) * ctx.alpha).to(ctx.src_dtype)          # common.py:109
        if ctx.needs_input_grad[1] and ctx.reference_grad:
            rb, _, n = ctx.ref_shape
            ix = idx
            if rb != 1:             # one library per item: indices are relative to the item's own frames
                ix = idx + (torch.arange(B, device=idx.device, dtype=idx.dtype) * n).view(B, 1, 1)
            rows = knn_scatter_grad(g_out.reshape(B * T, D), ix.reshape(B * T, ctx.k), rb * n,
                                    (1.0 - ctx.alpha) / ctx.k)
            grad_reference = rows.view(rb, n, D).transpose(1, 2).to(ctx.ref_dtype)
    return grad_source, grad_reference, None, None, None, None, None


knn_match.register_autograd(_backward, setup_context=_setup_context)
