"""Drop-in for `module.voice_library.VoiceLibrary` (module/voice_library.py:6-33).

Same constructor, same `tokens` parameter ([1, hubert_dim, num_tokens], state_dict key
"tokens", so reference checkpoints load unchanged), same `forward` / `match` signatures.
`match` stays differentiable w.r.t. `tokens` exactly as in the reference: the gradient
flows only through the gather (voice_library.py:31), i.e.
    tokens.grad[0, :, j] = (1-alpha)/k * sum over (b,t) with j in topk(b,t) of g[b, :, t]
    source.grad          = alpha * g
(the similarity / top-k path contributes exactly zero).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _cabi
from . import matching as M


class _LibraryMatchFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, source, tokens, k, alpha, mode, variant):
        B, D, T = source.shape
        lib = M.cached_pack(tokens, tokens.detach()[0].float() if tokens.dtype != torch.float32 else tokens.detach()[0])
        src32 = source.detach() if source.dtype == torch.float32 else source.detach().float()
        out_btd, idx, _ = M.match_packed(src32, lib, k, alpha, mode, variant)
        ctx.save_for_backward(idx)
        ctx.meta = (k, alpha, tuple(tokens.shape), tokens.dtype, source.dtype)
        out = out_btd.transpose(1, 2)
        return out if out.dtype == source.dtype else out.to(source.dtype)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        k, alpha, tok_shape, tok_dtype, src_dtype = ctx.meta
        grad_source = grad_tokens = None
        if ctx.needs_input_grad[0]:
            grad_source = (g * alpha).to(src_dtype)
        if ctx.needs_input_grad[1]:
            B, D, T = g.shape
            n = tok_shape[2]
            g_rows = g.transpose(1, 2).reshape(B * T, D).contiguous().float()
            grad_rows = torch.zeros((n, D), dtype=torch.float32, device=g.device)
            rc = _cabi.load().alive_knn_scatter_grad(
                g_rows.data_ptr(), idx.reshape(B * T, k).contiguous().data_ptr(), B * T, k, D,
                float((1.0 - alpha) / k), grad_rows.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
            _cabi.check(rc, "alive_knn_scatter_grad")
            grad_tokens = grad_rows.t().unsqueeze(0).to(tok_dtype)
        return grad_source, grad_tokens, None, None, None, None


class VoiceLibrary(nn.Module):
    def __init__(self, num_tokens=512, hubert_dim=768):
        super().__init__()
        self.tokens = nn.Parameter(torch.randn(1, hubert_dim, num_tokens))   # voice_library.py:9
        self.hubert_dim = hubert_dim

    def forward(self, source):
        return self.match(source)                                             # voice_library.py:12-13

    def match(self, source, k=4, alpha=0.0, *, return_indices=False, mode="auto", variant=0):
        """voice_library.py:15-33.  source [B, hubert_dim, T] -> [B, hubert_dim, T]."""
        if source.dim() != 3 or source.shape[1] != self.hubert_dim:
            raise RuntimeError(
                f"Expected size for first two dimensions of batch2 tensor to be: "
                f"[{source.shape[0]}, {source.shape[1] if source.dim() == 3 else '?'}] but got: "
                f"[{source.shape[0]}, {self.hubert_dim}].")
        n = self.tokens.shape[2]
        if not isinstance(k, int) or k < 1 or k > n:
            raise RuntimeError("selected index k out of range")
        M._require_cuda(source, "source")
        M._require_cuda(self.tokens, "tokens")
        needs_grad = torch.is_grad_enabled() and (self.tokens.requires_grad or source.requires_grad)
        if needs_grad:
            out = _LibraryMatchFn.apply(source, self.tokens, k, float(alpha), mode, variant)
            if return_indices:
                with torch.no_grad():
                    idx, _ = M.match_indices(source, self.tokens.expand(source.shape[0], -1, -1), k, mode, variant)
                return out, idx
            return out
        with torch.no_grad():
            lib = M.cached_pack(self.tokens, self.tokens.detach()[0].float())
            src32 = source if source.dtype == torch.float32 else source.float()
            out_btd, idx, _ = M.match_packed(src32, lib, k, float(alpha), mode, variant)
        out = out_btd.transpose(1, 2)
        if out.dtype != source.dtype:
            out = out.to(source.dtype)
        return (out, idx) if return_indices else out

    def packed(self) -> M.PackedFrames:
        """The packed (bf16 + raw row-major) form of the current tokens, cached until they change."""
        return M.cached_pack(self.tokens, self.tokens.detach()[0].float())
