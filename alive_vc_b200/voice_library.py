"""Drop-in for `module.voice_library.VoiceLibrary` (module/voice_library.py:6-33).

Same constructor, same `tokens` parameter ([1, hubert_dim, num_tokens], state_dict key
"tokens", so reference checkpoints load unchanged), same `forward` / `match` signatures.
`match` stays differentiable w.r.t. `tokens` exactly as in the reference: the gradient
flows only through the gather (voice_library.py:31), i.e.
    tokens.grad[0, :, j] = (1-alpha)/k * sum over (b,t) with j in topk(b,t) of g[b, :, t]
    source.grad          = alpha * g
(the similarity / top-k path contributes exactly zero).  The work and its autograd formula
live in the registered custom op `torch.ops.alive_vc_b200.knn_match` (alive_vc_b200/ops.py).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import matching as M


class VoiceLibrary(nn.Module):
    def __init__(self, num_tokens=512, hubert_dim=768):
        super().__init__()
        self.tokens = nn.Parameter(torch.randn(1, hubert_dim, num_tokens))   # voice_library.py:9
        self.hubert_dim = hubert_dim

    def forward(self, source):
        return self.match(source)                                             # voice_library.py:12-13

    def match(self, source, k=4, alpha=0.0, *, return_indices=False, mode="auto", variant=0):
        """voice_library.py:15-33.  source [B, hubert_dim, T] -> [B, hubert_dim, T]."""
        from . import ops
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
        M._check_dtypes(source, self.tokens)
        out_btd, idx, _ = ops.knn_match(source, self.tokens, k, float(alpha), mode, variant, True)
        out = out_btd.transpose(1, 2)
        want = torch.promote_types(source.dtype, self.tokens.dtype)
        if out.dtype != want:
            out = out.to(want)
        return (out, idx) if return_indices else out

    def packed(self) -> M.PackedFrames:
        """The packed (bf16 + raw row-major) form of the current tokens (cached until they change when the
        library holds at least matching.PACK_CACHE_MIN_ELEMENTS elements; smaller ones are re-packed per call)."""
        t = self.tokens.detach()
        return M.cached_pack(self.tokens, t[0] if t.dtype == torch.float32 else t[0].float())

    def invalidate(self):
        """Forget the cached packed copy of `tokens`.  Only needed after writes the version counter does not see
        (`VL.tokens.data[:, :, n] = t` as in generate_voice_library.py:38, custom kernels) on a library of at
        least matching.PACK_CACHE_MIN_ELEMENTS elements; in-place ops and optimizer steps are tracked."""
        M.clear_pack_cache(self.tokens)
