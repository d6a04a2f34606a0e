"""Drop-in for /root/reference/sampling/autoregressive_sampling.py:9-61:
the target-only baseline every driver times against speculative sampling.  One fused
filter+softmax+sample launch per token (kernel 1b) on a static KV cache."""
from __future__ import annotations

from typing import Optional

import torch

from .. import ops
from ..engine import ModelStepper


@torch.no_grad()
def autoregressive_sampling(x: torch.Tensor, model: torch.nn.Module, N: int, eos_token_id: Optional[int] = None,
                            temperature: float = 1, top_k: int = 0, top_p: float = 0, pad_token_id=None, *,
                            uniforms: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Generates exactly N tokens per row (the reference's `n = len(x)` is the batch dim, :13-21) unless
    EOS is sampled (:55, batch 1).  uniforms: (N, B) fp32, drawn with torch.rand if absent."""
    if not x.is_cuda:
        raise RuntimeError("autoregressive_sampling needs CUDA tensors: there is no CPU path")
    B, L = x.shape
    dev = x.device
    S = (L + N + 2 + 63) // 64 * 64
    st = ModelStepper(model, B, S, dev)
    tokens = torch.zeros(B, S, dtype=torch.int64, device=dev)
    tokens[:, :L] = x
    seq = torch.full((B,), L, dtype=torch.int32, device=dev)
    if uniforms is None:
        uniforms = torch.rand(N, B, device=dev)
    uniforms = uniforms.to(device=dev, dtype=torch.float32).contiguous()
    cur = torch.zeros(B, dtype=torch.int64, device=dev)
    flag = ops.default_flag(dev)
    if L > 1:
        st.prefill(tokens, L - 1)
    n_done = 0
    for i in range(N):
        # step i consumes the token at position L-1+i (the previously sampled one for i > 0)
        logits = st.forward(tokens, seq, i - 1, 1, cur if i > 0 else None)[:, 0]
        ops.norm_sample(logits, temperature, top_k or 0, top_p or 0.0, uniforms[i], tok_out=cur, err=flag)
        n_done = i + 1
        if eos_token_id is not None and B == 1 and int(cur[0]) == eos_token_id:
            break
    tokens[:, L + n_done - 1] = cur
    flag.check()
    return tokens[:, :L + n_done].clone()
