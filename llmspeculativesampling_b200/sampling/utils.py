"""Drop-in for /root/reference/sampling/utils.py:152-245 — same names, arguments and error behaviour,
served by the sm_100a kernels (no CPU path)."""
from __future__ import annotations

from typing import Optional

import torch

from .. import ops


def top_k_top_p_filter(logits: torch.Tensor, top_k: int = 0, top_p: float = 0.0) -> torch.Tensor:
    """In place, like the reference (utils.py:152-179): entries outside the kept set become -inf.
    The fused kernel never materialises filtered logits, so the keep mask is taken from the kept
    set's probabilities (an entry whose probability underflows to exactly 0 is masked too — it has
    probability 0 either way)."""
    assert logits.dim() == 2
    if not top_k and not top_p:
        return logits                                                     # both stages disabled: nothing to mask
    probs = ops.norm_probs(logits, 1.0, top_k or 0, top_p or 0.0)
    logits.masked_fill_(probs == 0, float("-inf"))
    return logits


def norm_logits(logits: torch.Tensor, temperature: float, top_k: float, top_p: float) -> torch.Tensor:
    """(rows, V) logits -> (rows, V) fp32 probabilities (utils.py:182-210).  Raises
    RuntimeError('norm logits error') where the reference does (one flag read instead of 3 syncs)."""
    assert logits.dim() == 2                                              # utils.py:194
    ops._require_cuda(logits, "logits")
    flag = ops.default_flag(logits.device)
    probs = ops.norm_probs(logits, temperature, top_k or 0, top_p or 0.0, err=flag)
    flag.check()
    return probs


def sample(probs: torch.Tensor, num_samples: int = 1, u: Optional[torch.Tensor] = None,
           generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """(rows, V) weights -> (rows, 1) int64 (utils.py:213-233).  torch.multinomial is replaced by the
    inverse-CDF rule of include/specdec_b200.h on one uniform per row (`u`, drawn here if absent)."""
    if num_samples != 1:
        raise NotImplementedError("only num_samples=1 is on the speculative hot path")
    p2 = probs if probs.dim() == 2 else probs.reshape(1, -1)
    ops._require_cuda(p2, "probs")
    if u is None:
        u = torch.rand(p2.shape[0], device=p2.device, dtype=torch.float32, generator=generator)
    flag = ops.default_flag(p2.device)
    tok = ops.sample_rows(p2, u.reshape(-1).to(torch.float32).contiguous(), err=flag)
    flag.check()                                                          # 'prob error', utils.py:224
    return tok.view(-1, 1)


def max_fn(x: torch.Tensor) -> torch.Tensor:
    """norm(max(x, 0)) with the reference's 1e-6 (utils.py:236-245)."""
    return ops.max_fn(x)
