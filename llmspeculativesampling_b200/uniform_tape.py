"""Pre-drawn uniform tape of the speculative loop (product side).

One iteration of one request consumes a block of 2*gamma+2 fp32 uniforms:
    [0, gamma)           draft samples            (reference sampling/kvcache_model.py:283)
    [gamma]              the sample the reference draws and discards in target.generate(x, 1)
                         (sampling/speculative_sampling.py:1956) — kept so that tapes line up
    [gamma+1, 2gamma+1)  accept tests             (speculative_sampling.py:1978)
    [2gamma+1]           residual / bonus sample  (speculative_sampling.py:2007 / :2019)
Request r uses torch.Generator(seed_of(base_seed, r)): a request's tokens do not depend on which
GPU or batch slot serves it, so sharded runs reproduce the single-GPU run bit for bit.
"""
from __future__ import annotations

from typing import Sequence

import torch


def seed_of(base_seed: int, request_id: int) -> int:
    return (int(base_seed) * 1000003 + int(request_id) * 7919 + 12345) & 0x7FFFFFFF


def block(gamma: int) -> int:
    return 2 * gamma + 2


def make_tape(seed: int, iterations: int, gamma: int) -> torch.Tensor:
    """(iterations, 2*gamma+2) fp32 uniforms in [0, 1), drawn on the CPU (multiples of 2**-24)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return torch.rand(iterations, block(gamma), generator=g, dtype=torch.float32)


def batch_tape(base_seed: int, request_ids: Sequence[int], iterations: int, gamma: int) -> torch.Tensor:
    """(iterations, B, 2*gamma+2): per-request tapes stacked along the batch dimension."""
    return torch.stack([make_tape(seed_of(base_seed, r), iterations, gamma) for r in request_ids], dim=1)
