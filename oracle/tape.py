"""Uniform tape for the parity contract (TEST INFRASTRUCTURE ONLY).

One speculative iteration of one request consumes a fixed block of ``2*gamma + 2`` fp32
uniforms, all drawn up-front from ``torch.Generator(seed)`` on the CPU:

    [0 .. gamma)            u_draft   one per drafted token  (kvcache_model.py:283 sample(q))
    [gamma]                 u_discard the sample() whose result the reference throws away in
                                      target.generate(x, 1)  (speculative_sampling.py:1956)
    [gamma+1 .. 2gamma+1)   u_acc     accept tests (the reference draws lazily and stops at the
                                      first reject, :1978; the tape draws all gamma, uses a prefix)
    [2gamma+1]              u_final   residual / bonus sample (:2007 / :2019)

Request ``r`` of a batch uses the generator seeded with ``seed_of(base_seed, r)`` so that shards
of a multi-GPU run reproduce the single-GPU tokens bit-for-bit.
"""
from __future__ import annotations

import torch


def seed_of(base_seed: int, request_id: int) -> int:
    return (int(base_seed) * 1000003 + int(request_id) * 7919 + 12345) & 0x7FFFFFFF


def block(gamma: int) -> int:
    return 2 * gamma + 2


def make_tape(seed: int, iterations: int, gamma: int) -> torch.Tensor:
    """(iterations, 2*gamma+2) fp32 uniforms in [0,1), multiples of 2**-24."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return torch.rand(iterations, block(gamma), generator=g, dtype=torch.float32)


def split(row: torch.Tensor, gamma: int):
    """row (2gamma+2,) -> (u_draft[gamma], u_discard, u_acc[gamma], u_final)."""
    return row[:gamma], row[gamma], row[gamma + 1:2 * gamma + 1], row[2 * gamma + 1]
