"""Shared helpers for the parity tests (oracle = checker, kernels = thing under test)."""
import numpy as np
import torch

from oracle import ref_ops

ATOL = 2.0 ** -40        # probabilities below the sampling resolution are irrelevant
RTOL = 1e-5              # BASELINE.json north_star: probabilities within 1e-5 relative (fp32)


def make_logits(rows, V, scale, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(rows, V, generator=g) * scale).to(dtype)


def oracle_probs(logits_cpu, T, k, p):
    torch.set_num_threads(1)
    return torch.cat([ref_ops.norm_probs(logits_cpu[i:i + 1].float(), T, k, p) for i in range(logits_cpu.shape[0])], 0)


def compare_probs(got, want, what=""):
    """Returns the number of rows whose support differs (boundary ties); asserts values elsewhere."""
    got = got.detach().cpu().double()
    want = want.detach().cpu().double()
    assert got.shape == want.shape, what
    boundary_rows = 0
    for r in range(got.shape[0]):
        sg, sw = got[r] > 0, want[r] > 0
        if not torch.equal(sg, sw):
            boundary_rows += 1
            continue
        err = (got[r] - want[r]).abs()
        tol = RTOL * want[r].abs() + ATOL
        bad = err > tol
        assert not bool(bad.any()), (f"{what} row {r}: {int(bad.sum())} entries off, max rel "
                                     f"{float((err / want[r].clamp_min(1e-30))[sw].max()):.3e}")
    return boundary_rows
