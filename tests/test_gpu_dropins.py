"""GPU parity of the drop-in API itself (the functions a user of the reference calls) against fixtures generated
from the UNMODIFIED reference by oracle/make_golden.py:
  sampling.utils.{norm_logits, top_k_top_p_filter, sample, max_fn}   utils.py:152-245    norm_logits.npz, max_fn.npz
  sampling.speculative_sampling_v2                                   speculative_sampling.py:2080-2194   v2_runs.json
  sampling.autoregressive_sampling                                   autoregressive_sampling.py:9-61     ar_runs.json
Token ids and accept counts are bit-exact; probabilities within 1e-5 relative (helpers.RTOL)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ref_ops, replay_model, tape
from tests.helpers import compare_probs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _golden_norm_cases():
    from oracle.make_golden import NORM_CASES, _logits
    blob = np.load(os.path.join(GOLD, "norm_logits.npz"))
    for ci, (V, rows, T, k, p, scale, seed, dtype) in enumerate(NORM_CASES):
        x = _logits(V, rows, scale, seed, dtype)
        want = torch.zeros(rows, V)
        nz = torch.from_numpy(blob[f"c{ci}_nz_idx"].astype(np.int64))
        want[nz[:, 0], nz[:, 1]] = torch.from_numpy(blob[f"c{ci}_nz_val"])
        yield ci, x, T, k, p, want


def test_utils_norm_logits_and_filter_match_reference_vectors(cuda_lib):
    from llmspeculativesampling_b200.sampling import utils
    for ci, x, T, k, p, want in _golden_norm_cases():
        xg = x.cuda()
        before = xg.clone()
        got = utils.norm_logits(xg, T, k, p)
        assert torch.equal(xg, before), "norm_logits must not modify the caller's logits (utils.py:197)"
        assert got.dtype == torch.float32 and got.shape == want.shape
        assert compare_probs(got, want, f"utils.norm_logits golden case {ci}") == 0
        # top_k_top_p_filter acts IN PLACE on logits / T and returns the same tensor (utils.py:152-179): the finite
        # entries are exactly the support of the reference's probabilities and keep their values
        scaled = (x.float() / T).cuda()
        keep_vals = scaled.clone()
        ret = utils.top_k_top_p_filter(scaled, k, p)
        assert ret.data_ptr() == scaled.data_ptr()
        finite = torch.isfinite(scaled).cpu()
        assert torch.equal(finite, want > 0), f"filter support differs in golden case {ci}"
        assert torch.equal(scaled[finite.cuda()], keep_vals[finite.cuda()])
        assert bool((scaled[~finite.cuda()] == float("-inf")).all())


def test_utils_sample_and_max_fn_match_reference(cuda_lib):
    from llmspeculativesampling_b200.sampling import utils
    blob = np.load(os.path.join(GOLD, "max_fn.npz"))
    x, y = torch.from_numpy(blob["x"]), torch.from_numpy(blob["y"])
    got = utils.max_fn(x.cuda())
    assert torch.allclose(got.cpu(), y, rtol=1e-5, atol=1e-12) and float(got[5].abs().sum()) == 0.0
    # sample(): (rows, V) -> (rows, 1) int64, the inverse-CDF token of the row's uniform (the parity contract's
    # replacement for torch.multinomial, utils.py:221), on the reference's own probability vectors
    g = torch.Generator().manual_seed(5)
    for ci, _, _, _, _, want in _golden_norm_cases():
        u = torch.rand(want.shape[0], generator=g)
        tok = utils.sample(want.cuda(), u=u.cuda())
        assert tok.shape == (want.shape[0], 1) and tok.dtype == torch.int64
        assert tok.view(-1).tolist() == [ref_ops.icdf_sample(want[r], float(u[r])) for r in range(want.shape[0])], ci
    # residual sampling exactly as the reference composes it: sample(max_fn(p - q)) (speculative_sampling.py:2007)
    p_row, q_row = y[0:1], y[1:2]
    tok = utils.sample(utils.max_fn((p_row - q_row).cuda()), u=torch.tensor([0.37], device="cuda"))
    assert int(tok) == ref_ops.icdf_sample(ref_ops.max_fn(p_row - q_row)[0], 0.37)
    with pytest.raises(RuntimeError, match="prob error"):                  # utils.py:224
        utils.sample(torch.zeros(2, 100, device="cuda"), u=torch.tensor([0.1, 0.2], device="cuda"))
    bad = torch.randn(2, 1000, device="cuda")
    bad[1, 17] = float("nan")
    with pytest.raises(RuntimeError, match="norm logits error"):           # utils.py:207
        utils.norm_logits(bad, 1.0, 20, 0.9)
    with pytest.raises(AssertionError):                                    # utils.py:194
        utils.norm_logits(torch.randn(1000, device="cuda"), 1.0, 20, 0.9)
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        utils.norm_logits(torch.randn(2, 1000), 1.0, 20, 0.9)


def test_speculative_sampling_v2_matches_reference_golden_runs(cuda_lib):
    """speculative_sampling_v2 on the KV-cached engine vs runs of the UNMODIFIED reference v2 (full re-forwards, strict
    accept test u < min(1, p/q)): same tokens, accepted runs and acceptance statistic."""
    from llmspeculativesampling_b200.sampling import speculative_sampling_v2
    runs = json.load(open(os.path.join(GOLD, "v2_runs.json")))
    assert len(runs) >= 6
    for r in runs:
        d, t = replay_model.make_pair(r["V"], seed=r["seed"], noise=r["noise"], device="cuda")
        prefix = torch.tensor([r["prefix"]], device="cuda")
        tp = tape.make_tape(r["seed"], r["max_len"] + 1, r["gamma"]).unsqueeze(1)
        out, det = speculative_sampling_v2(prefix, d, t, r["max_len"], r["gamma"], r["temperature"], r["top_k"], r["top_p"],
                                           details=True, uniforms=tp)
        assert out[0].tolist() == r["tokens"], f"V={r['V']} k={r['top_k']} p={r['top_p']} gamma={r['gamma']}"
        assert det["acc_len"] == r["acc_len"]
        # the reference's v2 statistic is lazy: min(1, p/q) of the tested tokens only (:2155)
        assert abs(det["acc_rate"] - r["acc_rate"]) < 1e-5


def test_autoregressive_sampling_matches_reference_golden_runs(cuda_lib):
    from oracle.make_golden import ar_uniforms
    from llmspeculativesampling_b200.sampling import autoregressive_sampling
    runs = json.load(open(os.path.join(GOLD, "ar_runs.json")))
    assert len(runs) >= 6
    for r in runs:
        _, t = replay_model.make_pair(r["V"], seed=r["seed"], noise=0.5, device="cuda")
        x = torch.tensor([r["prefix"]], device="cuda")
        out = autoregressive_sampling(x, t, r["N"], r["eos"], r["temperature"], r["top_k"], r["top_p"],
                                      uniforms=ar_uniforms(r["seed"], r["N"]).view(-1, 1))
        assert out[0].tolist() == r["tokens"], f"V={r['V']} k={r['top_k']} p={r['top_p']} eos={r['eos']}"
