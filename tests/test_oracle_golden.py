"""CPU tests: the oracle (restatement) against fixtures produced by the UNMODIFIED reference
(tests/golden, written by oracle/make_golden.py) and, when /root/reference is present (dev
container), against the live reference."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ref_loader, ref_ops, replay_model, spec_loop, tape
from oracle.make_golden import NORM_CASES, _logits

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_norm_probs_restatement_matches_reference_vectors():
    blob = np.load(os.path.join(GOLD, "norm_logits.npz"))
    torch.set_num_threads(1)
    for ci, (V, rows, T, k, p, scale, seed, dtype) in enumerate(NORM_CASES):
        x = _logits(V, rows, scale, seed, dtype)
        got = torch.cat([ref_ops.norm_probs(x[i:i + 1].float(), T, k, p) for i in range(rows)], 0)
        nz = torch.from_numpy(blob[f"c{ci}_nz_idx"].astype(np.int64))
        want = torch.zeros(rows, V)
        want[nz[:, 0], nz[:, 1]] = torch.from_numpy(blob[f"c{ci}_nz_val"])
        assert torch.equal(got > 0, want > 0), f"case {ci}: support differs"
        assert torch.allclose(got, want, rtol=1e-6, atol=0), f"case {ci}"


def test_max_fn_restatement_matches_reference_vectors():
    blob = np.load(os.path.join(GOLD, "max_fn.npz"))
    got = ref_ops.max_fn(torch.from_numpy(blob["x"]))
    assert torch.equal(got, torch.from_numpy(blob["y"]))


@pytest.mark.parametrize("residual", ["normalised", "raw"])
def test_spec_loop_restatement_matches_reference_runs(residual):
    """End-to-end sampling.speculative_sampling of the reference (tape-driven) vs the oracle loop.
    'normalised' replays exactly what the reference computes; 'raw' is the rule the kernels implement —
    they may only differ when a uniform lands within ~1e-7 of a CDF step (none in these fixtures)."""
    runs = json.load(open(os.path.join(GOLD, "spec_runs.json")))
    torch.set_num_threads(1)
    for r in runs:
        if r["V"] > 4096 and residual == "raw":
            continue                                    # keep the CPU suite short; covered by 'normalised'
        d, t = replay_model.make_pair(r["V"], seed=r["seed"], noise=r["noise"])
        prefix = torch.tensor([r["prefix"]])
        tp = tape.make_tape(r["seed"], r["max_len"] + 1, r["gamma"])
        out, det = spec_loop.speculative_sampling(prefix, d, t, r["max_len"], r["gamma"], r["temperature"], r["top_k"],
                                                  r["top_p"], tape=tp, residual=residual)
        assert out[0].tolist() == r["tokens"], r
        assert det["acc_len"] == r["acc_len"]
        assert abs(det["acc_rate"] - r["acc_rate"]) < 1e-6


def test_multi_draft_restatement_matches_reference_runs():
    """The reference's multi_speculative_sampling(strategy='iid') (speculative_sampling.py:1379-1716, tape-driven,
    unmodified code) vs the oracle's restatement: tokens, longest accepted run per iteration, mean acceptance."""
    from oracle import make_golden
    runs = json.load(open(os.path.join(GOLD, "multi_runs.json")))
    torch.set_num_threads(1)
    for r in runs:
        if r["V"] > 4096:
            continue                                    # keep the CPU suite short (covered on the GPU side)
        d, t = replay_model.make_pair(r["V"], seed=r["seed"], noise=r["noise"])
        tp = make_golden.multi_tape(r["seed"], r["max_len"] + 1, r["gamma"], r["width"])
        out, det = spec_loop.multi_speculative_sampling(torch.tensor([r["prefix"]]), d, t, r["max_len"], r["gamma"], r["width"],
                                                        r["temperature"], r["top_k"], r["top_p"], tape=tp)
        assert out[0].tolist() == r["tokens"], r
        assert det["acc_len"] == r["acc_len"]
        assert abs(det["acc_rate"] - r["acc_rate"]) < 1e-6


def test_bild_restatement_matches_reference_runs():
    """The reference's BiLD_sampling (speculative_sampling.py:1718-1873, tape-driven, unmodified code) vs the oracle's
    restatement: tokens, accepted run lengths and the number of draft / target calls."""
    runs = json.load(open(os.path.join(GOLD, "bild_runs.json")))
    torch.set_num_threads(1)
    for r in runs:
        if r["V"] > 4096:
            continue                                    # keep the CPU suite short (covered on the GPU side)
        d, t = replay_model.make_pair(r["V"], seed=r["seed"], noise=r["noise"])
        tp = tape.make_tape(r["seed"], r["max_len"] + 1, r["gamma"])
        out, det = spec_loop.bild_sampling(torch.tensor([r["prefix"]]), d, t, r["max_len"], r["gamma"], r["fallback_thres"],
                                           r["rollback_thres"], r["temperature"], r["top_k"], r["top_p"], tape=tp)
        assert out[0].tolist() == r["tokens"], r
        assert det["acc_len"] == r["acc_len"]
        assert (det["target_call_times"], det["approx_call_times"]) == (r["target_call_times"], r["approx_call_times"])


def test_bild_restatement_matches_reference_runs_with_eos_inside_a_draft():
    runs = json.load(open(os.path.join(GOLD, "bild_eos_runs.json")))
    torch.set_num_threads(1)
    for r in runs:
        d, t = replay_model.make_pair(r["V"], seed=r["seed"], noise=r["noise"])
        tp = tape.make_tape(r["seed"], r["max_len"] + 1, r["gamma"])
        out, det = spec_loop.bild_sampling(torch.tensor([r["prefix"]]), d, t, r["max_len"], r["gamma"], r["fallback_thres"],
                                           r["rollback_thres"], r["temperature"], r["top_k"], r["top_p"], eos_token_id=r["eos"], tape=tp)
        assert out[0].tolist() == r["tokens"], r
        assert (det["target_call_times"], det["approx_call_times"]) == (r["target_call_times"], r["approx_call_times"])


@pytest.mark.parametrize("residual", ["normalised", "raw"])
def test_v2_restatement_matches_reference_runs(residual):
    """The reference's speculative_sampling_v2 (speculative_sampling.py:2080-2194: no KV cache, strict accept test,
    tape-driven, unmodified code) vs the oracle's restatement: tokens, accepted run per iteration, acc_rate."""
    runs = json.load(open(os.path.join(GOLD, "v2_runs.json")))
    torch.set_num_threads(1)
    for r in runs:
        if r["V"] > 4096:
            continue                                    # keep the CPU suite short (the GPU tests cover V = 32000)
        d, t = replay_model.make_pair(r["V"], seed=r["seed"], noise=r["noise"])
        tp = tape.make_tape(r["seed"], r["max_len"] + 1, r["gamma"])
        out, det = spec_loop.speculative_sampling_v2(torch.tensor([r["prefix"]]), d, t, r["max_len"], r["gamma"],
                                                     r["temperature"], r["top_k"], r["top_p"], tape=tp, residual=residual)
        assert out[0].tolist() == r["tokens"], r
        assert det["acc_len"] == r["acc_len"]
        assert abs(det["acc_rate"] - r["acc_rate"]) < 1e-6


def test_autoregressive_restatement_matches_reference_runs():
    """The reference's autoregressive_sampling (autoregressive_sampling.py:9-61, its sample fed from one uniform per
    token, unmodified code) vs the oracle's restatement, including the EOS stop (:55)."""
    from oracle import make_golden
    runs = json.load(open(os.path.join(GOLD, "ar_runs.json")))
    torch.set_num_threads(1)
    assert any(r["eos"] is not None and len(r["tokens"]) < len(r["prefix"]) + r["N"] for r in runs), "no run stops at EOS"
    for r in runs:
        if r["V"] > 4096:
            continue
        _, t = replay_model.make_pair(r["V"], seed=r["seed"], noise=0.5)
        out = spec_loop.autoregressive_sampling(torch.tensor([r["prefix"]]), t, r["N"], r["eos"], r["temperature"], r["top_k"],
                                                r["top_p"], uniforms=make_golden.ar_uniforms(r["seed"], r["N"]))
        assert out[0].tolist() == r["tokens"], r


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference only exists in the dev container")
def test_v2_and_autoregressive_oracle_match_live_reference():
    torch.set_num_threads(1)
    d, t = replay_model.make_pair(640, seed=47, noise=0.6)
    prefix = torch.randint(3, 640, (1, 6), generator=torch.Generator().manual_seed(47))
    for (T, k, p, gamma) in [(0.9, 20, 0.9, 4), (1.0, 0, 0.0, 3)]:
        tp = tape.make_tape(47, 25, gamma)
        a, da = ref_loader.run_reference_v2(prefix, d, t, 24, gamma, T, k, p, tape=tp)
        b, db = spec_loop.speculative_sampling_v2(prefix, d, t, 24, gamma, T, k, p, tape=tp, residual="normalised")
        assert torch.equal(a, b) and [int(x) for x in da["acc_len"]] == db["acc_len"]
        assert abs(float(da["acc_rate"]) - db["acc_rate"]) < 1e-6
        u = torch.rand(20, generator=torch.Generator().manual_seed(48))
        assert torch.equal(ref_loader.run_reference_ar(prefix, t, 20, T, k, p, u),
                           spec_loop.autoregressive_sampling(prefix, t, 20, None, T, k, p, uniforms=u))


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference only exists in the dev container")
def test_bild_oracle_matches_live_reference():
    torch.set_num_threads(1)
    d, t = replay_model.make_pair(640, seed=44, noise=0.6)
    prefix = torch.randint(3, 640, (1, 6), generator=torch.Generator().manual_seed(44))
    for (fb, rb, k, p) in [(0.5, 2.5, 20, 0.9), (0.25, 3.0, 0, 0.0)]:
        tp = tape.make_tape(44, 31, 4)
        a, da = ref_loader.run_reference_bild(prefix, d, t, 30, 4, fb, rb, 1.0, k, p, tape=tp)
        b, db = spec_loop.bild_sampling(prefix, d, t, 30, 4, fb, rb, 1.0, k, p, tape=tp)
        assert torch.equal(a, b) and [int(x) for x in da["acc_len"]] == db["acc_len"]


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference only exists in the dev container")
def test_oracle_matches_live_reference():
    torch.set_num_threads(1)
    utils = ref_loader.load_utils()
    g = torch.Generator().manual_seed(123)
    for (V, T, k, p) in [(257, 0.9, 7, 0.6), (1031, 1.0, 0, 0.95), (640, 1.7, 33, 0.0)]:
        x = torch.randn(3, V, generator=g) * 2.5
        a = torch.cat([utils.norm_logits(x[i:i + 1].clone(), T, k, p) for i in range(3)])
        b = ref_ops.norm_probs(x, T, k, p)
        assert torch.allclose(a, b, rtol=1e-6, atol=0) and torch.equal(a > 0, b > 0)
        f1 = utils.top_k_top_p_filter((x / T).clone(), k, p)
        f2 = ref_ops.filter_logits_((x / T).clone(), k, p)
        assert torch.equal(f1, f2)
        assert torch.equal(utils.max_fn(x), ref_ops.max_fn(x))
    d, t = replay_model.make_pair(600, seed=21, noise=0.4)
    prefix = torch.randint(3, 600, (1, 5), generator=g)
    tp = tape.make_tape(5, 40, 4)
    out_ref, det_ref = ref_loader.run_reference(prefix, d, t, 30, 4, 0.9, 10, 0.9, tape=tp)
    out_o, det_o = spec_loop.speculative_sampling(prefix, d, t, 30, 4, 0.9, 10, 0.9, tape=tp, residual="normalised")
    assert torch.equal(out_ref, out_o) and det_ref["acc_len"] == det_o["acc_len"]


def test_icdf_rule_properties():
    g = torch.Generator().manual_seed(0)
    p = torch.rand(500, generator=g) ** 6
    p[::7] = 0
    us = torch.linspace(0, 1 - 2 ** -24, 400)
    toks = [ref_ops.icdf_sample(p, float(u)) for u in us]
    assert all(p[t] > 0 for t in toks)                       # zero-weight entries are never drawn
    assert toks == sorted(toks)                              # monotone in u
    assert toks[0] == int((p > 0).nonzero()[0]) and toks[-1] == int((p > 0).nonzero()[-1])
    assert ref_ops.icdf_sample(p * 1024, 0.3) == ref_ops.icdf_sample(p, 0.3)   # power-of-two scale invariance
    with pytest.raises(RuntimeError, match="prob error"):
        ref_ops.icdf_sample(torch.zeros(10), 0.5)
    # empirical frequencies follow the weights (chi-square style bound)
    q = torch.tensor([0.4, 0.0, 0.6])
    draws = [ref_ops.icdf_sample(q, float(u)) for u in torch.rand(20000, generator=g)]
    f0 = draws.count(0) / len(draws)
    assert draws.count(1) == 0 and abs(f0 - 0.4) < 0.015


def test_accept_rule_variants():
    p_at = np.array([0.5, 0.2, 0.3], dtype=np.float32)
    q_at = np.array([0.5, 0.4, 0.1], dtype=np.float32)
    n, ratio, ties = ref_ops.accept_scan(p_at, q_at, np.array([1.0, 0.5, 0.9], dtype=np.float32), strict=False)
    assert (n, ties) == (3, 2)                               # u == p/q is accepted by `not (u > ratio)`
    n, _, _ = ref_ops.accept_scan(p_at, q_at, np.array([0.99, 0.5, 0.9], dtype=np.float32), strict=True)
    assert n == 1                                            # strict: u < min(1, ratio) fails on the tie
    with pytest.raises(RuntimeError, match="^s$"):
        ref_ops.accept_scan(p_at, np.zeros(3, dtype=np.float32), np.zeros(3, dtype=np.float32))
