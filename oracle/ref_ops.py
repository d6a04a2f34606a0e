"""Row-wise CPU restatement of the reference's sampling math (TEST INFRASTRUCTURE ONLY).

Follows, one logits row at a time and with the same ATen op chain:

* ``filter_logits_``  -> /root/reference/sampling/utils.py:152-179  (top_k_top_p_filter)
* ``norm_probs``      -> /root/reference/sampling/utils.py:182-210  (norm_logits)
* ``max_fn``          -> /root/reference/sampling/utils.py:236-245
* ``icdf_sample``     -> replaces ``torch.multinomial`` in utils.py:213-233 (sample) by the
                         explicit inverse-CDF rule of the parity contract (see below); keeps the
                         ``< 1e-9 -> argmax`` guard of utils.py:228-230 and the 'prob error' raise
* ``verify_request``  -> /root/reference/sampling/speculative_sampling.py:1966-2023
                         (accept loop, residual / bonus sample)   [``strict=False``]
                         and :2147-2181 (speculative_sampling_v2)  [``strict=True``]

Deviations from the reference that are part of the parity contract (BASELINE.json north_star):

1. RNG.  ``torch.multinomial`` (an exponential race over V draws, not reproducible from one
   uniform) is replaced by inverse-CDF sampling on ONE pre-drawn uniform ``u`` in [0, 1):

       e      = frexp-exponent of max(probs)            (max = f * 2**e, f in [0.5, 1))
       w_i    = floor(probs_i * 2**(40 - e))            (exact power-of-two scaling, uint64)
       m      = floor(u * 2**24)                        (torch.rand fp32 is a multiple of 2**-24)
       t      = (sum(w) * m) >> 24
       token  = first i with w_0 + ... + w_i > t

   Integer arithmetic makes the rule independent of summation order, so a parallel GPU scan and
   this sequential loop agree bit-for-bit.  Zero-weight entries can never be selected.
2. Tie order.  ``torch.sort(descending=True)`` is unstable on CPU (probed: equal values come
   back in arbitrary index order), so which of several equal logits straddling the top-p cut
   survives is undefined in the reference.  Here (and in the kernels): equal values are ordered
   by ascending vocabulary index.
3. Logits are upcast to fp32 before anything else (reference Llama does that itself,
   sampling/models/modeling_llama.py:870).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch

SCALE_BITS = 40          # fixed-point fraction bits of the sampling weights
U_BITS = 24              # uniforms are used as 24-bit integers
PROB_GUARD = 1e-9        # utils.py:228 (zero-prob guard)
NEG_INF = float("-inf")


# --------------------------------------------------------------------------- filter / normalise
def filter_logits_(x: torch.Tensor, top_k: Optional[int] = 0, top_p: Optional[float] = 0.0) -> torch.Tensor:
    """In-place top-k then top-p mask on a (rows, V) fp32 tensor.  utils.py:166-178."""
    assert x.dim() == 2
    if top_k is not None and top_k > 0:                                   # utils.py:166-169
        k = min(int(top_k), x.size(-1))
        kth = torch.topk(x, k, dim=-1).values[:, -1:]
        x.masked_fill_(x < kth, NEG_INF)                                  # ties with kth survive
    if top_p is not None and top_p > 0.0:                                 # utils.py:170-178
        srt = torch.sort(x, dim=-1, descending=True, stable=True)         # deviation 2: stable
        cum = torch.cumsum(torch.softmax(srt.values, dim=-1), dim=-1)     # fp64 accumulate on CPU
        drop_sorted = cum > top_p                                         # scalar is cast to fp32
        drop_sorted = torch.cat(
            [torch.zeros_like(drop_sorted[:, :1]), drop_sorted[:, :-1]], dim=-1)  # shift right
        drop = torch.zeros_like(drop_sorted).scatter_(1, srt.indices, drop_sorted)
        x.masked_fill_(drop, NEG_INF)
    return x


def norm_probs(logits: torch.Tensor, temperature: float, top_k: Optional[int], top_p: Optional[float]) -> torch.Tensor:
    """(rows, V) logits -> (rows, V) fp32 probabilities.  utils.py:182-210."""
    assert logits.dim() == 2
    x = logits.to(torch.float32) / temperature                            # utils.py:197 (new tensor)
    x = filter_logits_(x, top_k, top_p)
    probs = torch.log_softmax(x, dim=1).exp()                             # utils.py:199
    if bool(probs.isnan().any()) or bool(probs.isinf().any()) or bool((probs < 0).any()):
        raise RuntimeError("norm logits error")                           # utils.py:203-207
    return probs


def max_fn(x: torch.Tensor) -> torch.Tensor:
    """norm(max(x, 0)) with the reference's +1e-6 in the denominator.  utils.py:236-245."""
    pos = torch.clamp_min(x, 0.0)
    denom = pos.sum(dim=1, keepdim=True) if x.dim() > 1 else pos.sum()
    return pos / (denom + 1e-6)


# --------------------------------------------------------------------------- inverse-CDF rule
def sampling_weights(probs_row: np.ndarray) -> Tuple[np.ndarray, int]:
    """uint64 fixed-point weights of one row and the frexp exponent of its maximum."""
    p = np.ascontiguousarray(probs_row, dtype=np.float32).reshape(-1)
    if not np.all(np.isfinite(p)) or np.any(p < 0):
        raise RuntimeError("prob error")
    mx = float(p.max()) if p.size else 0.0
    if not mx > 0.0:
        raise RuntimeError("prob error")                                   # all-zero row: multinomial raises
    _, e = math.frexp(mx)
    scaled = np.ldexp(p.astype(np.float64), SCALE_BITS - e)                # exact (power of two)
    return np.floor(scaled).astype(np.uint64), e


def u_to_int(u: float) -> int:
    m = int(math.floor(float(np.float32(u)) * (1 << U_BITS)))
    return min(max(m, 0), (1 << U_BITS) - 1)


def icdf_sample(probs_row, u: float, return_margin: bool = False):
    """Token index drawn from one (V,) row of non-negative weights with the uniform ``u``."""
    p = probs_row.detach().cpu().numpy() if isinstance(probs_row, torch.Tensor) else np.asarray(probs_row)
    p = p.reshape(-1).astype(np.float32)
    w, _ = sampling_weights(p)
    csum = np.cumsum(w, dtype=np.uint64)
    total = int(csum[-1])
    t = (total * u_to_int(u)) >> U_BITS
    idx = int(np.searchsorted(csum, np.uint64(t), side="right"))
    if p[idx] < PROB_GUARD:                                                # utils.py:228-230
        idx = int(np.argmax(p))
    if not return_margin:
        return idx
    lo = int(csum[idx - 1]) if idx > 0 else 0
    hi = int(csum[idx])
    margin = min(t - lo, hi - 1 - t) / float(total) if hi > lo else 0.0   # distance to a CDF step
    return idx, margin


def residual_weights(p_row: torch.Tensor, q_row: torch.Tensor) -> torch.Tensor:
    """max(0, p - q) in fp32: the un-normalised residual the kernels sample from."""
    return torch.clamp_min(p_row.float() - q_row.float(), 0.0)


# --------------------------------------------------------------------------- accept / resample
def accept_scan(p_at: np.ndarray, q_at: np.ndarray, u_acc: np.ndarray, strict: bool = False):
    """First-rejection scan.  Returns (n_accepted, ratios fp32, n_exact_ties).

    strict=False: speculative_sampling.py:1975-1990 — reject iff ``r > p/q`` (fp32 divide; the
                  reference divides two Python floats and compares against an fp32 tensor, which
                  rounds the quotient back to fp32 — SURVEY.md §3.2 [probe]).
    strict=True : speculative_sampling.py:2152-2160 (v2) — accept iff ``r < min(1, p/q)``.
    """
    p_at = np.asarray(p_at, dtype=np.float32)
    q_at = np.asarray(q_at, dtype=np.float32)
    u_acc = np.asarray(u_acc, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = (p_at / q_at).astype(np.float32)
    n_acc, ties = len(ratio), 0
    for i in range(len(ratio)):
        if q_at[i] == 0.0:
            raise RuntimeError("s")                                        # ZeroDivisionError -> 's' (:2044-2046)
        thr = min(np.float32(1.0), ratio[i]) if strict else ratio[i]
        if u_acc[i] == thr:
            ties += 1
        ok = (u_acc[i] < thr) if strict else (not (u_acc[i] > thr))
        if not ok:
            n_acc = i
            break
    return n_acc, ratio, ties


def verify_request(p_rows: torch.Tensor, q_rows: torch.Tensor, draft: torch.Tensor,
                   u_acc, u_final: float, strict: bool = False, residual: str = "raw",
                   return_margin: bool = False):
    """One request's verify step.

    p_rows (gamma+1, V) target probs for positions L-1 .. L+gamma-1, q_rows (gamma, V) draft
    probs for positions L-1 .. L+gamma-2, draft (gamma,) drafted token ids.
    residual='raw'        sample from max(0, p-q) directly (the kernels' rule);
    residual='normalised' sample from max_fn(p-q), exactly what the patched reference does
                          (speculative_sampling.py:2007); the two differ only when u lands within
                          ~1e-7 of a CDF step.
    Returns (n_accepted, next_token, ratios, ties[, margin]).
    """
    gamma = q_rows.shape[0]
    d = draft.reshape(-1).tolist()
    p_at = np.array([float(p_rows[i, d[i]]) for i in range(gamma)], dtype=np.float32)
    q_at = np.array([float(q_rows[i, d[i]]) for i in range(gamma)], dtype=np.float32)
    n_acc, ratio, ties = accept_scan(p_at, q_at, u_acc, strict)
    if n_acc < gamma:                                                      # :2005-2015
        res = residual_weights(p_rows[n_acc], q_rows[n_acc])
        if residual == "normalised":
            res = max_fn(res.unsqueeze(0))[0]
        try:
            out = icdf_sample(res, u_final, return_margin=True)
        except RuntimeError:
            if strict:                                                     # v2 re-raises (:2163-2170)
                raise
            out = icdf_sample(p_rows[n_acc].float(), u_final, return_margin=True)   # :2009-2010 fallback
    else:                                                                  # :2016-2023 bonus token
        out = icdf_sample(p_rows[gamma].float(), u_final, return_margin=True)
    tok, margin = out
    if return_margin:
        return n_acc, tok, ratio, ties, margin
    return n_acc, tok, ratio, ties


def verify_multi_request(p_rows: torch.Tensor, q_rows: torch.Tensor, draft: torch.Tensor, u_seq, u_final: float,
                         residual: str = "raw"):
    """One request's verify step of multi_speculative_sampling(strategy='iid'), speculative_sampling.py:1612-1667.

    p_rows (W, gamma+1, V), q_rows (W, gamma, V), draft (W, gamma); u_seq: the accept uniforms IN DRAWING ORDER (draft 0
    until its first reject, then draft 1, ...).  Accept iff r < min(1, p/q) (fp32 divide; a NaN ratio rejects); the first
    draft with the longest accepted run wins, an all-accepted draft ends the scan.  Residual / bonus sampling as in
    verify_request (non-strict: an empty residual falls back to p).  Returns (choice, n_accepted, next_token, ratios)."""
    W, gamma = draft.shape
    n_rand = 0
    max_l, choice, all_acc = 0, 0, False
    ratios = np.zeros((W, gamma), dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        for w in range(W):
            for i in range(gamma):
                j = int(draft[w, i])
                ratios[w, i] = np.float32(p_rows[w, i, j]) / np.float32(q_rows[w, i, j])
    for w in range(W):
        cur_l, cur_all = 0, True
        for i in range(gamma):
            r = np.float32(u_seq[n_rand])
            n_rand += 1
            ratio = ratios[w, i]
            thr = ratio if np.isnan(ratio) else min(np.float32(1.0), ratio)
            if r < thr:
                cur_l += 1
            else:
                cur_all = False
                break
        if cur_l > max_l:
            max_l, choice = cur_l, w
            if cur_all:
                all_acc = True
                break
    if not all_acc:
        res = residual_weights(p_rows[choice, max_l], q_rows[choice, max_l])
        if residual == "normalised":
            res = max_fn(res.unsqueeze(0))[0]
        try:
            tok = icdf_sample(res, u_final)
        except RuntimeError:
            tok = icdf_sample(p_rows[choice, max_l].float(), u_final)
    else:
        tok = icdf_sample(p_rows[choice, gamma].float(), u_final)
    return choice, max_l, tok, ratios
