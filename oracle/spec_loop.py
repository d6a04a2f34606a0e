"""Batch-1 restatement of the reference's KV-cached stepping and speculative loop
(TEST INFRASTRUCTURE ONLY).

* ``OracleStepper``           -> /root/reference/sampling/kvcache_model.py:23-36 (state),
                                 :141-252 (_forward_with_kvcache, decoder-only branch),
                                 :255-310 (generate, multi=1), :360-384,429-431 (rollback, choice=None)
* ``speculative_sampling``    -> /root/reference/sampling/speculative_sampling.py:1934-2043
* ``speculative_sampling_v2`` -> /root/reference/sampling/speculative_sampling.py:2118-2185
* ``autoregressive_sampling`` -> /root/reference/sampling/autoregressive_sampling.py:9-61
* ``multi_speculative_sampling`` -> /root/reference/sampling/speculative_sampling.py:1379-1716 (strategy='iid'),
                                 with kvcache_model.py:180-200,272-276 (multi) and :390-396,433-436 (rollback choice)
* ``bild_sampling``           -> /root/reference/sampling/speculative_sampling.py:1718-1873 (BiLD_sampling, decoder-only)

Randomness comes from a uniform tape (oracle/tape.py) instead of torch's global RNG.
Works with ``oracle.replay_model.ReplayLM`` (legacy tuple cache) and with stock Hugging Face
causal LMs (``DynamicCache``; cropped with ``.crop``).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import ref_ops, tape as tape_mod


def _cached_len(kv) -> int:
    if hasattr(kv, "get_seq_length"):
        return int(kv.get_seq_length())
    return kv[0][0].shape[2]                                                # kvcache_model.py:175


def _crop(kv, end_pos: int):
    if hasattr(kv, "crop"):
        if _cached_len(kv) > end_pos:
            kv.crop(end_pos)
        return kv
    return tuple((k[:, :, :end_pos, :], v[:, :, :end_pos, :]) for k, v in kv)   # :381-382


class OracleStepper:
    def __init__(self, model, temperature: float = 1.0, top_k: int = 0, top_p: float = 0.0):
        self.model = model
        self.kv = None
        self.hist: Optional[torch.Tensor] = None                            # (seq, V) fp32 probs
        self.temperature, self.top_k, self.top_p = temperature, top_k, top_p

    @torch.no_grad()
    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        if self.kv is None:                                                 # prefill, :151-171
            out = self.model(ids, use_cache=True)
        else:                                                               # incremental, :173-214
            out = self.model(ids[:, _cached_len(self.kv):], past_key_values=self.kv, use_cache=True)
        logits = out.logits[0].to(torch.float32)
        rows = [ref_ops.norm_probs(logits[i:i + 1], self.temperature, self.top_k, self.top_p)
                for i in range(logits.shape[0])]                            # one row per call, :166-168/:235-236
        new = torch.cat(rows, dim=0)
        self.hist = new if self.hist is None else torch.cat([self.hist, new], dim=0)   # :246
        self.kv = out.past_key_values
        return new[-1]

    def generate(self, ids: torch.Tensor, gamma: int, uniforms) -> torch.Tensor:
        x = ids
        for i in range(gamma):                                              # :279-293
            q = self.forward(x)
            tok = ref_ops.icdf_sample(q, float(uniforms[i]))
            x = torch.cat([x, torch.tensor([[tok]], dtype=x.dtype, device=x.device)], dim=1)
        return x

    def rollback(self, end_pos: int) -> None:                               # :360-431
        self.kv = _crop(self.kv, end_pos)
        if self.hist is not None:
            self.hist = self.hist[:end_pos]


@torch.no_grad()
def speculative_sampling(prefix: torch.Tensor, approx_model, target_model, max_len: int, gamma: int = 4,
                         temperature: float = 1.0, top_k: int = 0, top_p: float = 0.0,
                         eos_token_id: Optional[int] = None, tape: Optional[torch.Tensor] = None,
                         seed: int = 0, residual: str = "raw", record: Optional[list] = None):
    """Returns (tokens (1, n), details).  One tape row (2*gamma+2 uniforms) per iteration."""
    assert prefix.shape[0] == 1, "input batch size must be 1"               # :1905
    seq_len = prefix.shape[1]
    T = seq_len + max_len
    if tape is None:
        tape = tape_mod.make_tape(seed, max_len + 1, gamma)
    ori_eos = int((prefix == eos_token_id).sum()) if eos_token_id is not None else 0
    approx = OracleStepper(approx_model, temperature, top_k, top_p)
    target = OracleStepper(target_model, temperature, top_k, top_p)
    acc_len: List[int] = []
    acc_rate: List[float] = []
    ties = 0
    min_margin = 1.0
    it = 0
    out = prefix
    while prefix.shape[1] < T:                                              # :1934
        u_draft, u_discard, u_acc, u_final = tape_mod.split(tape[it], gamma)
        x = approx.generate(prefix, gamma, u_draft)                         # :1943
        L = prefix.shape[1]
        _ = target.generate(x, 1, [u_discard])                              # :1956 (sample discarded)
        p_rows = target.hist[L - 1:L + gamma]                               # gamma+1 rows
        q_rows = approx.hist[L - 1:L + gamma - 1]                           # gamma rows
        draft = x[0, L:L + gamma]
        n_acc, tok, ratio, t, margin = ref_ops.verify_request(
            p_rows, q_rows, draft, u_acc.numpy(), float(u_final), strict=False,
            residual=residual, return_margin=True)
        ties += t
        min_margin = min(min_margin, margin)
        acc_rate.extend(np.minimum(1.0, ratio.astype(np.float64)).tolist())  # :1966-1971
        acc_len.append(n_acc)                                               # :1991
        if record is not None:
            record.append(dict(p=p_rows.clone(), q=q_rows.clone(), draft=draft.clone(),
                               u_acc=u_acc.clone(), u_final=float(u_final), n_acc=n_acc, tok=tok))
        n = L + n_acc - 1
        prefix = x[:, :n + 1]                                               # :1996
        approx.rollback(n + 1)                                              # :2000
        target.rollback(n + 1 if n_acc < gamma else n + 2)                  # :2015 / :2023
        prefix = torch.cat([prefix, torch.tensor([[tok]], dtype=prefix.dtype)], dim=1)   # :2027
        out = prefix
        it += 1
        if eos_token_id is not None:                                        # :2033-2041
            mask = out == eos_token_id
            if int(mask.sum()) > ori_eos:
                keep = torch.cumsum(mask.float(), dim=1) < ori_eos + 1
                end = int(keep.sum())
                if end < keep.shape[1]:
                    keep[:, end] = True
                out = out[keep][None, :]
                break
    details = dict(acc_len=acc_len, acc_rate=float(np.mean(acc_rate)) if acc_rate else 0.0,
                   iterations=it, target_call_times=it, approx_call_times=it,
                   exact_ties=ties, min_sample_margin=min_margin)
    return out, details


@torch.no_grad()
def speculative_sampling_v2(prefix: torch.Tensor, approx_model, target_model, max_len: int, gamma: int = 4,
                            temperature: float = 1.0, top_k: int = 0, top_p: float = 0.0,
                            tape: Optional[torch.Tensor] = None, seed: int = 0, residual: str = "raw"):
    """DeepMind variant without KV cache (full re-forward), strict accept test.  Tape layout as
    above except that slot ``gamma`` (u_discard) is unused."""
    assert prefix.shape[0] == 1
    T = prefix.shape[1] + max_len
    if tape is None:
        tape = tape_mod.make_tape(seed, max_len + 1, gamma)
    acc_len, acc_rate, it = [], [], 0

    def rows_of(model, ids):
        lg = model(ids).logits[0].to(torch.float32)
        return torch.cat([ref_ops.norm_probs(lg[i:i + 1], temperature, top_k, top_p) for i in range(lg.shape[0])], 0)

    while prefix.shape[1] < T:                                              # :2118
        u_draft, _, u_acc, u_final = tape_mod.split(tape[it], gamma)
        x, L = prefix, prefix.shape[1]
        for i in range(gamma):                                              # :2123-2128
            q_last = rows_of(approx_model, x)[-1]
            tok = ref_ops.icdf_sample(q_last, float(u_draft[i]))
            x = torch.cat([x, torch.tensor([[tok]], dtype=x.dtype)], dim=1)
        q = rows_of(approx_model, x[:, :-1])                                # rows of the last draft forward, :2131-2133
        p = rows_of(target_model, x)                                        # :2137-2140
        n_acc, tok, ratio, _ = ref_ops.verify_request(p[L - 1:L + gamma], q[L - 1:L + gamma - 1], x[0, L:L + gamma],
                                                      u_acc.numpy(), float(u_final), strict=True, residual=residual)
        acc_rate.extend(np.minimum(1.0, ratio.astype(np.float64))[:min(n_acc + 1, gamma)].tolist())   # :2155 (lazy)
        acc_len.append(n_acc)
        prefix = torch.cat([x[:, :L + n_acc], torch.tensor([[tok]], dtype=x.dtype)], dim=1)
        it += 1
    return prefix, dict(acc_len=acc_len, acc_rate=float(np.mean(acc_rate)) if acc_rate else 0.0, iterations=it)


@torch.no_grad()
def autoregressive_sampling(x: torch.Tensor, model, N: int, eos_token_id: Optional[int] = None,
                            temperature: float = 1.0, top_k: int = 0, top_p: float = 0.0,
                            uniforms: Optional[torch.Tensor] = None, seed: int = 0) -> torch.Tensor:
    """Target-only loop, one uniform per generated token.  autoregressive_sampling.py:9-61."""
    if uniforms is None:
        uniforms = torch.rand(N, generator=torch.Generator().manual_seed(int(seed)))
    kv = None
    for i in range(N):                                                      # n = len(x) = 1 -> exactly N tokens (:13-21)
        out = model(x, use_cache=True) if kv is None else model(x[:, -1:], past_key_values=kv, use_cache=True)
        kv = out.past_key_values
        probs = ref_ops.norm_probs(out.logits[:, -1, :].to(torch.float32), temperature, top_k, top_p)
        tok = ref_ops.icdf_sample(probs[0], float(uniforms[i]))
        x = torch.cat([x, torch.tensor([[tok]], dtype=x.dtype)], dim=1)
        if eos_token_id is not None and tok == eos_token_id:
            break
    return x


@torch.no_grad()
def bild_sampling(prefix: torch.Tensor, approx_model, target_model, max_len: int, gamma: int, fallback_thres: float,
                  rollback_thres: float, temperature: float = 1.0, top_k: int = 0, top_p: float = 0.0,
                  eos_token_id: Optional[int] = None, tape: Optional[torch.Tensor] = None, seed: int = 0):
    """Big-Little decoding, speculative_sampling.py:1718-1873 (decoder-only branch): the draft model emits one token at
    a time; the target only looks when the draft is unsure (max q < fallback_thres, :1784) or gamma tokens are unchecked;
    it keeps tokens while -log p[token] <= rollback_thres (:1800), then ALWAYS samples its own next token (:1812).
    One tape row per check cycle: u_draft[i] for the i-th draft token of the cycle, u_discard for the sample that
    target.generate(x, 1) throws away (:1788), u_final for the target's token (:1812).
    Returns (tokens (1, n), details)."""
    assert prefix.shape[0] == 1, "input batch size must be 1"               # :1729
    seq_len = prefix.shape[1]
    T = seq_len + max_len
    if tape is None:
        tape = tape_mod.make_tape(seed, max_len + 1, gamma)
    ori_eos = int((prefix == eos_token_id).sum()) if eos_token_id is not None else 0
    approx = OracleStepper(approx_model, temperature, top_k, top_p)
    target = OracleStepper(target_model, temperature, top_k, top_p)
    acc_len: List[int] = []
    last_check = seq_len - 1                                                # :1759
    cycle, n_draft = 0, 0
    approx_calls = target_calls = 0
    out = prefix
    while prefix.shape[1] < T:                                              # :1764
        u_draft, u_discard, _, u_final = tape_mod.split(tape[cycle], gamma)
        x = approx.generate(prefix, 1, [u_draft[n_draft]])                  # :1772
        n_draft += 1
        approx_calls += 1
        q_last = approx.hist[-1]                                            # :1778 (q[:, -1, :])
        if float(q_last.max()) < fallback_thres or x.shape[1] - last_check - 1 >= gamma:   # :1784
            _ = target.generate(x, 1, [u_discard])                          # :1788
            target_calls += 1
            p = target.hist
            n = x.shape[1] - 1
            l = 0
            for i in range(last_check, x.shape[1] - 1):                     # :1797-1803
                j = int(x[0, i + 1])
                if float(-p[i, j].log()) > rollback_thres:
                    n = i
                    break
                l += 1
            acc_len.append(l)
            prefix = x[:, :n + 1]                                           # :1806
            approx.rollback(n + 1)                                          # :1811
            tok = ref_ops.icdf_sample(p[n], float(u_final))                 # :1812
            target.rollback(n + 1)                                          # :1813
            last_check = n + 1
            prefix = torch.cat([prefix, torch.tensor([[tok]], dtype=prefix.dtype)], dim=1)   # :1817
            cycle += 1
            n_draft = 0
        else:
            prefix = x                                                      # :1826
        out = prefix
        if eos_token_id is not None:                                        # :1833-1841
            mask = out == eos_token_id
            if int(mask.sum()) > ori_eos:
                keep = torch.cumsum(mask.float(), dim=1) < ori_eos + 1
                end = int(keep.sum())
                if end < keep.shape[1]:
                    keep[:, end] = True
                out = out[keep][None, :]
                break
    details = dict(acc_len=acc_len, target_call_times=target_calls, approx_call_times=approx_calls, cycles=cycle)
    return out, details


class MultiStepper:
    """KVCacheModel with a batch of W drafts (legacy tuple caches): kvcache_model.py:141-252 incl. the cache / history
    expansion to W rows (:180-200, :240-245) and rollback(end_pos, choice) (:390-396, :433-436)."""

    def __init__(self, model, temperature: float = 1.0, top_k: int = 0, top_p: float = 0.0):
        self.model = model
        self.kv = None
        self.hist: Optional[torch.Tensor] = None                            # (B, seq, V)
        self.temperature, self.top_k, self.top_p = temperature, top_k, top_p

    def _norm(self, logits: torch.Tensor) -> torch.Tensor:                  # (B, n, V) -> probabilities, one row per call
        B, n, _ = logits.shape
        return torch.stack([torch.cat([ref_ops.norm_probs(logits[b, i:i + 1].float(), self.temperature, self.top_k, self.top_p)
                                       for i in range(n)], 0) for b in range(B)], 0)

    @torch.no_grad()
    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        B = ids.shape[0]
        if self.kv is None:
            out = self.model(ids, use_cache=True)
            self.hist = self._norm(out.logits)
        else:
            if self.kv[0][0].shape[0] < B:                                  # :180-190
                self.kv = tuple((k.repeat(B, 1, 1, 1), v.repeat(B, 1, 1, 1)) for k, v in self.kv)
            out = self.model(ids[:, _cached_len(self.kv):], past_key_values=self.kv, use_cache=True)
            new = self._norm(out.logits)
            if self.hist.shape[0] < B:                                      # :240-244
                self.hist = self.hist.repeat(B // self.hist.shape[0], 1, 1)
            self.hist = torch.cat([self.hist, new], dim=1)
        self.kv = out.past_key_values
        return self.hist[:, -1]

    def generate(self, ids: torch.Tensor, gamma: int, uniforms) -> torch.Tensor:
        """uniforms[i][b]: the uniform of row b's i-th token."""
        x = ids
        for i in range(gamma):
            q = self.forward(x)
            toks = [ref_ops.icdf_sample(q[b], float(uniforms[i][b])) for b in range(x.shape[0])]
            x = torch.cat([x, torch.tensor(toks, dtype=x.dtype).view(-1, 1)], dim=1)
        return x

    def rollback(self, end_pos: int, choice: int) -> None:
        self.kv = tuple((k[choice:choice + 1, :, :end_pos, :], v[choice:choice + 1, :, :end_pos, :]) for k, v in self.kv)
        self.hist = self.hist[choice:choice + 1, :end_pos]


def multi_block(gamma: int, width: int) -> int:
    return 2 * width * gamma + width + 1


@torch.no_grad()
def multi_speculative_sampling(prefix: torch.Tensor, approx_model, target_model, max_len: int, gamma: int, width: int,
                               temperature: float = 1.0, top_k: int = 0, top_p: float = 0.0,
                               eos_token_id: Optional[int] = None, tape: Optional[torch.Tensor] = None):
    """W independent drafts per iteration, one target pass over all of them, the first draft with the longest accepted
    run wins (speculative_sampling.py:1612-1640, accept iff r < min(1, p/q)).  One tape row of multi_block(gamma, W)
    uniforms per iteration: [gamma x W draft | W discarded target samples | accept tests in the order the reference
    draws them | final].  Returns (tokens (1, n), details)."""
    assert prefix.shape[0] == 1, "input batch size must be 1"
    W, g = width, gamma
    T = prefix.shape[1] + max_len
    ori_eos = int((prefix == eos_token_id).sum()) if eos_token_id is not None else 0
    approx = MultiStepper(approx_model, temperature, top_k, top_p)
    target = MultiStepper(target_model, temperature, top_k, top_p)
    acc_len: List[int] = []
    acc_rate: List[float] = []
    choices: List[int] = []
    out = prefix
    it = 0
    while out.shape[1] < T:                                                 # :1441
        blk = tape[it]
        L = out.shape[1]
        x = approx.generate(out.repeat(W, 1), g, blk[:g * W].view(g, W))    # :1531-1534
        q = approx.hist[:, L - 1:, :]                                       # :1542
        _ = target.generate(x, 1, blk[g * W:g * W + W].view(1, W))          # :1558 (W samples discarded)
        p = target.hist
        for w in range(W):                                                  # :1600-1609
            for i in range(g):
                j = int(x[w, L + i])
                qv = float(q[w, i, j])
                r = float((p[w, L + i - 1, j] / q[w, i, j])) if qv != 0 else 0.0
                acc_rate.append(0.0 if qv == 0 else min(r, 1.0))
        n_rand = 0
        is_all_accept = False
        max_n, max_l, choice = L - 1, 0, 0
        for w in range(W):                                                  # :1616-1640
            cur_n, cur_l, cur_all = L - 1, 0, True
            for i in range(g):
                r = blk[g * W + W + n_rand]
                n_rand += 1
                j = int(x[w, L + i])
                thr = torch.min(torch.tensor([1.0]), p[w, L + i - 1, j] / q[w, i, j])
                if bool(r < thr):
                    cur_l += 1
                    cur_n += 1
                else:
                    cur_all = False
                    break
            if cur_l > max_l:
                max_n, max_l, choice = cur_n, cur_l, w
                if cur_all:
                    is_all_accept = True
                    break
        acc_len.append(max_l)
        choices.append(choice)
        n = max_n
        out = x[choice:choice + 1, :n + 1]                                  # :1644
        approx.rollback(n + 1, choice)                                      # :1646
        u_final = float(blk[multi_block(g, W) - 1])
        if is_all_accept:
            tok = ref_ops.icdf_sample(p[choice, -1], u_final)               # :1649
            target.rollback(n + 2, choice)
        else:
            new_p = ref_ops.max_fn(p[choice:choice + 1, n, :] - q[choice:choice + 1, max_l, :])   # :1657
            try:
                tok = ref_ops.icdf_sample(new_p[0], u_final)
            except Exception:                                               # :1661-1663 (empty residual)
                tok = ref_ops.icdf_sample(p[choice, n], u_final)
            target.rollback(n + 1, choice)
        out = torch.cat([out, torch.tensor([[tok]], dtype=out.dtype)], dim=1)   # :1677
        it += 1
        if eos_token_id is not None:                                        # :1681-1689
            mask = out == eos_token_id
            if int(mask.sum()) > ori_eos:
                keep = torch.cumsum(mask.float(), dim=1) < ori_eos + 1
                end = int(keep.sum())
                if end < keep.shape[1]:
                    keep[:, end] = True
                out = out[keep][None, :]
                break
    details = dict(acc_len=acc_len, acc_rate=float(np.mean(acc_rate)) if acc_rate else 0.0, choices=choices,
                   target_call_times=it, approx_call_times=it)
    return out, details
