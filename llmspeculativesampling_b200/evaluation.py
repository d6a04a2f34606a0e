"""Engine side of the reference's evaluation driver (/root/reference/evaluation.py): the request loop
with its statistics (evaluation.py:515-583), the per-output target log-probability score `get_score` (evaluation.py:109-132)
and the autoregressive baseline loop that the reference times first (evaluation.py:421-440) — on the batched B200 engine.

Datasets, tokenizers, ROUGE and the GPU power monitor are outside the hot path and not part of this image;
the loop takes token-id prompts.  What the reference does one request at a time runs here `batch` requests at a time
(ragged prompts, one CUDA graph per iteration); per-request uniform tapes keep every request's tokens independent of the
batch it shared (sharding.py), so the statistics are those of the reference's loop over the same requests.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Union

import numpy as np
import torch


@torch.no_grad()
def get_score(output: Union[torch.Tensor, Sequence[torch.Tensor]], target_model, input_len: Union[int, Sequence[int]]):
    """evaluation.py:109-132 (decoder-only branch): mean over the generated positions of log p_target(token | prefix).

    `output` (1, n) with an int `input_len` returns a 0-d tensor exactly as the reference does; a list of (1, n_b) / (n_b,)
    outputs with a list of input lengths is scored in ONE target forward (right-padded: a causal model's positions do not
    see the padding behind them) and returns a list of 0-d tensors."""
    if getattr(target_model.config, "is_encoder_decoder", False):
        raise NotImplementedError("encoder-decoder scoring is outside the hot path (SURVEY.md §2)")
    single = isinstance(output, torch.Tensor)
    outs = [output.reshape(-1)] if single else [o.reshape(-1) for o in output]
    lens_in = [int(input_len)] if single else [int(x) for x in input_len]
    if outs[0].device.type != "cuda":
        raise RuntimeError("get_score needs CUDA tensors: there is no CPU path")
    B, n_max = len(outs), max(int(o.numel()) for o in outs)
    ids = torch.zeros(B, n_max, dtype=torch.int64, device=outs[0].device)
    mask = torch.zeros(B, n_max, dtype=torch.int64, device=outs[0].device)
    for b, o in enumerate(outs):
        ids[b, :o.numel()] = o
        mask[b, :o.numel()] = 1
    logits = target_model(ids, attention_mask=mask).logits[:, :-1, :]
    logp = torch.nn.functional.log_softmax(logits.float(), dim=-1)
    tok_lp = torch.gather(logp, dim=-1, index=ids[:, 1:, None])[:, :, 0]                 # (B, n_max - 1)
    scores = []
    for b, o in enumerate(outs):
        scores.append(tok_lp[b, lens_in[b] - 1:int(o.numel()) - 1].mean())              # evaluation.py:122
    return scores[0] if single else scores


@dataclass
class EvalStats:
    """The sums evaluation.py:515-542 keeps and the lines evaluation.py:567-583 prints from them."""
    name: str = "google speculative decoding (with KVCache)"
    total_time: int = 0                  # ns
    total_token: int = 0
    approx_time: int = 0
    target_time: int = 0
    other_time: int = 0
    total_acc_len: int = 0
    target_times: int = 0
    approx_times: int = 0
    target_model_time: int = 0
    target_pre_cache_time: int = 0
    target_post_prob_time: int = 0
    acc_rate: List[float] = field(default_factory=list)
    scores: List[float] = field(default_factory=list)
    requests: int = 0

    def add(self, elapsed_ns: int, new_tokens: int, details: Optional[dict], n_requests: int = 1) -> None:
        self.total_time += int(elapsed_ns)
        self.total_token += int(new_tokens)
        self.requests += n_requests
        if details is None:
            return
        acc = details["acc_len"]
        flat = [a for row in acc for a in row] if acc and isinstance(acc[0], (list, tuple)) else list(acc)
        self.total_acc_len += int(np.sum(flat)) if flat else 0
        # one engine iteration = one target call per request of the batch, as in the reference's one-request loop
        calls = len(flat)
        self.target_times += calls
        self.approx_times += calls
        self.acc_rate.append(float(details["acc_rate"]))
        for k in ("approx_time", "target_time", "other_time", "target_model_time", "target_pre_cache_time", "target_post_prob_time"):
            setattr(self, k, getattr(self, k) + int(details.get(k, 0)))

    def summary(self) -> dict:
        tok = max(self.total_token, 1)
        return {
            "total_time_s": self.total_time / 1e9, "total_tokens": self.total_token, "s_per_token": self.total_time / 1e9 / tok,
            "tokens_per_s": self.total_token / max(self.total_time / 1e9, 1e-12),
            "approx_time_s": self.approx_time / 1e9, "target_time_s": self.target_time / 1e9, "other_time_s": self.other_time / 1e9,
            "average_accepted_len": self.total_acc_len / max(self.target_times, 1), "target_call_times": self.target_times,
            "acc_rate": float(np.mean(self.acc_rate)) if self.acc_rate else 0.0, "approx_call_times": self.approx_times,
            "prob_score": float(np.mean(self.scores)) if self.scores else float("nan"),
            "target_model_time_s": self.target_model_time / 1e9, "pre_cache_time_s": self.target_pre_cache_time / 1e9,
            "post_prob_time_s": self.target_post_prob_time / 1e9, "requests": self.requests,
        }

    def lines(self) -> List[str]:
        """The report lines of evaluation.py:567-583 (power and ROUGE are not computed here)."""
        s = self.summary()
        return [
            f"\n {self.name} total time {s['total_time_s']} s, total tokens {s['total_tokens']}, average time {s['s_per_token']} s/token",
            f"approx time {s['approx_time_s']}, target time {s['target_time_s']}, other time {s['other_time_s']}",
            f"average accepted len {s['average_accepted_len']}, target call times {s['target_call_times']}, acc rate {s['acc_rate']}, approx call times {s['approx_call_times']}",
            f"prob score = {s['prob_score']}",
            f"target_model_time: {s['target_model_time_s']}, pre cache time: {s['pre_cache_time_s']}, post prob time: {s['post_prob_time_s']}",
        ]


def _batches(n: int, batch: int):
    for i in range(0, n, batch):
        yield list(range(i, min(i + batch, n)))


@torch.no_grad()
def evaluate_speculative(ds: Sequence[torch.Tensor], small_model, large_model, num_tokens: int, eos_token_id=None,
                         pad_token_id=None, gamma: int = 4, temperature: float = 1, top_k: int = 0, top_p: float = 0,
                         random_seed: Optional[int] = None, batch: int = 32, score: bool = True,
                         max_seconds: float = float("inf"), sampler=None):
    """evaluation.py:505-566: every prompt of `ds` ((1, n) or (n,) int64 CUDA tensors) through speculative_sampling with
    details, `batch` requests at a time; returns (outputs in dataset order, EvalStats).  `sampler` = another drop-in with
    the same signature (e.g. speculative_sampling_v2 wrapped by the caller)."""
    from .sampling import speculative_sampling
    fn = sampler or speculative_sampling
    stats = EvalStats()
    outputs: List[torch.Tensor] = []
    for ids in _batches(len(ds), batch):
        prompts = [ds[i].reshape(-1) for i in ids]
        t = time.perf_counter_ns()
        outs, details = fn(prompts, small_model, large_model, eos_token_id, pad_token_id, num_tokens, gamma, temperature,
                           top_k, top_p, False, random_seed, True, request_ids=ids)
        if isinstance(outs, torch.Tensor):
            outs = [outs]
        torch.cuda.synchronize()
        elapsed = time.perf_counter_ns() - t
        new = sum(int(o.numel()) - int(p.numel()) for o, p in zip(outs, prompts))
        stats.add(elapsed, new, details, len(ids))
        if score:
            stats.scores += [float(s) for s in get_score(outs, large_model, [int(p.numel()) for p in prompts])]
        outputs += list(outs)
        if stats.total_time / 1e9 > max_seconds:                              # evaluation.py:558-561
            break
    return outputs, stats


@torch.no_grad()
def evaluate_autoregressive(ds: Sequence[torch.Tensor], large_model, num_tokens: int, eos_token_id=None, pad_token_id=None,
                            temperature: float = 1, top_k: int = 0, top_p: float = 0, random_seed: Optional[int] = None,
                            batch: int = 32, score: bool = True):
    """evaluation.py:421-440: the large model alone (the speed-up denominator of the reference's report)."""
    from .sampling import autoregressive_sampling
    stats = EvalStats(name="large model autoregressive sampling")
    outputs: List[torch.Tensor] = []
    for ids in _batches(len(ds), batch):
        for i in ids:                                                         # (the drop-in takes one request per call)
            x = ds[i].reshape(1, -1)
            un = None
            if random_seed is not None:                                       # per-request tape: tokens independent of the order
                un = torch.rand(num_tokens, 1, generator=torch.Generator().manual_seed(int(random_seed) * 1000003 + i))
            t = time.perf_counter_ns()
            out = autoregressive_sampling(x, large_model, num_tokens, eos_token_id, temperature, top_k, top_p, pad_token_id,
                                          uniforms=un)
            torch.cuda.synchronize()
            stats.add(time.perf_counter_ns() - t, int(out.numel()) - int(x.numel()), None)
            outputs.append(out)
        if score:
            chunk = outputs[-len(ids):]
            stats.scores += [float(s) for s in get_score(chunk, large_model, [int(ds[i].numel()) for i in ids])]
    return outputs, stats
