"""The reference's op chain for one speculative iteration, restated for CUDA tensors (TEST / BASELINE INFRASTRUCTURE ONLY).

SURVEY.md §2.2 and BASELINE.md §3 call "the reference's ATen op chain run on the same B200" the bar the fused kernels
must beat.  /root/reference cannot travel to the GPU box, so this file restates, launch for launch and host sync for
host sync, what the reference does between the models' logits and the next step's input — one request at a time, one
row per Python iteration, with its own RNG (torch.multinomial / torch.rand on the device):

  norm_logits        /root/reference/sampling/utils.py:182-210 (+ top_k_top_p_filter :152-179), 3 host syncs per row
  sample             utils.py:213-233 (nonzero count sync, multinomial, zero-prob guard sync)
  max_fn             utils.py:236-245
  one iteration      sampling/kvcache_model.py:235-236, 280-283 (per-row norm_logits, sample per drafted token, the
                     discarded sample of target.generate(x, 1)) and sampling/speculative_sampling.py:1966-2027
                     (statistics loop, accept loop with .item(), residual / bonus sample)

bench.py times it as `gpu_aten_baseline`; nothing in the product imports it.
"""
from __future__ import annotations

import torch


def top_k_top_p_filter(logits: torch.Tensor, top_k: int = 0, top_p: float = 0.0) -> torch.Tensor:   # utils.py:152-179
    if top_k is not None and top_k > 0:
        filt = torch.topk(logits, min(top_k, logits.size(-1)))[0]
        logits[logits < filt[:, [-1]]] = float("-inf")
    if top_p is not None and top_p > 0.0:
        sorted_logits, sorted_indices = torch.sort(logits, descending=True)
        cumulative = torch.cumsum(torch.softmax(sorted_logits, dim=-1), dim=-1)
        remove = cumulative > top_p
        remove[..., 1:] = remove[..., :-1].clone()
        remove[..., 0] = 0
        logits[remove.scatter(1, sorted_indices, remove)] = float("-inf")
    return logits


def norm_logits(logits: torch.Tensor, temperature: float, top_k: int, top_p: float) -> torch.Tensor:  # utils.py:182-210
    assert logits.dim() == 2
    logits = logits / temperature
    logits = top_k_top_p_filter(logits, top_k=top_k, top_p=top_p)
    probs = torch.log_softmax(logits, dim=1).exp()
    if torch.any(torch.isnan(probs)) or torch.any(torch.isinf(probs)) or torch.any(probs < 0):       # 3 host syncs
        raise RuntimeError("norm logits error")
    return probs


def sample(probs: torch.Tensor, num_samples: int = 1) -> torch.Tensor:                                # utils.py:213-233
    nnz = probs.nonzero().shape[0]                                                                    # host sync
    idx_next = torch.multinomial(probs, num_samples=num_samples, replacement=(nnz < num_samples))
    if torch.any(torch.gather(probs, 1, idx_next) < 1e-9):                                            # host sync
        idx_next = torch.argmax(probs).reshape(1, 1)
    return idx_next


def max_fn(x: torch.Tensor) -> torch.Tensor:                                                          # utils.py:236-245
    x_max = torch.where(x > 0, x, torch.zeros_like(x))
    return x_max / (torch.sum(x_max, dim=1, keepdim=True) + 1e-6)


def iteration(draft_logits: torch.Tensor, target_logits: torch.Tensor, temperature: float, top_k: int, top_p: float):
    """One request: draft_logits (gamma, V), target_logits (gamma+1, V) on the device -> (n_accepted, next token)."""
    gamma = draft_logits.shape[0]
    q_rows, toks = [], []
    for i in range(gamma):                                                # kvcache_model.py:279-293
        q = norm_logits(draft_logits[i:i + 1].float(), temperature, top_k, top_p)
        q_rows.append(q)
        toks.append(sample(q))
    p_rows = [norm_logits(target_logits[i:i + 1].float(), temperature, top_k, top_p) for i in range(gamma + 1)]   # :235-236
    _ = sample(p_rows[gamma])                                             # speculative_sampling.py:1956 (discarded)
    acc_rate = []
    for i in range(gamma):                                                # :1966-1971 statistics loop (3 .item() per token)
        j = toks[i].item()
        acc_rate.append(min(1.0, p_rows[i][0, j].item() / q_rows[i][0, j].item()))
    n = gamma
    for i in range(gamma):                                                # :1975-1990 accept loop
        r = torch.rand(1, device=draft_logits.device)
        j = toks[i].item()
        if r > p_rows[i][0, j].item() / q_rows[i][0, j].item():
            n = i
            break
    if n < gamma:                                                         # :2005-2015
        try:
            t = sample(max_fn(p_rows[n] - q_rows[n]))
        except RuntimeError:
            t = sample(max_fn(p_rows[n]))
    else:
        t = sample(p_rows[gamma])                                         # :2016-2023
    return n, int(t.item())
