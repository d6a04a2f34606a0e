"""Drop-in for /root/reference/sampling/speculative_sampling.py:1718-1873 (`BiLD_sampling`, decoder-only path).

SURVEY.md §8f row N3, first version: the reference's own loop (one draft token at a time; the target looks only when the
draft is unsure or gamma tokens are unchecked) on the GPU building blocks of the hot path — `KVCacheModel` (static KV
cache, fused filter + softmax rows, inverse-CDF sampling from a uniform tape) — so tokens are reproducible and equal to
the reference's on the same uniforms.  The two policy quantities are read from the probability rows the kernels already
wrote: max q of the newest draft row (:1784), and the target's check (-log p[token] of the unchecked tokens, first failure,
the target's own token, :1797-1812) is one launch of the BiLD variant of kernel 2 (`sd_verify_bild`).
Batch 1 as in the reference (:1729); the batched / CUDA-graph engine serves `speculative_sampling` only.
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import ops, uniform_tape
from .kvcache_model import KVCacheModel


@torch.no_grad()
def BiLD_sampling(prefix: torch.Tensor, approx_model: torch.nn.Module, target_model: torch.nn.Module, gamma,
                  eos_token_id, pad_token_id, fallback_thres, rollback_thres, max_len: int, temperature: float = 1,
                  top_k: int = 0, top_p: float = 0, verbose: bool = False, random_seed: Optional[int] = None,
                  details: bool = False, *, uniforms: Optional[torch.Tensor] = None):
    """Same positional signature as the reference.  `uniforms`: optional (cycles, 2*gamma+2) tape — row c serves check
    cycle c: columns 0..gamma-1 the draft tokens of the cycle, column gamma the sample `target.generate(x, 1)` throws
    away (:1788), column 2*gamma+1 the target's own token (:1812); `random_seed` derives such a tape."""
    assert prefix.shape[0] == 1, "input batch size must be 1"               # :1729
    if prefix.device.type != "cuda":
        raise RuntimeError("BiLD_sampling needs CUDA tensors/models: there is no CPU path")
    dev = prefix.device
    gamma = int(gamma)
    seq_len = prefix.shape[1]
    T = seq_len + int(max_len)
    if uniforms is None:
        seed = int(random_seed) if random_seed is not None else int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
        uniforms = uniform_tape.make_tape(seed, int(max_len) + 1, gamma)
    tape = uniforms.to(device=dev, dtype=torch.float32)
    ori_eos_cnt = int((prefix == eos_token_id).sum()) if eos_token_id is not None else 0
    approx = KVCacheModel(approx_model, temperature, top_k, top_p, max_len=T + gamma + 2)
    target = KVCacheModel(target_model, temperature, top_k, top_p, max_len=T + gamma + 2)
    acc_len = []
    approx_call_times = target_call_times = 0
    last_check = seq_len - 1                                                # :1759
    cycle, n_draft = 0, 0
    out = prefix
    while prefix.shape[1] < T:                                              # :1764
        row = tape[min(cycle, tape.shape[0] - 1)]
        x = approx.generate(prefix, 1, uniforms=row[n_draft].view(1, 1))    # :1772
        n_draft += 1
        approx_call_times += 1
        q_max = float(approx._prob_history[0, x.shape[1] - 2].max())        # :1778-1784: the row the new token came from
        if q_max < fallback_thres or x.shape[1] - last_check - 1 >= gamma:
            _ = target.generate(x, 1, uniforms=row[gamma].view(1, 1))       # :1788 (sample discarded)
            target_call_times += 1
            # kernel 2, BiLD variant: -log p[token] of every unchecked token, first failure, the target's own token
            p = target._prob_history                                        # (1, len, V)
            c = x.shape[1] - 1 - last_check                                 # unchecked tokens (<= gamma)
            kept, t = ops.verify_bild(p[:, last_check:last_check + c + 1], x[:, last_check + 1:last_check + 1 + c],
                                      rollback_thres, row[2 * gamma + 1].view(1).contiguous())   # :1797-1812
            l = int(kept[0])
            n = last_check + l
            acc_len.append(l)
            prefix = x[:, :n + 1]                                           # :1806
            approx.rollback(n + 1)                                          # :1811
            target.rollback(n + 1)                                          # :1813
            last_check = n + 1
            prefix = torch.cat((prefix, t.view(1, 1)), dim=1)               # :1817
            cycle += 1
            n_draft = 0
        else:
            prefix = x                                                      # :1826
        out = prefix
        if eos_token_id is not None:                                        # :1833-1841
            mask = out == eos_token_id
            if int(mask.sum()) > ori_eos_cnt:
                keep = torch.cumsum(mask.float(), dim=1) < ori_eos_cnt + 1
                end = int(keep.sum())
                if end < keep.shape[1]:
                    keep[:, end] = True
                out = out[keep][None, :]
                break
    ops.default_flag(dev).check()
    if details:
        return out, {"approx_time": approx.forward_time_dict["_model_time"], "target_time": target.forward_time_dict["_model_time"],
                     "other_time": approx.forward_time_dict["norm_prob_time"] + target.forward_time_dict["norm_prob_time"],
                     "acc_len": acc_len, "acc_rate": float("nan"),           # the reference averages an empty list here (:1865)
                     "target_call_times": target_call_times, "approx_call_times": approx_call_times}
    return out
