"""Drop-in for /root/reference/sampling/speculative_sampling.py:1718-1873 (`BiLD_sampling`, decoder-only path).

SURVEY.md §8f row N3.  Default: the batched BiLD engine (`bild_engine.BiLDEngine`: gamma tokens drafted up front in a
fixed-shape CUDA graph, `sd_verify_bild` derives how many of them the reference would have drafted, checks them, samples
the target's token and appends).  `use_engine=False`: the reference's own loop (one draft token at a time; the target
looks only when the draft is unsure or gamma tokens are unchecked) on `KVCacheModel` (static KV cache, fused filter +
softmax rows, inverse-CDF sampling from a uniform tape).  Either way tokens are reproducible and equal to the
reference's on the same uniforms.  The two policy quantities are read from the probability rows the kernels already
wrote: max q of the newest draft row (:1784), and the target's check (-log p[token] of the unchecked tokens, first failure,
the target's own token, :1797-1812) is one launch of the BiLD variant of kernel 2 (`sd_verify_bild`).
Batch 1 as in the reference (:1729); the batched / CUDA-graph engine serves `speculative_sampling` only.
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import ops, uniform_tape
from .kvcache_model import KVCacheModel


_ENGINES = {}


def _engine_for(approx_model, target_model, batch, total_len, gamma, fb, rb, temperature, top_k, top_p, device, use_cuda_graph):
    from ..bild_engine import BiLDEngine
    bucket = (total_len + 255) // 256 * 256
    key = (id(approx_model), id(target_model), batch, bucket, gamma, float(fb), float(rb), float(temperature), int(top_k or 0),
           float(top_p or 0.0), str(device), bool(use_cuda_graph))
    eng = _ENGINES.get(key)
    if eng is None:
        if len(_ENGINES) >= 2:
            _ENGINES.clear()
        eng = _ENGINES[key] = BiLDEngine(approx_model, target_model, batch, bucket, gamma, fb, rb, temperature, top_k, top_p,
                                         device, use_cuda_graph=use_cuda_graph)
    return eng


@torch.no_grad()
def BiLD_sampling(prefix, approx_model: torch.nn.Module, target_model: torch.nn.Module, gamma,
                  eos_token_id, pad_token_id, fallback_thres, rollback_thres, max_len: int, temperature: float = 1,
                  top_k: int = 0, top_p: float = 0, verbose: bool = False, random_seed: Optional[int] = None,
                  details: bool = False, *, uniforms: Optional[torch.Tensor] = None, use_engine: bool = True,
                  use_cuda_graph: bool = True):
    """Same positional signature as the reference (speculative_sampling.py:1719-1724).  Runs on the batched BiLD engine
    (`bild_engine.BiLDEngine`, one CUDA graph per check cycle); `prefix` may then also be a list of ragged 1-D prompts and
    `uniforms` (cycles, B, 2*gamma+2).  `use_engine=False` runs the reference's own token-granular loop on `KVCacheModel`
    (batch 1)."""
    if use_engine:
        return _bild_on_engine(prefix, approx_model, target_model, int(gamma), eos_token_id, fallback_thres, rollback_thres,
                               int(max_len), temperature, top_k, top_p, random_seed, details, uniforms, use_cuda_graph)
    return _bild_host_loop(prefix, approx_model, target_model, gamma, eos_token_id, pad_token_id, fallback_thres,
                           rollback_thres, max_len, temperature, top_k, top_p, verbose, random_seed, details, uniforms=uniforms)


def _bild_on_engine(prefix, approx_model, target_model, gamma, eos_token_id, fb, rb, max_len, temperature, top_k, top_p,
                    random_seed, details, uniforms, use_cuda_graph):
    if isinstance(prefix, torch.Tensor):
        assert prefix.dim() == 2 and prefix.shape[0] == 1, "input batch size must be 1"       # :1729 (lists: extension)
        prompts = [prefix[0]]
    else:
        prompts = [p.reshape(-1) for p in prefix]
    B = len(prompts)
    dev = prompts[0].device
    if dev.type != "cuda":
        raise RuntimeError("BiLD_sampling needs CUDA tensors/models: there is no CPU path")
    if uniforms is None:
        seed = int(random_seed) if random_seed is not None else int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
        uniforms = uniform_tape.batch_tape(seed, list(range(B)), max_len + 1, gamma)
    if uniforms.dim() == 2:
        uniforms = uniforms.unsqueeze(1)
    assert uniforms.shape[1] == B and uniforms.shape[2] == 2 * gamma + 2
    total = max(int(p.numel()) for p in prompts) + max_len
    eng = _engine_for(approx_model, target_model, B, total, gamma, fb, rb, temperature, top_k, top_p, dev, use_cuda_graph)
    eng.load_prompts(prompts, max_len, eos_token_id)
    iters = eng.run(uniforms.to(device=dev, dtype=torch.float32))
    outs = eng.results(eos_token_id)
    out = outs[0] if B == 1 else outs
    if not details:
        return out
    acc = eng.acc_hist[:iters].cpu().numpy()
    drafted = eng.drafted_hist[:iters].cpu().numpy()
    acc_len = [[int(a) for a in acc[:, b] if a >= 0] for b in range(B)]
    tcalls = [len(a) for a in acc_len]
    acalls = [int(drafted[:, b].sum()) for b in range(B)]
    one = B == 1
    return out, {"approx_time": 0, "target_time": 0, "other_time": 0, "acc_len": acc_len[0] if one else acc_len,
                 "acc_rate": float("nan"), "target_call_times": tcalls[0] if one else tcalls,
                 "approx_call_times": acalls[0] if one else acalls, "cycles": iters, "cuda_graph": eng.graph_captured}


@torch.no_grad()
def _bild_host_loop(prefix: torch.Tensor, approx_model: torch.nn.Module, target_model: torch.nn.Module, gamma,
                    eos_token_id, pad_token_id, fallback_thres, rollback_thres, max_len: int, temperature: float = 1,
                    top_k: int = 0, top_p: float = 0, verbose: bool = False, random_seed: Optional[int] = None,
                    details: bool = False, *, uniforms: Optional[torch.Tensor] = None):
    """The reference's own loop on `KVCacheModel`.  `uniforms`: optional (cycles, 2*gamma+2) tape — row c serves check
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
