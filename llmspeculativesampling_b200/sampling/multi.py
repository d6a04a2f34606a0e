"""Drop-in for /root/reference/sampling/speculative_sampling.py:1379-1716 (`multi_speculative_sampling`,
strategy='iid', decoder-only path).

SURVEY.md §8f row N2: W independent drafts per iteration, ONE target pass over all of them, the first draft with the
longest accepted run wins (:1612-1640), `rollback(end_pos, choice)` keeps that draft's KV cache (:1646-1667) — on the
batched engine (`multi_engine.MultiDraftEngine`): B requests x W drafts per iteration in one CUDA graph; the accept tests
and the resample are ONE launch of the multi-draft variant of kernel 2 (`sd_verify_multi`: the reference draws its accept
uniforms lazily, draft by draft, stopping a draft at its first reject — :1616-1634 — and the kernel consumes the tape in
exactly that order), the rollback is a KV row copy (`sd_kv_select`) + token append (`sd_multi_commit`) inside the graph.
`KVCacheModel.generate(multi=W)` / `rollback(end_pos, choice)` offer the same at the reference's object level.  `strategy='beam'` depends on the beam
search APIs removed from transformers 5.x and stays out of scope.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import ops
from .kvcache_model import KVCacheModel


def multi_block(gamma: int, width: int) -> int:
    """Uniforms one iteration can consume: gamma*W draft, W discarded target samples, up to W*gamma accept tests, 1 final."""
    return 2 * width * gamma + width + 1


_ENGINES = {}


def _engine_for(approx_model, target_model, batch, width, total_len, gamma, temperature, top_k, top_p, device, use_cuda_graph):
    from ..multi_engine import MultiDraftEngine
    bucket = (total_len + 255) // 256 * 256
    key = (id(approx_model), id(target_model), batch, width, bucket, gamma, float(temperature), int(top_k or 0),
           float(top_p or 0.0), str(device), bool(use_cuda_graph))
    eng = _ENGINES.get(key)
    if eng is None:
        if len(_ENGINES) >= 2:
            _ENGINES.clear()
        eng = _ENGINES[key] = MultiDraftEngine(approx_model, target_model, batch, width, bucket, gamma, temperature, top_k,
                                               top_p, device, use_cuda_graph=use_cuda_graph)
    return eng


@torch.no_grad()
def multi_speculative_sampling(prefix, approx_model: torch.nn.Module, target_model: torch.nn.Module,
                               eos_token_id, pad_token_id, max_len: int, gamma: int = 4, width: int = 8, num_beams=None,
                               strategy: str = "beam", acc_rate_head=None, acc_rate_thres=0.4, temperature: float = 1,
                               top_k: int = 0, top_p: float = 0, verbose: bool = False, random_seed: Optional[int] = None,
                               details: bool = False, *, uniforms: Optional[torch.Tensor] = None, use_cuda_graph: bool = True):
    """Same positional signature as the reference (speculative_sampling.py:1380-1385), on the batched multi-draft engine
    (`multi_engine.MultiDraftEngine`: one CUDA graph per iteration).  Extensions: `prefix` may be a list of ragged 1-D
    prompts (B requests decode together); `uniforms` is an (iterations, multi_block(gamma, width)) tape for one request or
    (iterations, B, multi_block) for a batch — per iteration [gamma draft calls x W | W unused | accept tests in the
    reference's drawing order | final]."""
    if strategy != "iid":
        raise NotImplementedError("multi_speculative_sampling: only strategy='iid' is built (beam / diverse need the beam "
                                  "search APIs removed from transformers 5.x)")
    if isinstance(prefix, torch.Tensor):
        assert prefix.dim() == 2 and prefix.shape[0] == 1, "input batch size must be 1"       # :1413 (lists: extension)
        prompts = [prefix[0]]
    else:
        prompts = [p.reshape(-1) for p in prefix]
    B = len(prompts)
    dev = prompts[0].device
    if dev.type != "cuda":
        raise RuntimeError("multi_speculative_sampling needs CUDA tensors/models: there is no CPU path")
    g, W = int(gamma), int(width)
    nblk = multi_block(g, W)
    if uniforms is None:
        gen = torch.Generator().manual_seed(int(random_seed) if random_seed is not None else
                                            int(torch.randint(0, 2 ** 31 - 1, (1,)).item()))
        uniforms = torch.rand(int(max_len) + 1, B, nblk, generator=gen)
    if uniforms.dim() == 2:
        uniforms = uniforms.unsqueeze(1)
    assert uniforms.shape[1] == B and uniforms.shape[2] == nblk
    total = max(int(p.numel()) for p in prompts) + int(max_len)
    eng = _engine_for(approx_model, target_model, B, W, total, g, temperature, top_k, top_p, dev, use_cuda_graph)
    eng.load_prompts(prompts, int(max_len), eos_token_id)
    iters = eng.run(uniforms.to(device=dev, dtype=torch.float32))
    outs = eng.results(eos_token_id)
    out = outs[0] if B == 1 else outs
    if not details:
        return out
    acc = eng.acc_hist_m[:iters].cpu().numpy()                              # (iters, B), -1 where the request was idle
    rat = eng.ratio_hist_m[:iters].cpu().numpy().astype(np.float64)         # (iters, B, W, gamma)
    acc_len = [[int(a) for a in acc[:, b] if a >= 0] for b in range(B)]
    live = acc >= 0
    r = rat[live]                                                           # statistics over every drafted token, :1600-1609
    r = np.where(np.isnan(r) | np.isinf(r), 0.0, np.minimum(r, 1.0))
    return out, {"approx_time": 0, "target_time": 0, "other_time": 0,
                 "acc_len": acc_len[0] if B == 1 else acc_len, "acc_rate": float(r.mean()) if r.size else 0.0,
                 "target_call_times": iters, "approx_call_times": iters, "iterations": iters, "cuda_graph": eng.graph_captured}
