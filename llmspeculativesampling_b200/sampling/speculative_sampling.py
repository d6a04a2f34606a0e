"""Drop-in for /root/reference/sampling/speculative_sampling.py:1876-2194
(`speculative_sampling`, `speculative_sampling_v2`) on the batched B200 engine.

Per request the semantics are the reference's (draft gamma -> one target pass -> accept scan ->
residual / bonus sample -> rollback -> append -> EOS cut, including the possible overshoot of
max_len by up to gamma tokens); extensions: B >= 1 prefixes per call and an explicit uniform tape.
"""
from __future__ import annotations

import time
from collections import OrderedDict
from typing import List, Optional, Sequence, Union

import numpy as np
import torch

from ..engine import SpecDecEngine
from .. import uniform_tape

_ENGINES: "OrderedDict[tuple, SpecDecEngine]" = OrderedDict()
_MAX_ENGINES = 4
_PROFILE_EVERY = 8          # details=True: every 8th iteration runs eagerly with CUDA events at the phase boundaries


def _engine_for(approx_model, target_model, batch, total_len, gamma, temperature, top_k, top_p, device, strict,
                use_cuda_graph) -> SpecDecEngine:
    bucket = (total_len + 255) // 256 * 256
    key = (id(approx_model), id(target_model), batch, bucket, gamma, float(temperature), int(top_k or 0),
           float(top_p or 0.0), str(device), bool(strict), bool(use_cuda_graph))
    eng = _ENGINES.get(key)
    if eng is None:
        eng = SpecDecEngine(approx_model, target_model, batch, bucket, gamma, temperature, top_k, top_p, device,
                            strict=strict, use_cuda_graph=use_cuda_graph)
        _ENGINES[key] = eng
        while len(_ENGINES) > _MAX_ENGINES:
            _ENGINES.popitem(last=False)
    else:
        _ENGINES.move_to_end(key)
    return eng


def clear_engine_cache() -> None:
    _ENGINES.clear()


def _as_prompts(prefix) -> List[torch.Tensor]:
    if isinstance(prefix, torch.Tensor):
        assert prefix.dim() == 2, "prefix must be (batch, prefix_seqlen)"
        return [prefix[b] for b in range(prefix.shape[0])]
    return [p.reshape(-1) for p in prefix]


def _run(prefix, approx_model, target_model, eos_token_id, max_len, gamma, temperature, top_k, top_p, random_seed,
         details, uniforms, strict, use_cuda_graph, request_ids, pad_batch=False):
    prompts = _as_prompts(prefix)
    B = len(prompts)
    device = prompts[0].device
    if device.type != "cuda":
        raise RuntimeError("speculative_sampling needs CUDA tensors/models: there is no CPU path")
    total = max(int(p.numel()) for p in prompts) + int(max_len)
    # pad_batch (serving): the engine (static KV caches, probability buffers, captured graph) is keyed on the batch size, so
    # arbitrary sizes from a request queue are rounded up to a power of two with dummy requests that may generate nothing
    B_eng = 1 << (B - 1).bit_length() if pad_batch else B
    eng = _engine_for(approx_model, target_model, B_eng, total, gamma, temperature, top_k, top_p, device, strict,
                      use_cuda_graph)
    if uniforms is None:
        if random_seed is not None:
            ids = list(request_ids) if request_ids is not None else list(range(B))
            uniforms = uniform_tape.batch_tape(random_seed, ids, int(max_len) + 1, gamma)
        else:
            uniforms = torch.rand(int(max_len) + 1, B, uniform_tape.block(gamma))
    assert uniforms.shape[1] == B and uniforms.shape[2] == uniform_tape.block(gamma)
    tape_dev = uniforms.to(device=device, dtype=torch.float32)
    new_tokens = [int(max_len)] * B
    if B_eng > B:
        tape_dev = torch.cat([tape_dev, tape_dev.new_zeros(tape_dev.shape[0], B_eng - B, tape_dev.shape[2])], dim=1)
        prompts = prompts + [prompts[0][:2]] * (B_eng - B)
        new_tokens = new_tokens + [0] * (B_eng - B)
    t0 = time.perf_counter_ns()
    try:
        eng.load_prompts(prompts, new_tokens, eos_token_id)
        iters = eng.run(tape_dev, profile_every=_PROFILE_EVERY if details else 0)
        outs = eng.results(eos_token_id, device)[:B]
    except RuntimeError as e:
        if str(e) in ("norm logits error", "prob error", "s"):
            print(e)
            raise RuntimeError("s")                                        # reference :2044-2046
        raise
    elapsed = time.perf_counter_ns() - t0
    out = outs[0] if B == 1 else outs
    if not details:
        return out
    acc = eng.acc_hist[:iters, :B].cpu().numpy()                            # (iters, B), -1 where the request was idle
    rat = eng.ratio_hist[:iters, :B].cpu().numpy()
    acc_len = [[int(a) for a in acc[:, b] if a >= 0] for b in range(B)]
    live = acc >= 0
    if strict:
        # v2 appends min(1, p/q) only for the tokens it tests — up to and including the first reject (:2155)
        tested = np.arange(rat.shape[2])[None, None, :] <= acc[:, :, None]
        rates = np.minimum(1.0, rat.astype(np.float64))[live[:, :, None] & tested]
    else:
        rates = np.minimum(1.0, rat.astype(np.float64))[live]              # every drafted token (:1966-1971)
    # phase times (ns, as the reference's process_time_ns sums, :2062-2073): CUDA-event times of the eager twin of the graph,
    # every _PROFILE_EVERY-th iteration, scaled to the run.  approx_time = the gamma draft steps (forward + kernel 1b),
    # target_time = target forward + kernel 1, other_time = kernel 2 + bookkeeping; the static caches need no per-call
    # cache preparation (target_pre_cache_time = 0).
    ph = eng.phase_ns or {}
    d = {
        "approx_time": ph.get("approx_time", 0), "target_time": ph.get("target_time", 0),
        "other_time": ph.get("other_time", 0),
        "acc_len": acc_len[0] if B == 1 else acc_len,
        "acc_rate": float(rates.mean()) if rates.size else 0.0,
        "target_call_times": iters, "approx_call_times": iters,
        "target_model_time": ph.get("target_model_time", 0), "target_pre_cache_time": 0,
        "target_post_prob_time": ph.get("target_post_prob_time", 0),
        "total_time_ns": elapsed, "iterations": iters, "cuda_graph": eng.graph_captured,
        "exact_ties": int(eng.ties.item()), "timed_iterations": ph.get("timed_iterations", 0),
    }
    return out, d


@torch.no_grad()
def speculative_sampling(prefix: Union[torch.Tensor, Sequence[torch.Tensor]], approx_model: torch.nn.Module,
                         target_model: torch.nn.Module, eos_token_id=None, pad_token_id=None, max_len: int = 128,
                         gamma: int = 4, temperature: float = 1, top_k: int = 0, top_p: float = 0,
                         verbose: bool = False, random_seed: int = None, details: bool = False, *,
                         uniforms: Optional[torch.Tensor] = None, use_cuda_graph: bool = True,
                         request_ids: Optional[Sequence[int]] = None, pad_batch: bool = False):
    """Google / Leviathan speculative sampling with KV caches (reference :1877-2076).

    Positional order is the reference's: (prefix, approx_model, target_model, eos_token_id,
    pad_token_id, max_len, gamma, temperature, top_k, top_p, verbose, random_seed, details).
    Accept rule: reject iff u > p/q (:1981).  `random_seed` seeds the per-request uniform tape
    (the reference re-seeds torch before every accept draw, :1976-1977, which makes all its accept
    uniforms equal — that quirk is not reproduced)."""
    return _run(prefix, approx_model, target_model, eos_token_id, max_len, gamma, temperature, top_k, top_p,
                random_seed, details, uniforms, False, use_cuda_graph, request_ids, pad_batch)


@torch.no_grad()
def speculative_sampling_v2(prefix: Union[torch.Tensor, Sequence[torch.Tensor]], approx_model: torch.nn.Module,
                            target_model: torch.nn.Module, max_len: int, gamma: int = 4, temperature: float = 1,
                            top_k: int = 0, top_p: float = 0, random_seed: int = None, details: bool = False, *,
                            uniforms: Optional[torch.Tensor] = None, use_cuda_graph: bool = True,
                            request_ids: Optional[Sequence[int]] = None):
    """DeepMind variant (reference :2080-2194): accept iff u < min(1, p/q) (:2156).  The reference
    re-runs full forwards; here the same KV-cached engine is used (identical maths)."""
    return _run(prefix, approx_model, target_model, None, max_len, gamma, temperature, top_k, top_p, random_seed,
                details, uniforms, True, use_cuda_graph, request_ids)
