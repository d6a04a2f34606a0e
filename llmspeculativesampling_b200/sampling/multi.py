"""Drop-in for /root/reference/sampling/speculative_sampling.py:1379-1716 (`multi_speculative_sampling`,
strategy='iid', decoder-only path).

SURVEY.md §8f row N2, first version: W independent drafts per iteration (`KVCacheModel.generate(multi=W)`), ONE target
pass over all of them, the first draft with the longest accepted run wins (:1612-1640), `rollback(end_pos, choice)`
keeps that draft's KV cache (:1646-1667).  Everything per row runs on the hot-path kernels (fused filter + softmax,
inverse-CDF sampling); the accept tests and the resample of an iteration are ONE launch of the multi-draft variant of
kernel 2 (`sd_verify_multi`: the reference draws its accept uniforms lazily, draft by draft, stopping a draft at its first
reject — :1616-1634 — and the kernel consumes the tape in exactly that order).  Batch 1 as in the reference (:1413).  `strategy='beam'` depends on the beam
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


@torch.no_grad()
def multi_speculative_sampling(prefix: torch.Tensor, approx_model: torch.nn.Module, target_model: torch.nn.Module,
                               eos_token_id, pad_token_id, max_len: int, gamma: int = 4, width: int = 8, num_beams=None,
                               strategy: str = "beam", acc_rate_head=None, acc_rate_thres=0.4, temperature: float = 1,
                               top_k: int = 0, top_p: float = 0, verbose: bool = False, random_seed: Optional[int] = None,
                               details: bool = False, *, uniforms: Optional[torch.Tensor] = None):
    """Same positional signature as the reference.  `uniforms`: optional (iterations, multi_block(gamma, width)) tape —
    per iteration [gamma draft calls x W | W discarded | accept tests in drawing order | final]."""
    if strategy != "iid":
        raise NotImplementedError("multi_speculative_sampling: only strategy='iid' is built (beam / diverse need the beam "
                                  "search APIs removed from transformers 5.x)")
    assert prefix.shape[0] == 1, "input batch size must be 1"               # :1413
    if prefix.device.type != "cuda":
        raise RuntimeError("multi_speculative_sampling needs CUDA tensors/models: there is no CPU path")
    dev = prefix.device
    g, W = int(gamma), int(width)
    T = prefix.shape[1] + int(max_len)
    nblk = multi_block(g, W)
    if uniforms is None:
        gen = torch.Generator().manual_seed(int(random_seed) if random_seed is not None else
                                            int(torch.randint(0, 2 ** 31 - 1, (1,)).item()))
        uniforms = torch.rand(int(max_len) + 1, nblk, generator=gen)
    tape_h = uniforms.to("cpu", torch.float32)
    tape_d = tape_h.to(dev)
    ori_eos_cnt = int((prefix == eos_token_id).sum()) if eos_token_id is not None else 0
    approx = KVCacheModel(approx_model, temperature, top_k, top_p, max_len=T + g + 2)
    target = KVCacheModel(target_model, temperature, top_k, top_p, max_len=T + g + 2)
    acc_len, acc_rate = [], []
    out = prefix
    it = 0
    while out.shape[1] < T:                                                 # :1441
        blk_h, blk_d = tape_h[min(it, tape_h.shape[0] - 1)], tape_d[min(it, tape_d.shape[0] - 1)]
        L = out.shape[1]
        x = approx.generate(out, g, uniforms=blk_d[:g * W].view(g, W), multi=W, strategy="iid")   # :1531 -> (W, L + g)
        _ = target.generate(x, 1, uniforms=blk_d[g * W:g * W + W].view(1, W))                     # :1558 (W samples discarded)
        q_hist, p_hist = approx._prob_history, target._prob_history
        # kernel 2, multi-draft variant: accept scan over the W drafts (uniforms consumed in the reference's drawing
        # order), first longest run wins, residual / bonus sample — one launch, one small read-back
        ratios = torch.empty(1, W, g, dtype=torch.float32, device=dev)
        ch, na, nt = ops.verify_multi(p_hist[:, L - 1:L + g].unsqueeze(0), q_hist[:, L - 1:L - 1 + g].unsqueeze(0),
                                      x[:, L:L + g].unsqueeze(0), blk_d[g * W + W:g * W + W + W * g].view(1, W * g),
                                      blk_d[nblk - 1].view(1).contiguous(), ratios=ratios)
        choice, max_l, t = int(ch[0]), int(na[0]), nt
        r_h = ratios[0].cpu()
        for w in range(W):                                                  # :1600-1609 (statistics over ALL drafted tokens)
            for i in range(g):
                rv = float(r_h[w, i])
                acc_rate.append(0.0 if (rv != rv or rv == float("inf")) else min(rv, 1.0))
        acc_len.append(max_l)
        n = L - 1 + max_l
        out = x[choice:choice + 1, :n + 1]                                  # :1644
        approx.rollback(n + 1, choice)                                      # :1646
        target.rollback(n + 2 if max_l == g else n + 1, choice)             # :1650 / :1667
        out = torch.cat((out, t.view(1, 1)), dim=1)                         # :1677
        it += 1
        if eos_token_id is not None:                                        # :1681-1689
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
                     "acc_len": acc_len, "acc_rate": float(np.mean(acc_rate)) if acc_rate else 0.0,
                     "target_call_times": it, "approx_call_times": it}
    return out
