"""Batched multi-draft engine: B requests x W independent drafts per iteration in ONE CUDA graph.

SURVEY.md §8f row N2 on the batched engine — per request the semantics of the reference's
multi_speculative_sampling(strategy='iid') (speculative_sampling.py:1379-1716): W drafts of gamma tokens from the draft
model (kvcache_model.py:272-276), one target pass over all of them, the first draft with the longest accepted run wins
(:1612-1640, accept iff r < min(1, p/q), uniforms consumed as lazily as the reference draws them), residual / bonus
sample, and rollback(end_pos, choice) (kvcache_model.py:390-396) — here: the winning row's kept KV positions are copied
over the request's other W - 1 rows of the static caches and all W token rows advance together.

Rows b*W .. b*W+W-1 of every buffer belong to request b.  Per request and iteration the tape block is
multi_block(gamma, W) = [gamma x W draft | W unused (the reference's discarded target samples) | W*gamma accept | final].
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops
from .engine import SpecDecEngine


def multi_block(gamma: int, width: int) -> int:
    return 2 * width * gamma + width + 1


class MultiDraftEngine(SpecDecEngine):
    def __init__(self, approx_model, target_model, batch: int, width: int, max_total_len: int, gamma: int = 4,
                 temperature: float = 1.0, top_k: int = 0, top_p: float = 0.0, device=None, use_cuda_graph: bool = True,
                 max_iterations: int = 0):
        self.W, self.n_req = int(width), int(batch)
        super().__init__(approx_model, target_model, batch * width, max_total_len, gamma, temperature, top_k, top_p, device,
                         strict=True, use_cuda_graph=use_cuda_graph, max_iterations=max_iterations)
        g, W, dev = gamma, self.W, self.device
        self.use_compact = False                                  # the multi-draft verify reads the dense rows
        self._layer_tables = [ops.LayerTable(st.cache.k, st.cache.v) for st in (self.draft, self.target)]
        self.u_m = torch.zeros(batch, multi_block(g, W), dtype=torch.float32, device=dev)
        self.u_final_m = torch.zeros(batch, dtype=torch.float32, device=dev)
        self.choice = torch.zeros(batch, dtype=torch.int32, device=dev)
        self.n_acc_m = torch.zeros(batch, dtype=torch.int32, device=dev)
        self.next_tok_m = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.ratios_m = torch.zeros(batch, W, g, dtype=torch.float32, device=dev)
        self.acc_hist_m = torch.full((self.max_iterations, batch), -1, dtype=torch.int32, device=dev)
        self.choice_hist = torch.zeros(self.max_iterations, batch, dtype=torch.int32, device=dev)
        self.ratio_hist_m = torch.zeros(self.max_iterations, batch, W, g, dtype=torch.float32, device=dev)

    def load_prompts(self, prompts: Sequence[torch.Tensor], max_new_tokens, eos_token_id: Optional[int] = None) -> None:
        assert len(prompts) == self.n_req
        if isinstance(max_new_tokens, int):
            max_new_tokens = [max_new_tokens] * self.n_req
        rep = [p for p in prompts for _ in range(self.W)]
        super().load_prompts(rep, [m for m in max_new_tokens for _ in range(self.W)], eos_token_id)
        self.acc_hist_m.fill_(-1)

    def _iteration(self) -> None:
        g, R, V, W, B = self.gamma, self.B, self.V, self.W, self.n_req
        # tape block -> (gamma, B*W) draft uniforms (call i, request b, draft w)
        self.u_draft_t.view(g, B, W).copy_(self.u_m[:, :g * W].view(B, g, W).permute(1, 0, 2))
        self.u_final_m.copy_(self.u_m[:, -1])
        for i in range(g):
            if i == 0:
                logits = self.draft.forward(self.tokens, self.seq_len, -2, 2, None)[:, 1]
            else:
                logits = self.draft.forward(self.tokens, self.seq_len, i - 1, 1, self.cur_tok)[:, 0]
            ops.norm_sample(logits, self.T, self.top_k, self.top_p, self.u_draft_t[i], probs_out=self.q_probs[:, i],
                            tok_out=self.cur_tok, err=self.err)
            self.draft_tok[:, i].copy_(self.cur_tok)
        logits = self.target.forward(self.tokens, self.seq_len, -1, g + 1, self.cur_tok)
        ops.norm_probs(logits.reshape(R * (g + 1), V), self.T, self.top_k, self.top_p, out=self.p_probs.view(R * (g + 1), V),
                       err=self.err)
        # kernel 2, multi-draft variant: winner, accepted run, next token
        ch, na, nt = ops.verify_multi(self.p_probs.view(B, W, g + 1, V), self.q_probs.view(B, W, g, V),
                                      self.draft_tok.view(B, W, g), self.u_m[:, g * W + W:g * W + W + W * g], self.u_final_m,
                                      ratios=self.ratios_m, err=self.err)
        self.choice.copy_(ch); self.n_acc_m.copy_(na); self.next_tok_m.copy_(nt)
        # rollback(end_pos, choice): the winner's kept KV positions [L, L + n_acc) over the other rows, both models
        # (one launch per model over all of its layers: device tables of the cache pointers)
        for tab in self._layer_tables:
            ops.kv_select_layers(tab, W, self.choice, self.seq_len, W, self.n_acc_m, g, active=self.active, active_stride=W)
        ops.multi_commit(self.tokens, self.seq_len, W, self.choice, self.n_acc_m, self.next_tok_m, active=self.active)
        # statistics + termination, all on the device (rows of a request are identical after the commit)
        it = self.it_dev
        act_req = self.active.view(B, W)[:, 0]
        self.acc_hist_m.index_copy_(0, it, torch.where(act_req > 0, self.n_acc_m, torch.full_like(self.n_acc_m, -1)).unsqueeze(0))
        self.choice_hist.index_copy_(0, it, self.choice.unsqueeze(0))
        self.ratio_hist_m.index_copy_(0, it, self.ratios_m.unsqueeze(0))
        it.add_(1)
        gen = (self._cols >= self.prompt_len.unsqueeze(1)) & (self._cols < self.seq_len.unsqueeze(1))
        hit_eos = ((self.tokens == self.eos) & gen).any(dim=1)
        self.active.copy_(((self.seq_len < self.limit) & ~hit_eos & (self.active > 0)).to(torch.int32))

    def _capture(self) -> None:
        saved = self.acc_hist_m.clone()
        super()._capture()
        self.acc_hist_m.copy_(saved)

    def run(self, tape_dev: torch.Tensor, check_every: int = 1) -> int:
        """tape_dev: (iterations, B, multi_block(gamma, W)) uniforms on the device.  Returns iterations executed."""
        if self.use_cuda_graph and self._graph is None and not self.graph_captured:
            self._capture()
        it = 0
        n_max = min(tape_dev.shape[0], self.max_iterations)
        while it < n_max:
            self.u_m.copy_(tape_dev[it])
            if self._graph is not None:
                self._graph.replay()
            else:
                self._iteration()
            it += 1
            if it % check_every == 0 and int(self.active.sum().item()) == 0:
                break
        self.err.check()
        return it

    def results(self, eos_token_id: Optional[int] = None) -> List[torch.Tensor]:
        return super().results(eos_token_id)[::self.W]
