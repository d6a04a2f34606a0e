"""Batched BiLD engine: B requests of Big-Little decoding, one CUDA graph per check cycle.

SURVEY.md §8f row N3 on the batched engine — per request the semantics of the reference's BiLD_sampling
(speculative_sampling.py:1718-1873).  The reference drafts token by token and calls the target when the draft is unsure
(max q < fallback_thres, :1784) or gamma tokens are unchecked; a fixed-shape graph drafts gamma tokens up front instead
and the BiLD variant of kernel 2 (`sd_verify_bild`, engine mode) derives how many of them the reference would have
drafted, checks those (-log p[token] <= rollback_thres, :1800), samples the target's own token (:1812), appends and
advances the lengths — tokens drafted beyond the fallback point are simply overwritten by the next cycle.  The
reference tests its length limit before every draft token (:1764): a request with less room than drafted tokens ends
with those tokens unchecked, as the reference's loop does.

Tape block per request and cycle: the speculative block (2*gamma+2): u_draft[i] for the i-th draft token of the cycle,
[gamma] unused here (the reference's discarded target sample), [2*gamma+1] the target's token.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .engine import SpecDecEngine


class BiLDEngine(SpecDecEngine):
    def __init__(self, approx_model, target_model, batch: int, max_total_len: int, gamma: int, fallback_thres: float,
                 rollback_thres: float, temperature: float = 1.0, top_k: int = 0, top_p: float = 0.0, device=None,
                 use_cuda_graph: bool = True, max_iterations: int = 0):
        super().__init__(approx_model, target_model, batch, max_total_len, gamma, temperature, top_k, top_p, device,
                         strict=False, use_cuda_graph=use_cuda_graph, max_iterations=max_iterations)
        self.fallback_thres, self.rollback_thres = float(fallback_thres), float(rollback_thres)
        # compact lists of the q rows only (kernel 2's BiLD variant takes max q from them); the check itself reads dense p rows
        self.use_q_compact = 0 < self.top_k <= 128
        self.use_compact = False
        self.q_cmp = ops.CompactRows(batch * gamma, self.device) if self.use_q_compact else None
        self._eos_id = -1
        self.n_drafted = torch.zeros(batch, dtype=torch.int32, device=self.device)
        self.drafted_hist = torch.zeros(self.max_iterations, batch, dtype=torch.int32, device=self.device)

    def _iteration(self) -> None:
        g, B, V = self.gamma, self.B, self.V
        self.u_draft_t.copy_(self.u_rows[:, :g].t())
        self.u_final.copy_(self.u_rows[:, 2 * g + 1])
        for i in range(g):
            if i == 0:
                logits = self.draft.forward(self.tokens, self.seq_len, -2, 2, None)[:, 1]
            else:
                logits = self.draft.forward(self.tokens, self.seq_len, i - 1, 1, self.cur_tok)[:, 0]
            ops.norm_sample(logits, self.T, self.top_k, self.top_p, self.u_draft_t[i], probs_out=self.q_probs[:, i],
                            tok_out=self.cur_tok, err=self.err, compact=self.q_cmp.view(i, g) if self.use_q_compact else None)
            self.draft_tok[:, i].copy_(self.cur_tok)
        logits = self.target.forward(self.tokens, self.seq_len, -1, g + 1, self.cur_tok)
        ops.norm_probs(logits.reshape(B * (g + 1), V), self.T, self.top_k, self.top_p, out=self.p_probs.view(B * (g + 1), V),
                       err=self.err)
        # kernel 2, BiLD variant (engine mode): drafted length, check, the target's token, append, lengths
        ops.verify_bild(self.p_probs, self.draft_tok, self.rollback_thres, self.u_final, q_probs=self.q_probs,
                        fallback_thres=self.fallback_thres, n_drafted=self.n_drafted, tokens=self.tokens, seq_len=self.seq_len,
                        limit=self.limit, active=self.active, n_accepted=self.n_acc, next_tok=self.next_tok, err=self.err,
                        q_compact=self.q_cmp.view() if self.use_q_compact else None, q_cmp_req_stride=g, eos_token_id=self._eos_id)
        it = self.it_dev
        idle = torch.full_like(self.n_acc, -1000)
        self.acc_hist.index_copy_(0, it, torch.where(self.active > 0, self.n_acc, idle).unsqueeze(0))
        self.drafted_hist.index_copy_(0, it, torch.where(self.active > 0, self.n_drafted, torch.zeros_like(self.n_drafted)).unsqueeze(0))
        it.add_(1)
        gen = (self._cols >= self.prompt_len.unsqueeze(1)) & (self._cols < self.seq_len.unsqueeze(1))
        hit_eos = ((self.tokens == self.eos) & gen).any(dim=1)
        self.active.copy_(((self.seq_len < self.limit) & ~hit_eos & (self.active > 0)).to(torch.int32))

    def load_prompts(self, prompts, max_new_tokens, eos_token_id: Optional[int] = None) -> None:
        eos = -1 if eos_token_id is None else int(eos_token_id)
        if eos != self._eos_id:                                   # the EOS id is a launch argument: re-capture the graph
            self._eos_id = eos
            self._graph = None
            self.graph_captured = False
        super().load_prompts(prompts, max_new_tokens, eos_token_id)
        self.acc_hist.fill_(-1000)
