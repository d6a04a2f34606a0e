"""Batched, ragged speculative-decoding engine on static KV caches (B200 host side).

What the reference does per request in Python is restructured here for B
requests in flight on one GPU, with NO host synchronisation inside an iteration:

  reference (batch 1)                                    here (batch B, device-resident state)
  ---------------------------------------------------    -------------------------------------------------
  KVCacheModel._past_key_values: tuple cache grown by    StaticKVCache: (B, H, S, D) buffers per layer,
    torch.cat, cropped by slicing                          rows written at per-request offsets by
    (sampling/kvcache_model.py:175,214,381-382)             sd_kv_append; rollback = smaller offset
  KVCacheModel._prob_history (1, seq, V) grown by cat    q_probs (B, gamma, V) / p_probs (B, gamma+1, V):
    (kvcache_model.py:246)                                  only the rows verify can still read
  generate(): gamma x (forward, norm_logits, sample,     gamma x (sd_build_step, HF forward, sd_norm_sample)
    torch.cat)  (kvcache_model.py:279-293)                  — fixed shapes, captured in ONE CUDA graph
  accept loop with ~7 .item() per token, max_fn,         sd_norm_probs on the gamma+1 target rows, then
    sample, rollback, cat                                   sd_verify (accept/ballot/residual/sample/append/
    (sampling/speculative_sampling.py:1966-2027)            length update) in the same graph

The draft model is always fed two tokens at its first step of an iteration (positions L-2, L-1):
after a fully accepted iteration the reference's draft cache misses exactly those two
(kvcache_model.py:206 feeds `input_ids[:, cached_len:]`), otherwise re-writing position L-2 is
idempotent.  That keeps every step's shape static, which is what makes the graph possible.
"""
from __future__ import annotations

import warnings
from typing import List, Optional, Sequence

import torch

from . import ops
from .uniform_tape import block as tape_block


class StaticKVCache:
    """Duck-typed Hugging Face cache: `update()` appends at `write_pos[b]` and returns the full buffers."""

    is_compileable = False

    def __init__(self, num_layers: int, batch: int, num_kv_heads: int, max_len: int, head_dim: int,
                 dtype: torch.dtype, device):
        self.k = [torch.zeros(batch, num_kv_heads, max_len, head_dim, dtype=dtype, device=device) for _ in range(num_layers)]
        self.v = [torch.zeros_like(t) for t in self.k]
        self.write_pos = torch.zeros(batch, dtype=torch.int32, device=device)
        self.dtype = dtype
        self.max_len = max_len
        self._seen = 0

    def update(self, key_states, value_states, layer_idx: int, cache_kwargs=None):
        ops.kv_append(key_states, value_states, self.k[layer_idx], self.v[layer_idx], self.write_pos)
        return self.k[layer_idx], self.v[layer_idx]

    def peek(self, layer_idx: int, positions: torch.Tensor) -> torch.Tensor:
        b = torch.arange(positions.shape[0], device=positions.device)
        return self.k[layer_idx][b, 0, positions, 0]

    # the HF modelling code only calls these when position_ids / a 4-D mask are NOT supplied
    def get_seq_length(self, layer_idx: int = 0) -> int:
        return self._seen

    def get_max_cache_shape(self, layer_idx: int = 0) -> int:
        return self.max_len

    def __len__(self):
        return len(self.k)


def _model_geometry(model):
    cfg = model.config
    layers = cfg.num_hidden_layers
    heads = cfg.num_attention_heads
    kv_heads = getattr(cfg, "num_key_value_heads", None) or heads
    head_dim = getattr(cfg, "head_dim", None) or cfg.hidden_size // heads
    return layers, kv_heads, head_dim, cfg.vocab_size


class ModelStepper:
    """One model + its static cache + the fixed-shape step buffers."""

    def __init__(self, model, batch: int, max_len: int, device, cache_dtype: Optional[torch.dtype] = None):
        self.model = model
        self.B, self.S = batch, max_len
        layers, kv_heads, head_dim, self.V = _model_geometry(model)
        if cache_dtype is None:
            cache_dtype = getattr(model, "kv_cache_dtype", None) or getattr(model, "dtype", torch.float32)
        self.cache = StaticKVCache(layers, batch, kv_heads, max_len, head_dim, cache_dtype, device)
        self.device = device
        self._bufs = {}
        impl = getattr(model.config, "_attn_implementation", "sdpa")
        self._additive_mask = impl == "eager"
        self._zero_len = torch.zeros(batch, dtype=torch.int32, device=device)

    def _step_bufs(self, q: int):
        if q not in self._bufs:
            ids = torch.zeros(self.B, q, dtype=torch.int64, device=self.device)
            pos = torch.zeros(self.B, q, dtype=torch.int64, device=self.device)
            mask = torch.zeros(self.B, 1, q, self.S, dtype=torch.uint8, device=self.device)
            self._bufs[q] = (ids, pos, mask)
        return self._bufs[q]

    def forward(self, tokens: torch.Tensor, seq_len: torch.Tensor, offset: int, q: int,
                prev_tok: Optional[torch.Tensor], last_only: bool = False) -> torch.Tensor:
        """Consume the q tokens at positions seq_len+offset .. of every request; returns logits (B, q, V)."""
        ids, pos, mask = self._step_bufs(q)
        ops.build_step(tokens, seq_len, offset, q, prev_tok, self.S, ids, pos, self.cache.write_pos, mask)
        attn = mask.view(torch.bool)
        if self._additive_mask:
            attn = torch.zeros(mask.shape, dtype=self.cache.dtype, device=self.device).masked_fill_(~attn, float("-inf"))
        kw = {"logits_to_keep": 1} if last_only else {}
        out = self.model(input_ids=ids, position_ids=pos, attention_mask=attn, past_key_values=self.cache,
                         use_cache=True, **kw)
        return out.logits

    def prefill(self, tokens: torch.Tensor, n_positions: int) -> None:
        """Fill the cache for positions 0 .. n_positions-1 of every request (padded prompts)."""
        chunk = 256
        for s in range(0, n_positions, chunk):
            q = min(chunk, n_positions - s)
            self.forward(tokens, self._zero_len, s, q, None, last_only=True)


class SpecDecEngine:
    """B requests of speculative sampling (Leviathan et al.) with reference semantics per request."""

    def __init__(self, approx_model, target_model, batch: int, max_total_len: int, gamma: int = 4,
                 temperature: float = 1.0, top_k: int = 0, top_p: float = 0.0, device=None,
                 strict: bool = False, use_cuda_graph: bool = True, max_iterations: int = 0):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("SpecDecEngine needs a CUDA device: there is no CPU path")
        if not 1 <= gamma <= 32:
            raise ValueError("gamma must be in [1, 32]")
        self.B, self.gamma = batch, gamma
        self.T, self.top_k, self.top_p = float(temperature), int(top_k or 0), float(top_p or 0.0)
        self.strict = strict
        # a finished request is still stepped while others run: it can sit at limit + gamma and the target step then
        # appends at limit + 2 * gamma - 1 (sd_build_step also bounds-checks)
        self.S = (max_total_len + 2 * gamma + 2 + 63) // 64 * 64
        self.draft = ModelStepper(approx_model, batch, self.S, self.device)
        self.target = ModelStepper(target_model, batch, self.S, self.device)
        if self.draft.V != self.target.V:
            raise ValueError("draft and target vocabularies differ")
        V = self.V = self.target.V
        dev = self.device
        g = gamma
        self.tokens = torch.zeros(batch, self.S, dtype=torch.int64, device=dev)
        self.seq_len = torch.full((batch,), 2, dtype=torch.int32, device=dev)
        self.prompt_len = torch.full((batch,), 2, dtype=torch.int32, device=dev)
        self.limit = torch.full((batch,), 2, dtype=torch.int32, device=dev)
        self.active = torch.zeros(batch, dtype=torch.int32, device=dev)
        self.q_probs = torch.zeros(batch, g, V, dtype=torch.float32, device=dev)
        self.p_probs = torch.zeros(batch, g + 1, V, dtype=torch.float32, device=dev)
        self.draft_tok = torch.zeros(batch, g, dtype=torch.int64, device=dev)
        self.cur_tok = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.n_acc = torch.zeros(batch, dtype=torch.int32, device=dev)
        self.next_tok = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.ratios = torch.zeros(batch, g, dtype=torch.float32, device=dev)
        self.ties = torch.zeros(1, dtype=torch.int32, device=dev)
        self.u_rows = torch.zeros(batch, tape_block(g), dtype=torch.float32, device=dev)
        self.u_draft_t = torch.zeros(g, batch, dtype=torch.float32, device=dev)
        self.u_final = torch.zeros(batch, dtype=torch.float32, device=dev)
        self.max_iterations = max_iterations or max_total_len
        self.acc_hist = torch.full((self.max_iterations, batch), -1, dtype=torch.int32, device=dev)
        self.ratio_hist = torch.zeros(self.max_iterations, batch, g, dtype=torch.float32, device=dev)
        self.it_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.eos = torch.full((1,), -1, dtype=torch.int64, device=dev)
        self._cols = torch.arange(self.S, device=dev, dtype=torch.int32).unsqueeze(0)
        self.err = ops.ErrFlag(dev)
        # compact (index, prob) lists of the filtered rows: kernel 2 then verifies from a few hundred bytes per request
        self.use_compact = 0 < self.top_k <= 128
        self.q_cmp = ops.CompactRows(batch * g, dev) if self.use_compact else None
        self.p_cmp = ops.CompactRows(batch * (g + 1), dev) if self.use_compact else None
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self.graph_captured = False
        self.phase_ns = None                                       # filled by run(profile_every=N)

    # ------------------------------------------------------------------ state
    def load_prompts(self, prompts: Sequence[torch.Tensor], max_new_tokens, eos_token_id: Optional[int] = None) -> None:
        """prompts: B 1-D int64 tensors (ragged).  max_new_tokens: int or per-request list."""
        assert len(prompts) == self.B
        lens = [int(p.numel()) for p in prompts]
        if min(lens) < 2:
            raise ValueError("every prompt needs at least 2 tokens")
        if isinstance(max_new_tokens, int):
            max_new_tokens = [max_new_tokens] * self.B
        host = torch.zeros(self.B, self.S, dtype=torch.int64)
        for b, p in enumerate(prompts):
            if lens[b] + max_new_tokens[b] + self.gamma + 1 > self.S:
                raise ValueError("prompt + max_new_tokens exceeds the engine's max_total_len")
            host[b, :lens[b]] = p.reshape(-1).cpu()
        self.tokens.copy_(host.to(self.device, non_blocking=True))
        lt = torch.tensor(lens, dtype=torch.int32)
        self.prompt_len.copy_(lt)
        self.seq_len.copy_(lt)
        self.limit.copy_(lt + torch.tensor(max_new_tokens, dtype=torch.int32))
        self.active.fill_(1)
        self.eos.fill_(-1 if eos_token_id is None else int(eos_token_id))
        self.acc_hist.fill_(-1)
        self.it_dev.zero_()
        self.ties.zero_()
        n_pre = max(lens)
        self.draft.prefill(self.tokens, n_pre)
        self.target.prefill(self.tokens, n_pre)

    # ------------------------------------------------------------------ one iteration (graph body)
    def _iteration(self, marks: Optional[list] = None) -> None:
        """marks: a list that receives CUDA events at the phase boundaries (start, drafts done, target forward done,
        target rows normalised, verified + bookkeeping) — the eager, event-timed twin of the captured graph."""
        def mark():
            if marks is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append(ev)
        g, B, V = self.gamma, self.B, self.V
        mark()
        self.u_draft_t.copy_(self.u_rows[:, :g].t())
        self.u_final.copy_(self.u_rows[:, 2 * g + 1])
        for i in range(g):
            if i == 0:
                logits = self.draft.forward(self.tokens, self.seq_len, -2, 2, None)[:, 1]
            else:
                logits = self.draft.forward(self.tokens, self.seq_len, i - 1, 1, self.cur_tok)[:, 0]
            ops.norm_sample(logits, self.T, self.top_k, self.top_p, self.u_draft_t[i], probs_out=self.q_probs[:, i],
                            tok_out=self.cur_tok, err=self.err,
                            compact=self.q_cmp.view(i, g) if self.use_compact else None)
            self.draft_tok[:, i].copy_(self.cur_tok)
        mark()
        logits = self.target.forward(self.tokens, self.seq_len, -1, g + 1, self.cur_tok)
        mark()
        ops.norm_probs(logits.reshape(B * (g + 1), V), self.T, self.top_k, self.top_p,
                       out=self.p_probs.view(B * (g + 1), V), err=self.err,
                       compact=self.p_cmp.view() if self.use_compact else None)
        mark()
        ops.verify(self.p_probs, self.q_probs, self.draft_tok, self.u_rows[:, g + 1:2 * g + 1], self.u_final,
                   strict=self.strict, n_accepted=self.n_acc, next_tok=self.next_tok, ratios=self.ratios,
                   tie_count=self.ties, tokens=self.tokens, seq_len=self.seq_len, active=self.active, err=self.err,
                   p_compact=self.p_cmp.view() if self.use_compact else None, p_cmp_req_stride=g + 1,
                   q_compact=self.q_cmp.view() if self.use_compact else None, q_cmp_req_stride=g)
        # statistics + termination, all on the device
        it = self.it_dev
        self.acc_hist.index_copy_(0, it, torch.where(self.active > 0, self.n_acc, torch.full_like(self.n_acc, -1)).unsqueeze(0))
        self.ratio_hist.index_copy_(0, it, self.ratios.unsqueeze(0))
        it.add_(1)
        gen = (self._cols >= self.prompt_len.unsqueeze(1)) & (self._cols < self.seq_len.unsqueeze(1))
        hit_eos = ((self.tokens == self.eos) & gen).any(dim=1)
        self.active.copy_(((self.seq_len < self.limit) & ~hit_eos & (self.active > 0)).to(torch.int32))
        mark()

    def _capture(self) -> None:
        """Warm up on a side stream and capture one iteration into a CUDA graph."""
        saved = [t.clone() for t in (self.tokens, self.seq_len, self.active, self.it_dev, self.acc_hist)]
        try:
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                for _ in range(2):
                    self.it_dev.zero_()
                    self._iteration()
                    for t, s_ in zip((self.tokens, self.seq_len, self.active), saved[:3]):
                        t.copy_(s_)
            torch.cuda.current_stream(self.device).wait_stream(s)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            self.it_dev.zero_()
            with torch.cuda.graph(graph):
                self._iteration()
            self._graph = graph
            self.graph_captured = True
        except Exception as e:     # HF model not capturable: stay eager (still no host sync per token)
            warnings.warn(f"CUDA graph capture failed ({type(e).__name__}: {e}); running the iteration eagerly")
            self._graph = None
            torch.cuda.synchronize(self.device)
        finally:
            for t, s_ in zip((self.tokens, self.seq_len, self.active, self.it_dev, self.acc_hist), saved):
                t.copy_(s_)
            self.err.t.zero_()

    # ------------------------------------------------------------------ driver
    def run(self, tape_dev: torch.Tensor, check_every: int = 1, profile_every: int = 0) -> int:
        """tape_dev: (iterations, B, 2*gamma+2) uniforms on the device.  Returns iterations executed.
        profile_every = N > 0: every N-th iteration (the first included) runs as the eager twin of the graph with CUDA
        events at its phase boundaries; self.phase_ns then holds the phase times of the whole run, scaled from the timed
        iterations (keys as the reference's `details`, sampling/speculative_sampling.py:2062-2073)."""
        if self.use_cuda_graph and self._graph is None and not self.graph_captured:
            self._capture()
        it = 0
        n_max = min(tape_dev.shape[0], self.max_iterations)
        timed: List[list] = []
        can_time = profile_every > 0 and type(self)._iteration is SpecDecEngine._iteration
        while it < n_max:
            self.u_rows.copy_(tape_dev[it])
            if can_time and it % profile_every == 0:
                marks: list = []
                self._iteration(marks)
                timed.append(marks)
            elif self._graph is not None:
                self._graph.replay()
            else:
                self._iteration()
            it += 1
            if it % check_every == 0 and int(self.active.sum().item()) == 0:
                break
        self.err.check()
        self.phase_ns = None
        if timed:
            torch.cuda.synchronize(self.device)
            tot = [0.0, 0.0, 0.0, 0.0]
            for m in timed:
                for j in range(4):
                    tot[j] += m[j].elapsed_time(m[j + 1]) * 1e6          # ms -> ns
            k = it / len(timed)
            self.phase_ns = {"approx_time": int(tot[0] * k), "target_model_time": int(tot[1] * k),
                             "target_post_prob_time": int(tot[2] * k), "target_time": int((tot[1] + tot[2]) * k),
                             "other_time": int(tot[3] * k), "target_pre_cache_time": 0, "timed_iterations": len(timed)}
        return it

    def results(self, eos_token_id: Optional[int] = None, device=None) -> List[torch.Tensor]:
        """Per-request token tensors (1, n) with the reference's EOS cut (speculative_sampling.py:2033-2041); on `device`
        if given (the reference returns its output on the prefix's device), else on the host."""
        toks = self.tokens.cpu()
        lens = self.seq_len.cpu().tolist()
        plen = self.prompt_len.cpu().tolist()
        outs = []
        for b in range(self.B):
            row = toks[b, :lens[b]]
            if eos_token_id is not None:
                new = (row[plen[b]:] == eos_token_id).nonzero()
                if new.numel() > 0:
                    row = row[:plen[b] + int(new[0]) + 1]
            outs.append(row.unsqueeze(0))
        if device is not None:
            outs = [o.to(device, non_blocking=True) for o in outs]
        return outs
