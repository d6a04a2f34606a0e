"""Drop-in for /root/reference/sampling/kvcache_model.py:23-36,141-310,359-436 (decoder-only path).

Same constructor, `generate(input, gamma)`, `rollback(end_pos)`, and the attributes the algorithms
read (`_prob_history`, `_past_key_values`, `forward_time_dict`), but:
  * the KV cache is a static (B, H, S, D) buffer per layer written in place by sd_kv_append
    (no torch.cat growth, rollback = a smaller length),
  * `_prob_history` is a view into a pre-allocated (B, S, V) fp32 buffer that the fused kernel writes
    row-by-row in ONE launch per forward (the reference runs ~30 launches + 3 syncs per row),
  * sampling uses the inverse-CDF rule on a uniform per draw (pass `uniforms=` for reproducibility).
All rows of a batch must have the same length, as in the reference; ragged batches are served by
`llmspeculativesampling_b200.engine.SpecDecEngine`.
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import ops
from ..engine import ModelStepper


class _GpuTimes(dict):
    """forward_time_dict: nanoseconds measured with CUDA events, resolved lazily on read."""

    def __init__(self, keys):
        super().__init__({k: 0 for k in keys})
        self._pending = {k: [] for k in keys}

    def span(self, key):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self._pending[key].append((s, e))
        return s, e

    def __getitem__(self, key):
        pend = self._pending.get(key)
        if pend:
            torch.cuda.synchronize()
            total = sum(s.elapsed_time(e) for s, e in pend) * 1e6
            pend.clear()
            super().__setitem__(key, super().__getitem__(key) + int(total))
        return super().__getitem__(key)


class KVCacheModel:
    def __init__(self, model: torch.nn.Module, temperature: float = 1, top_k: int = 0, top_p: float = 0,
                 max_len: int = 1024) -> None:
        self._model = model
        self._temperature = temperature
        self._top_k = top_k
        self._top_p = top_p
        self._max_len = max_len
        self._stepper: Optional[ModelStepper] = None
        self._prob_buf: Optional[torch.Tensor] = None
        self._tokens: Optional[torch.Tensor] = None
        self._len_dev: Optional[torch.Tensor] = None
        self._n = 0                                       # cached positions == valid probability rows
        self.forward_time_dict = _GpuTimes(["_model_time", "norm_prob_time", "prepare_cache_time"])

    # ---- attributes the reference's algorithms read ------------------------------------------
    @property
    def _prob_history(self) -> Optional[torch.Tensor]:
        return None if self._prob_buf is None else self._prob_buf[:, :self._n]

    @property
    def _past_key_values(self):
        return None if self._stepper is None or self._n == 0 else self._stepper.cache

    def _ensure(self, input_ids: torch.Tensor) -> None:
        if self._stepper is not None:
            return
        if not input_ids.is_cuda:
            raise RuntimeError("KVCacheModel needs CUDA inputs: there is no CPU path")
        B = input_ids.shape[0]
        dev = input_ids.device
        self._stepper = ModelStepper(self._model, B, self._max_len, dev)
        self._prob_buf = torch.zeros(B, self._max_len, self._stepper.V, dtype=torch.float32, device=dev)
        self._tokens = torch.zeros(B, self._max_len, dtype=torch.int64, device=dev)
        self._len_dev = torch.zeros(B, dtype=torch.int32, device=dev)

    def _forward_with_kvcache(self, input_ids: torch.Tensor) -> torch.Tensor:
        """Feeds `input_ids[:, cached_len:]` (kvcache_model.py:206), normalises every new row
        (:166-168 / :235-236) and returns the last row's distribution (B, V)."""
        self._ensure(input_ids)
        cur = input_ids.shape[1]
        if cur > self._max_len:
            raise ValueError("sequence longer than KVCacheModel(max_len=...)")
        new = cur - self._n
        assert new >= 1, "nothing to feed (rollback first)"
        self._tokens[:, :cur].copy_(input_ids)
        self._len_dev.fill_(cur)
        B, V = input_ids.shape[0], self._stepper.V
        done = 0
        while done < new:
            q = min(256, new - done)
            s0, e0 = self.forward_time_dict.span("_model_time")
            s0.record()
            logits = self._stepper.forward(self._tokens, self._len_dev, -(new - done), q, None)
            e0.record()
            s1, e1 = self.forward_time_dict.span("norm_prob_time")
            s1.record()
            first = self._n + done
            if q == 1:                                            # a generation step: ONE launch for all B rows (strided rows)
                ops.norm_probs(logits[:, 0], self._temperature, self._top_k or 0, self._top_p or 0.0,
                               out=self._prob_buf[:, first])
            else:                                                 # prefill: the q rows of one request are contiguous in the buffer
                for b in range(B):
                    ops.norm_probs(logits[b], self._temperature, self._top_k or 0, self._top_p or 0.0,
                                   out=self._prob_buf[b, first:first + q])
            e1.record()
            done += q
        self._n = cur
        ops.default_flag(input_ids.device).check()                # 'norm logits error' once per forward
        return self._prob_buf[:, cur - 1]

    @torch.no_grad()
    def generate(self, input: torch.Tensor, gamma: int, uniforms: Optional[torch.Tensor] = None, multi: int = 1,
                 strategy: str = "beam") -> torch.Tensor:
        """gamma x (forward, sample, append) — kvcache_model.py:279-293.  uniforms: (gamma, B) or None.
        `multi=W, strategy='iid'` (kvcache_model.py:272-276): the single prefix is drafted W times independently."""
        x = input
        if multi > 1:
            if strategy != "iid":
                raise NotImplementedError("only strategy='iid' of the multi-draft path is built (beam needs HF 4.35 beam APIs)")
            if x.shape[0] == 1:
                x = x.repeat(multi, 1)
        for i in range(gamma):
            q = self._forward_with_kvcache(x)
            u = uniforms[i].reshape(-1) if uniforms is not None else torch.rand(x.shape[0], device=x.device)
            tok = ops.sample_rows(q, u.to(device=x.device, dtype=torch.float32).contiguous())
            x = torch.cat((x, tok.view(-1, 1)), dim=1)
        ops.default_flag(x.device).check()
        return x

    @torch.no_grad()
    def rollback(self, end_pos: int, choice: Optional[int] = None) -> None:
        """kvcache_model.py:360-436: crop cache and probability history to end_pos.  On static buffers this is a counter
        update; stale rows are overwritten by the next forward.  With `choice` (:390-396, :433-436: keep draft `choice`
        only — the next forward expands it to all rows again, :180-200) the kept positions of that row's KV cache are
        copied to every row of the static buffers."""
        assert self._stepper is not None and self._n > 0                   # reference: assert self._past_key_values
        self._n = min(self._n, int(end_pos))
        if choice is not None and self._stepper.B > 1:
            W, n, dev = self._stepper.B, self._n, self._tokens.device
            ch = torch.full((1,), int(choice), dtype=torch.int32, device=dev)
            start = torch.zeros(1, dtype=torch.int32, device=dev)
            count = torch.full((1,), n, dtype=torch.int32, device=dev)
            for k, v in zip(self._stepper.cache.k, self._stepper.cache.v):
                ops.kv_select(k, v, W, ch, start, 1, count, n)             # the W rows are one "request" of W drafts
            # the reference collapses _prob_history to the chosen row (:430-436); with static buffers every row gets it
            self._prob_buf[:, :n].copy_(self._prob_buf[int(choice):int(choice) + 1, :n].expand(W, -1, -1).clone())
