"""Fake causal LM for parity tests (TEST INFRASTRUCTURE ONLY).

The logits of position t are a pure, integer-indexed table look-up keyed by a rolling hash of the
token prefix 0..t, so every implementation — the reference (batch 1, legacy tuple KV cache,
/root/reference/sampling/kvcache_model.py:141-252), the oracle loop and the batched B200 engine
(static KV cache, ragged lengths) — sees BIT-IDENTICAL logits on CPU and on GPU, while a stale or
wrongly rolled-back cache entry still changes every later logit (the hash chain lives in the
"KV cache": one fp32-exact integer per position).

Two calling conventions:
* legacy:  ``model(input_ids, past_key_values=None | tuple-of-(k,v), use_cache=True)`` ->
           ``.logits`` (B, q, V), ``.past_key_values`` = ((k, v),) with k, v of shape (B, 1, seq, 1)
           — what the reference's KVCacheModel expects (it reads ``[0][0].shape[2]`` and slices
           ``[:, :, :end_pos, :]``).
* engine:  ``model(input_ids, position_ids=(B,q), past_key_values=<cache with .peek/.update>)``.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

_P = 8388593          # prime < 2**23 : hash states are exact in fp32
_A = 48271
_B = 69621
_TABLE = 1 << 20


class _Out(SimpleNamespace):
    pass


class ReplayLM(torch.nn.Module):
    def __init__(self, vocab_size: int, table_seed: int = 0, noise: float = 0.0, noise_seed: int = 1,
                 scale: float = 3.0, dtype: torch.dtype = torch.float32, device="cpu"):
        super().__init__()
        g = torch.Generator(device="cpu").manual_seed(1000 + table_seed)
        base = torch.randn(_TABLE, generator=g) * scale
        if noise > 0.0:
            g2 = torch.Generator(device="cpu").manual_seed(5000 + noise_seed)
            base = base + noise * torch.randn(_TABLE, generator=g2)
        # round once through the serving dtype so that bf16/fp16 variants are exact too
        self.register_buffer("table", base.to(dtype).to(device), persistent=False)
        self.vocab_size = vocab_size
        self.config = SimpleNamespace(is_encoder_decoder=False, vocab_size=vocab_size,
                                      num_hidden_layers=1, num_key_value_heads=1, head_dim=4,
                                      hidden_size=4, num_attention_heads=1)
        self.kv_cache_dtype = torch.float32            # hash states must stay exact in the engine's cache
        self._dummy = torch.nn.Parameter(torch.zeros(1), requires_grad=False)

    @property
    def device(self):
        return self.table.device

    @property
    def dtype(self):
        return self.table.dtype

    # -- hash chain -------------------------------------------------------------------------
    @staticmethod
    def hash_step(h_prev: torch.Tensor, tok: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
        return (h_prev * _A + (tok % _P) * _B + pos + 1) % _P

    def logits_of(self, h: torch.Tensor) -> torch.Tensor:
        v = torch.arange(self.vocab_size, device=h.device, dtype=torch.int64)
        idx = (h.unsqueeze(-1) * 2654435761 + v * 40503 + 17) % _TABLE
        return self.table[idx]

    def _chain(self, h_prev: torch.Tensor, input_ids: torch.Tensor, pos0: torch.Tensor) -> torch.Tensor:
        hs = []
        h = h_prev
        for j in range(input_ids.shape[1]):
            h = self.hash_step(h, input_ids[:, j].to(torch.int64), pos0 + j)
            hs.append(h)
        return torch.stack(hs, dim=1)                                   # (B, q) int64

    # -- forward ----------------------------------------------------------------------------
    def forward(self, input_ids, past_key_values=None, use_cache=True, position_ids=None,
                attention_mask=None, **kwargs):
        B, q = input_ids.shape
        dev = input_ids.device
        if past_key_values is not None and hasattr(past_key_values, "peek"):
            # engine convention: ragged batch, explicit positions, static cache object
            pos0 = position_ids[:, 0].to(torch.int64)
            prev = past_key_values.peek(0, (pos0 - 1).clamp_min(0))      # (B,) fp32 exact ints
            h_prev = torch.where(pos0 > 0, prev.to(torch.int64), torch.zeros_like(pos0))
            h = self._chain(h_prev, input_ids, pos0)
            kv = h.to(past_key_values.dtype).view(B, 1, q, 1).expand(B, 1, q, 4).contiguous()
            past_key_values.update(kv, kv, 0)
            return _Out(logits=self.logits_of(h), past_key_values=past_key_values)
        if past_key_values is None:
            h_prev = torch.zeros(B, dtype=torch.int64, device=dev)
            pos0 = torch.zeros(B, dtype=torch.int64, device=dev)
            old = None
        else:
            old = past_key_values[0][0]                                  # (B, 1, seq, 1)
            h_prev = old[:, 0, -1, 0].to(torch.int64)
            pos0 = torch.full((B,), old.shape[2], dtype=torch.int64, device=dev)
        h = self._chain(h_prev, input_ids, pos0)
        new = h.to(torch.float32).view(B, 1, q, 1)
        k = new if old is None else torch.cat([old, new], dim=2)
        return _Out(logits=self.logits_of(h), past_key_values=((k, k.clone()),))


def make_pair(vocab_size: int, seed: int = 0, noise: float = 0.5, dtype=torch.float32, device="cpu"):
    """(draft, target) with correlated logits so that acceptance lands around 0.5-0.9."""
    target = ReplayLM(vocab_size, table_seed=seed, noise=0.0, dtype=dtype, device=device)
    draft = ReplayLM(vocab_size, table_seed=seed, noise=noise, noise_seed=seed + 1, dtype=dtype, device=device)
    return draft, target
