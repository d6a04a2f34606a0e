"""Import the UNMODIFIED reference from /root/reference and make it deterministic
(TEST INFRASTRUCTURE ONLY; dev container only — the path does not exist on the GPU box).

Nothing is copied: the reference modules are imported in place.  Two compatibility shims
 are needed because the container has transformers 5.x while the reference pins
4.35.2:

* shim 1 — four class names that ``sampling/kvcache_model.py:8,12`` imports (used only by the
  out-of-scope beam code) were removed from transformers 5; stubs are registered before import.
* shim 2 — ``LegacyCacheAdapter`` converts between the 4.35 tuple-of-(k, v) cache that
  ``kvcache_model.py:175,381-382`` indexes/slices and the ``DynamicCache`` object HF 5 models use.

Determinism: ``TapeRNG`` replaces ``sample`` in the three modules that bound the name
(``sampling.utils``, ``sampling.kvcache_model``, ``sampling.speculative_sampling``) by the
inverse-CDF rule of ``oracle.ref_ops.icdf_sample`` and ``torch.rand`` inside
``speculative_sampling`` by reads from the same tape (layout: oracle/tape.py).
"""
from __future__ import annotations

import contextlib
import importlib
import importlib.util
import os
import sys
import warnings
from types import SimpleNamespace

import torch

from . import ref_ops, tape as tape_mod

REFERENCE_ROOT = os.environ.get("SPECDEC_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "sampling", "utils.py"))


def load_utils():
    """``sampling/utils.py`` loaded standalone (it only needs torch) — no shim."""
    path = os.path.join(REFERENCE_ROOT, "sampling", "utils.py")
    spec = importlib.util.spec_from_file_location("_reference_sampling_utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_PKG = None


def load_package():
    """The whole reference ``sampling`` package (shim 1)."""
    global _PKG
    if _PKG is not None:
        return _PKG
    import transformers
    import transformers.generation                       # noqa: F401  (lazy module must be realised first)
    import transformers.models.bloom.modeling_bloom      # noqa: F401
    for name in ("BeamSearchScorer", "BeamScorer"):
        if not hasattr(transformers, name):
            setattr(sys.modules["transformers"], name, type(name, (), {}))
    for name in ("BeamSampleDecoderOnlyOutput", "BeamSampleEncoderDecoderOutput"):
        if not hasattr(transformers.generation, name):
            setattr(sys.modules["transformers.generation"], name, type(name, (), {}))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)                # also needed for `from globals import Decoder`
    for shadow in ("sampling", "globals"):
        m = sys.modules.get(shadow)
        if m is not None and not str(getattr(m, "__file__", "")).startswith(REFERENCE_ROOT):
            raise RuntimeError(f"module {shadow!r} already imported from elsewhere: {m.__file__}")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _PKG = importlib.import_module("sampling")
    return _PKG


class LegacyCacheAdapter(torch.nn.Module):
    """shim 2: HF-5 causal LM behind the transformers-4.35 cache calling convention."""

    def __init__(self, model):
        super().__init__()
        self.inner = model
        self.config = model.config

    @property
    def device(self):
        return next(self.inner.parameters()).device

    def forward(self, input_ids, past_key_values=None, use_cache=True, **kw):
        from transformers.cache_utils import DynamicCache
        cache = None
        if past_key_values is not None:
            cache = DynamicCache(ddp_cache_data=[(k.contiguous(), v.contiguous()) for k, v in past_key_values])
        out = self.inner(input_ids, past_key_values=cache, use_cache=True)
        legacy = tuple((layer.keys, layer.values) for layer in out.past_key_values.layers)
        return SimpleNamespace(logits=out.logits, past_key_values=legacy)


class TapeRNG:
    """Feeds the reference's ``sample`` / ``torch.rand`` calls from a (iterations, 2*gamma+2) tape."""

    def __init__(self, tape: torch.Tensor, gamma: int):
        self.tape, self.gamma = tape, gamma
        self.it = 0
        self.n_sample = 0
        self.n_rand = 0

    def sample(self, probs: torch.Tensor, num_samples: int = 1):
        g = self.gamma
        c = self.n_sample
        if c < g:
            u = self.tape[self.it, c]                     # u_draft
        elif c == g:
            u = self.tape[self.it, g]                     # u_discard
        else:
            u = self.tape[self.it, 2 * g + 1]             # u_final
        tok = ref_ops.icdf_sample(probs.reshape(-1), float(u))
        self.n_sample += 1
        if c == g + 1:                                    # iteration finished
            self.it += 1
            self.n_sample = 0
            self.n_rand = 0
        return torch.tensor([[tok]], dtype=torch.long, device=probs.device)

    def rand(self, *size, **kw):
        u = self.tape[self.it, self.gamma + 1 + self.n_rand]
        self.n_rand += 1
        return u.reshape(1).clone().to(kw.get("device", "cpu"))


@contextlib.contextmanager
def patched(rng: TapeRNG):
    pkg = load_package()
    mods = [sys.modules["sampling.utils"], sys.modules["sampling.kvcache_model"],
            sys.modules["sampling.speculative_sampling"]]
    saved = [m.sample for m in mods]
    ss = sys.modules["sampling.speculative_sampling"]
    real_torch = ss.torch

    class _TorchProxy:
        def __getattr__(self, name):
            return rng.rand if name == "rand" else getattr(real_torch, name)

    try:
        for m in mods:
            m.sample = rng.sample
        ss.torch = _TorchProxy()
        yield pkg
    finally:
        for m, s in zip(mods, saved):
            m.sample = s
        ss.torch = real_torch


def run_reference(prefix, approx_model, target_model, max_len, gamma, temperature, top_k, top_p,
                  tape=None, seed=0, eos_token_id=-1, legacy_models=True):
    """Run the real ``sampling.speculative_sampling`` on the tape.  Returns (tokens, details)."""
    if tape is None:
        tape = tape_mod.make_tape(seed, max_len + 1, gamma)
    rng = TapeRNG(tape, gamma)
    a = approx_model if legacy_models else LegacyCacheAdapter(approx_model)
    t = target_model if legacy_models else LegacyCacheAdapter(target_model)
    with patched(rng) as pkg:
        out, d = pkg.speculative_sampling(prefix, a, t, eos_token_id, None, max_len, gamma=gamma,
                                          temperature=temperature, top_k=top_k, top_p=top_p, details=True)
    return out, d


class BiLDTapeRNG:
    """Feeds the ``sample`` calls of the reference's BiLD_sampling from a (cycles, 2*gamma+2) tape: draft tokens of a
    check cycle use u_draft[0..], the sample that ``target.generate(x, 1)`` throws away uses u_discard (recognised
    because the target's forward ran just before it), the target's own token uses u_final and closes the cycle."""

    def __init__(self, tape: torch.Tensor, gamma: int):
        self.tape, self.gamma = tape, gamma
        self.cycle = 0
        self.n_draft = 0
        self.discard_pending = False

    def target_called(self):
        self.discard_pending = True

    def _tok(self, probs, u):
        tok = ref_ops.icdf_sample(probs.reshape(-1), float(u))
        return torch.tensor([[tok]], dtype=torch.long, device=probs.device)

    def kv_sample(self, probs: torch.Tensor, num_samples: int = 1):
        if self.discard_pending:
            self.discard_pending = False
            return self._tok(probs, self.tape[self.cycle, self.gamma])
        u = self.tape[self.cycle, self.n_draft]
        self.n_draft += 1
        return self._tok(probs, u)

    def final_sample(self, probs: torch.Tensor, num_samples: int = 1):
        u = self.tape[self.cycle, 2 * self.gamma + 1]
        self.cycle += 1
        self.n_draft = 0
        return self._tok(probs, u)


class _NotifyForward(torch.nn.Module):
    """Transparent wrapper that tells the tape when the target model runs."""

    def __init__(self, inner, callback):
        super().__init__()
        self.inner, self.callback = inner, callback
        self.config = inner.config

    @property
    def device(self):
        return self.inner.device

    def forward(self, *args, **kwargs):
        self.callback()
        return self.inner(*args, **kwargs)


def run_reference_bild(prefix, approx_model, target_model, max_len, gamma, fallback_thres, rollback_thres, temperature,
                       top_k, top_p, tape=None, seed=0, eos_token_id=-1, legacy_models=True):
    """Run the real ``sampling.speculative_sampling.BiLD_sampling`` (speculative_sampling.py:1718-1873) on the tape."""
    if tape is None:
        tape = tape_mod.make_tape(seed, max_len + 1, gamma)
    rng = BiLDTapeRNG(tape, gamma)
    a = approx_model if legacy_models else LegacyCacheAdapter(approx_model)
    t = target_model if legacy_models else LegacyCacheAdapter(target_model)
    t = _NotifyForward(t, rng.target_called)
    pkg = load_package()
    kv, ss = sys.modules["sampling.kvcache_model"], sys.modules["sampling.speculative_sampling"]
    saved = (kv.sample, ss.sample)
    try:
        kv.sample, ss.sample = rng.kv_sample, rng.final_sample
        out, d = ss.BiLD_sampling(prefix, a, t, gamma, eos_token_id, None, fallback_thres, rollback_thres, max_len,
                                  temperature=temperature, top_k=top_k, top_p=top_p, details=True)
    finally:
        kv.sample, ss.sample = saved
    return out, d


def multi_block(gamma: int, width: int) -> int:
    """Uniforms one iteration of multi_speculative_sampling(strategy='iid') can consume."""
    return 2 * width * gamma + width + 1


class MultiTapeRNG:
    """Feeds the reference's multi_speculative_sampling (iid): per iteration a block of 2*W*gamma + W + 1 uniforms —
    [gamma draft calls x W rows | W discarded target samples | the accept tests IN THE ORDER THE REFERENCE DRAWS THEM
    (draft 0 position 0, 1, .. until its first reject, then draft 1, ..; speculative_sampling.py:1616-1634) | final]."""

    def __init__(self, tape: torch.Tensor, gamma: int, width: int):
        self.tape, self.gamma, self.width = tape, gamma, width
        self.it = 0
        self.n_kv = 0
        self.n_rand = 0

    def _rows(self, probs, us):
        toks = [ref_ops.icdf_sample(probs[w], float(us[w])) for w in range(probs.shape[0])]
        return torch.tensor(toks, dtype=torch.long, device=probs.device).view(-1, 1)

    def kv_sample(self, probs: torch.Tensor, num_samples: int = 1):
        W = self.width
        assert probs.shape[0] == W
        us = self.tape[self.it, self.n_kv * W:(self.n_kv + 1) * W]          # calls 0..gamma-1: drafts, call gamma: discarded
        self.n_kv += 1
        return self._rows(probs, us)

    def final_sample(self, probs: torch.Tensor, num_samples: int = 1):
        u = self.tape[self.it, multi_block(self.gamma, self.width) - 1]
        tok = self._rows(probs.reshape(1, -1), [u])            # raises on an empty row: the reference then retries with p
        self.it += 1
        self.n_kv = 0
        self.n_rand = 0
        return tok

    def rand(self, *size, **kw):
        W, g = self.width, self.gamma
        u = self.tape[self.it, g * W + W + self.n_rand]
        self.n_rand += 1
        return u.reshape(1).clone().to(kw.get("device", "cpu"))


def run_reference_multi(prefix, approx_model, target_model, max_len, gamma, width, temperature, top_k, top_p, tape,
                        eos_token_id=-1, legacy_models=True):
    """Run the real ``multi_speculative_sampling(strategy='iid')`` (speculative_sampling.py:1379-1716) on the tape."""
    rng = MultiTapeRNG(tape, gamma, width)
    a = approx_model if legacy_models else LegacyCacheAdapter(approx_model)
    t = target_model if legacy_models else LegacyCacheAdapter(target_model)
    load_package()
    kv, ss = sys.modules["sampling.kvcache_model"], sys.modules["sampling.speculative_sampling"]
    saved = (kv.sample, ss.sample, ss.torch)
    real_torch = ss.torch

    class _TorchProxy:
        def __getattr__(self, name):
            return rng.rand if name == "rand" else getattr(real_torch, name)

    try:
        kv.sample, ss.sample, ss.torch = rng.kv_sample, rng.final_sample, _TorchProxy()
        out, d = ss.multi_speculative_sampling(prefix, a, t, eos_token_id, None, max_len, gamma=gamma, width=width,
                                               strategy="iid", temperature=temperature, top_k=top_k, top_p=top_p,
                                               details=True)
    finally:
        kv.sample, ss.sample, ss.torch = saved
    return out, d


class V2TapeRNG:
    """Feeds the reference's speculative_sampling_v2 (speculative_sampling.py:2080-2194) from the (iterations, 2*gamma+2)
    tape: per iteration gamma draft samples (u_draft), the accept uniforms it draws lazily (u_acc, a prefix is used),
    then exactly ONE more sample — the residual (:2161) or the bonus (:2178) — which uses u_final and closes the
    iteration.  Slot gamma (u_discard) is unused: v2 has no discarded sample."""

    def __init__(self, tape: torch.Tensor, gamma: int):
        self.tape, self.gamma = tape, gamma
        self.it = 0
        self.n_sample = 0
        self.n_rand = 0

    def sample(self, probs: torch.Tensor, num_samples: int = 1):
        g, c = self.gamma, self.n_sample
        u = self.tape[self.it, c] if c < g else self.tape[self.it, 2 * g + 1]
        tok = ref_ops.icdf_sample(probs.reshape(-1), float(u))            # raises 'prob error' on an empty residual
        self.n_sample += 1
        if c == g:
            self.it += 1
            self.n_sample = 0
            self.n_rand = 0
        return torch.tensor([[tok]], dtype=torch.long, device=probs.device)

    def rand(self, *size, **kw):
        u = self.tape[self.it, self.gamma + 1 + self.n_rand]
        self.n_rand += 1
        return u.reshape(1).clone().to(kw.get("device", "cpu"))


def run_reference_v2(prefix, approx_model, target_model, max_len, gamma, temperature, top_k, top_p, tape=None, seed=0):
    """Run the real ``sampling.speculative_sampling.speculative_sampling_v2`` (speculative_sampling.py:2080-2194, no KV
    cache, so no cache shim is involved) on the tape.  Returns (tokens, details)."""
    if tape is None:
        tape = tape_mod.make_tape(seed, max_len + 1, gamma)
    rng = V2TapeRNG(tape, gamma)
    load_package()
    ss = sys.modules["sampling.speculative_sampling"]
    saved = (ss.sample, ss.torch)
    real_torch = ss.torch

    class _TorchProxy:
        def __getattr__(self, name):
            return rng.rand if name == "rand" else getattr(real_torch, name)

    try:
        ss.sample, ss.torch = rng.sample, _TorchProxy()
        out, d = ss.speculative_sampling_v2(prefix, approx_model, target_model, max_len, gamma=gamma,
                                            temperature=temperature, top_k=top_k, top_p=top_p, details=True)
    finally:
        ss.sample, ss.torch = saved
    return out, d


def run_reference_ar(x, model, N, temperature, top_k, top_p, uniforms, eos_token_id=-1):
    """Run the real ``sampling.autoregressive_sampling.autoregressive_sampling`` (autoregressive_sampling.py:9-61) with
    its ``sample`` fed from `uniforms` (one per generated token).  Returns the token tensor."""
    load_package()
    ar = importlib.import_module("sampling.autoregressive_sampling")
    state = {"i": 0}

    def tape_sample(probs, num_samples=1):
        tok = ref_ops.icdf_sample(probs.reshape(-1), float(uniforms[state["i"]]))
        state["i"] += 1
        return torch.tensor([[tok]], dtype=torch.long, device=probs.device)

    saved = ar.sample
    try:
        ar.sample = tape_sample
        out = ar.autoregressive_sampling(x, model, N, eos_token_id, temperature=temperature, top_k=top_k, top_p=top_p)
    finally:
        ar.sample = saved
    return out
