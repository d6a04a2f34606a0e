"""Tensor-level wrappers over the C ABI (include/specdec_b200.h).

Every function takes CUDA tensors, launches on torch's current stream and returns without
synchronising.  Numerical faults are accumulated in a device-side flag word (`ErrFlag`) that the
caller checks when it wants to (once per public-API call), mirroring where the reference raises.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _cabi

SD_F32, SD_BF16, SD_F16 = 0, 1, 2
_DT = {torch.float32: SD_F32, torch.bfloat16: SD_BF16, torch.float16: SD_F16}

ERR_NORM_LOGITS, ERR_PROB, ERR_ZERO_Q, ERR_BAD_TOKEN = 1, 2, 4, 8


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the speculative-decoding kernels have no CPU path")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


NORM_WORKSPACE_BYTES = 16 + 65536 // 8        # SD_NORM_WORKSPACE_BYTES (include/specdec_b200.h)


class ErrFlag:
    """Device int32 word the kernels OR error bits into, plus the norm kernels' scheduler workspace
    (SD_NORM_WORKSPACE_BYTES, include/specdec_b200.h).  One object serves launches that are ordered after one another
    (one stream, one CUDA graph); launches that may run concurrently need their own."""

    def __init__(self, device, shared: bool = False):
        self.t = torch.zeros(1, dtype=torch.int32, device=device)
        self.ws = torch.zeros(NORM_WORKSPACE_BYTES // 4, dtype=torch.int32, device=device)
        self.shared = shared          # per-device default flag: any stream may use it, so its workspace is not used

    def ptr(self) -> int:
        return self.t.data_ptr()

    def ws_ptr(self) -> int:
        return self.ws.data_ptr()

    def check(self) -> None:
        """Synchronising read; raises the reference's exceptions and clears the flag."""
        bits = int(self.t.item())
        if bits == 0:
            return
        self.t.zero_()
        if bits & ERR_NORM_LOGITS:
            raise RuntimeError("norm logits error")          # reference sampling/utils.py:207
        if bits & ERR_PROB:
            raise RuntimeError("prob error")                 # reference sampling/utils.py:224
        raise RuntimeError("s")                              # reference speculative_sampling.py:2046


_flags = {}


def default_flag(device) -> ErrFlag:
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _flags:
        _flags[key] = ErrFlag(torch.device("cuda", key), shared=True)
    return _flags[key]


_workspaces = {}


def _default_workspace(device) -> int:
    """Scheduler workspace of the calling stream (launches on different streams may overlap, so each stream gets its
    own).  None can be created while the stream is capturing: 0 (NULL) then selects the one-cluster-per-row kernel —
    pass an ErrFlag made before the capture to get the persistent kernel inside a graph."""
    dev = torch.device(device)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), _stream())
    ws = _workspaces.get(key)
    if ws is None:
        if torch.cuda.is_current_stream_capturing():
            return 0
        ws = _workspaces[key] = torch.zeros(NORM_WORKSPACE_BYTES // 4, dtype=torch.int32, device=dev)
    return ws.data_ptr()


NORM_NO_PIPELINE, NORM_FORCE_GENERAL, NORM_NO_RING = 1, 2, 4


def _norm_flags(general: bool, pipeline) -> int:
    """pipeline: True = the library's choice (ring kernel where a row fits one CTA, else the cluster pipeline),
    'cluster' = skip the ring kernel (SD_NORM_NO_RING), False = one-cluster-per-row kernel (SD_NORM_NO_PIPELINE)."""
    f = NORM_FORCE_GENERAL if general else 0
    if pipeline == "cluster":
        f |= NORM_NO_RING
    elif not pipeline:
        f |= NORM_NO_PIPELINE
    return f


class CompactRows:
    """Compact (index, probability) lists of top-k filtered rows: `cnt` (n_rows,), `idx` / `val` (n_rows, cap).
    `view(first_row, row_stride)` addresses logical row r of a kernel call at compact row first_row + r * row_stride."""

    def __init__(self, n_rows: int, device, cap: int = 64):
        self.cnt = torch.full((n_rows,), -1, dtype=torch.int32, device=device)
        self.idx = torch.zeros(n_rows, cap, dtype=torch.int32, device=device)
        self.val = torch.zeros(n_rows, cap, dtype=torch.float32, device=device)
        self.cap = cap

    def view(self, first_row: int = 0, row_stride: int = 1) -> "_cabi.Compact":
        return _cabi.Compact(self.cnt.data_ptr() + 4 * first_row, self.idx.data_ptr() + 4 * first_row * self.cap,
                             self.val.data_ptr() + 4 * first_row * self.cap, self.cap, row_stride)

    def to_dense(self, V: int) -> torch.Tensor:
        """(n_rows, V) dense rows rebuilt from the lists (rows with cnt < 0 are NaN) — test helper."""
        out = torch.zeros(self.cnt.numel(), V, dtype=torch.float32, device=self.cnt.device)
        for r in range(self.cnt.numel()):
            c = int(self.cnt[r])
            if c < 0:
                out[r] = float("nan")
            else:
                out[r, self.idx[r, :c].long()] = self.val[r, :c]
        return out


def _cref(c):
    import ctypes
    return None if c is None else ctypes.byref(c)


def set_tuning(norm_cluster: int = 0, norm_threads: int = 0, verify_cluster: int = 0) -> None:
    _cabi.load().sd_set_tuning(norm_cluster, norm_threads, verify_cluster)


def set_pdl(enable: bool) -> None:
    """Programmatic dependent launch for the pipelined norm kernel and the sparse verify kernel (see the header)."""
    _cabi.load().sd_set_pdl(1 if enable else 0)


def _rows2d(logits: torch.Tensor) -> torch.Tensor:
    if logits.dim() != 2:
        raise AssertionError("logits must be 2-D (rows, vocab)")      # reference utils.py:194
    if logits.stride(1) != 1:
        logits = logits.contiguous()
    return logits


def norm_probs(logits: torch.Tensor, temperature: float, top_k: int, top_p: float,
               out: Optional[torch.Tensor] = None, err: Optional[ErrFlag] = None,
               general: bool = False, pipeline: bool = True, compact=None) -> torch.Tensor:
    """(rows, V) logits (fp32/bf16/fp16) -> (rows, V) fp32 probabilities.  Kernel 1."""
    _require_cuda(logits, "logits")
    x = _rows2d(logits)
    rows, V = x.shape
    if out is None:
        out = torch.empty(rows, V, dtype=torch.float32, device=x.device)
    assert out.dtype == torch.float32 and out.shape == (rows, V) and out.stride(1) == 1
    ws = err.ws_ptr() if (err is not None and not err.shared) else _default_workspace(x.device)
    err = err or default_flag(x.device)
    lib = _cabi.load()
    k = int(top_k) if top_k else 0
    p = float(top_p) if top_p else 0.0
    flags = _norm_flags(general, pipeline)
    rc = lib.sd_norm_probs(x.data_ptr(), _DT[x.dtype], rows, V, x.stride(0), float(temperature), k, p,
                           out.data_ptr(), out.stride(0), _cref(compact), err.ptr(), flags, ws, _stream())
    _cabi.check(rc, "sd_norm_probs")
    return out


def norm_sample(logits: torch.Tensor, temperature: float, top_k: int, top_p: float, u: torch.Tensor,
                probs_out: Optional[torch.Tensor] = None, tok_out: Optional[torch.Tensor] = None,
                err: Optional[ErrFlag] = None, general: bool = False, pipeline: bool = True,
                compact=None) -> torch.Tensor:
    """Kernel 1b: probabilities (optional, written to probs_out) and one sampled token per row."""
    _require_cuda(logits, "logits")
    x = _rows2d(logits)
    rows, V = x.shape
    assert u.is_cuda and u.dtype == torch.float32 and u.numel() == rows and u.is_contiguous()
    if tok_out is None:
        tok_out = torch.empty(rows, dtype=torch.int64, device=x.device)
    assert tok_out.dtype == torch.int64 and tok_out.numel() == rows and tok_out.is_contiguous()
    if probs_out is not None:
        assert probs_out.dtype == torch.float32 and probs_out.shape == (rows, V) and probs_out.stride(1) == 1
    ws = err.ws_ptr() if (err is not None and not err.shared) else _default_workspace(x.device)
    err = err or default_flag(x.device)
    flags = _norm_flags(general, pipeline)
    rc = _cabi.load().sd_norm_sample(x.data_ptr(), _DT[x.dtype], rows, V, x.stride(0), float(temperature),
                                     int(top_k or 0), float(top_p or 0.0), _ptr(probs_out),
                                     probs_out.stride(0) if probs_out is not None else V, u.data_ptr(),
                                     tok_out.data_ptr(), _cref(compact), err.ptr(), flags, ws, _stream())
    _cabi.check(rc, "sd_norm_sample")
    return tok_out


def sample_rows(probs: torch.Tensor, u: torch.Tensor, tok_out: Optional[torch.Tensor] = None,
                err: Optional[ErrFlag] = None) -> torch.Tensor:
    """One inverse-CDF draw per row of non-negative fp32 weights."""
    _require_cuda(probs, "probs")
    x = _rows2d(probs)
    if x.dtype != torch.float32:
        x = x.float()
    rows, V = x.shape
    assert u.is_cuda and u.dtype == torch.float32 and u.numel() == rows and u.is_contiguous()
    if tok_out is None:
        tok_out = torch.empty(rows, dtype=torch.int64, device=x.device)
    err = err or default_flag(x.device)
    rc = _cabi.load().sd_sample(x.data_ptr(), rows, V, x.stride(0), u.data_ptr(), tok_out.data_ptr(), err.ptr(), _stream())
    _cabi.check(rc, "sd_sample")
    return tok_out


def max_fn(x: torch.Tensor) -> torch.Tensor:
    _require_cuda(x, "x")
    squeeze = x.dim() == 1
    x2 = x.unsqueeze(0) if squeeze else x
    x2 = _rows2d(x2.float() if x2.dtype != torch.float32 else x2)
    out = torch.empty_like(x2)
    rc = _cabi.load().sd_max_fn(x2.data_ptr(), x2.shape[0], x2.shape[1], x2.stride(0), out.data_ptr(), out.stride(0), _stream())
    _cabi.check(rc, "sd_max_fn")
    return out[0] if squeeze else out


def verify(p_probs: torch.Tensor, q_probs: torch.Tensor, draft_tok: torch.Tensor, u_acc: torch.Tensor,
           u_final: torch.Tensor, strict: bool = False, n_accepted: Optional[torch.Tensor] = None,
           next_tok: Optional[torch.Tensor] = None, ratios: Optional[torch.Tensor] = None,
           tie_count: Optional[torch.Tensor] = None, tokens: Optional[torch.Tensor] = None,
           seq_len: Optional[torch.Tensor] = None, active: Optional[torch.Tensor] = None,
           err: Optional[ErrFlag] = None, p_compact=None, p_cmp_req_stride: int = 0, q_compact=None,
           q_cmp_req_stride: int = 0, stats: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Kernel 2.  p_probs (B, gamma+1, V), q_probs (B, gamma, V) fp32 (last dim contiguous),
    draft_tok (B, gamma) int64, u_acc (B, gamma), u_final (B,).  Returns (n_accepted, next_tok)."""
    _require_cuda(p_probs, "p_probs")
    B, g1, V = p_probs.shape
    gamma = g1 - 1
    assert q_probs.shape == (B, gamma, V) and p_probs.dtype == q_probs.dtype == torch.float32
    assert p_probs.stride(2) == 1 and q_probs.stride(2) == 1
    assert draft_tok.dtype == torch.int64 and draft_tok.shape == (B, gamma) and draft_tok.stride(1) == 1
    assert u_acc.dtype == torch.float32 and u_acc.shape == (B, gamma) and u_acc.stride(1) == 1
    assert u_final.dtype == torch.float32 and u_final.numel() == B and u_final.is_contiguous()
    dev = p_probs.device
    if n_accepted is None:
        n_accepted = torch.empty(B, dtype=torch.int32, device=dev)
    if next_tok is None:
        next_tok = torch.empty(B, dtype=torch.int64, device=dev)
    err = err or default_flag(dev)
    rc = _cabi.load().sd_verify(
        p_probs.data_ptr(), p_probs.stride(0), p_probs.stride(1), q_probs.data_ptr(), q_probs.stride(0), q_probs.stride(1),
        draft_tok.data_ptr(), draft_tok.stride(0), u_acc.data_ptr(), u_acc.stride(0), u_final.data_ptr(),
        B, gamma, V, 1 if strict else 0, n_accepted.data_ptr(), next_tok.data_ptr(), _ptr(ratios), _ptr(tie_count),
        _ptr(tokens), tokens.stride(0) if tokens is not None else 0, _ptr(seq_len), _ptr(active),
        _cref(p_compact), p_cmp_req_stride, _cref(q_compact), q_cmp_req_stride, _ptr(stats), err.ptr(), _stream())
    _cabi.check(rc, "sd_verify")
    return n_accepted, next_tok


def verify_multi(p_probs: torch.Tensor, q_probs: torch.Tensor, draft_tok: torch.Tensor, u_acc: torch.Tensor,
                 u_final: torch.Tensor, ratios: Optional[torch.Tensor] = None, err: Optional[ErrFlag] = None):
    """Kernel 2, multi-draft variant (sd_verify_multi).  p_probs (B, W, gamma+1, V), q_probs (B, W, gamma, V) fp32,
    draft_tok (B, W, gamma) int64, u_acc (B, >= W*gamma) accept uniforms in the reference's drawing order, u_final (B,).
    Returns (choice, n_accepted, next_tok)."""
    _require_cuda(p_probs, "p_probs")
    B, W, g1, V = p_probs.shape
    gamma = g1 - 1
    assert q_probs.shape == (B, W, gamma, V) and p_probs.dtype == q_probs.dtype == torch.float32
    assert p_probs.stride(3) == 1 and q_probs.stride(3) == 1
    assert draft_tok.dtype == torch.int64 and draft_tok.shape == (B, W, gamma) and draft_tok.stride(2) == 1
    assert u_acc.dtype == torch.float32 and u_acc.dim() == 2 and u_acc.shape[0] == B and u_acc.shape[1] >= W * gamma and u_acc.stride(1) == 1
    assert u_final.dtype == torch.float32 and u_final.numel() == B and u_final.is_contiguous()
    dev = p_probs.device
    choice = torch.empty(B, dtype=torch.int32, device=dev)
    n_acc = torch.empty(B, dtype=torch.int32, device=dev)
    nxt = torch.empty(B, dtype=torch.int64, device=dev)
    err = err or default_flag(dev)
    rc = _cabi.load().sd_verify_multi(
        p_probs.data_ptr(), p_probs.stride(0), p_probs.stride(1), p_probs.stride(2),
        q_probs.data_ptr(), q_probs.stride(0), q_probs.stride(1), q_probs.stride(2),
        draft_tok.data_ptr(), draft_tok.stride(0), draft_tok.stride(1), u_acc.data_ptr(), u_acc.stride(0), u_final.data_ptr(),
        B, W, gamma, V, choice.data_ptr(), n_acc.data_ptr(), nxt.data_ptr(), _ptr(ratios), err.ptr(), _stream())
    _cabi.check(rc, "sd_verify_multi")
    return choice, n_acc, nxt


def verify_bild(p_probs: torch.Tensor, draft_tok: torch.Tensor, rollback_thres: float, u_final: torch.Tensor,
                n_check: Optional[torch.Tensor] = None, nll: Optional[torch.Tensor] = None, err: Optional[ErrFlag] = None,
                q_probs: Optional[torch.Tensor] = None, fallback_thres: float = 0.0, n_drafted: Optional[torch.Tensor] = None,
                tokens: Optional[torch.Tensor] = None, seq_len: Optional[torch.Tensor] = None,
                limit: Optional[torch.Tensor] = None, active: Optional[torch.Tensor] = None,
                n_accepted: Optional[torch.Tensor] = None, next_tok: Optional[torch.Tensor] = None, q_compact=None,
                q_cmp_req_stride: int = 0, eos_token_id: int = -1):
    """Kernel 2, BiLD variant (sd_verify_bild).  p_probs (B, C+1, V) fp32 target rows of the C unchecked draft tokens
    draft_tok (B, C) int64 (+ the row after them); keeps tokens while -log p[token] <= rollback_thres and samples the
    target's own token from the first row it did not keep.  Engine mode: q_probs (B, C, V) + fallback_thres derive the
    number of tokens the reference would have drafted (-> n_drafted), tokens / seq_len / limit / active as in `verify`.
    Returns (n_kept, next_tok)."""
    _require_cuda(p_probs, "p_probs")
    B, c1, V = p_probs.shape
    C = c1 - 1
    assert p_probs.dtype == torch.float32 and p_probs.stride(2) == 1
    assert draft_tok.dtype == torch.int64 and draft_tok.shape == (B, C) and draft_tok.stride(1) == 1
    assert u_final.dtype == torch.float32 and u_final.numel() == B and u_final.is_contiguous()
    if q_probs is not None:
        assert q_probs.shape == (B, C, V) and q_probs.dtype == torch.float32 and q_probs.stride(2) == 1
    dev = p_probs.device
    if n_accepted is None:
        n_accepted = torch.empty(B, dtype=torch.int32, device=dev)
    if next_tok is None:
        next_tok = torch.empty(B, dtype=torch.int64, device=dev)
    err = err or default_flag(dev)
    rc = _cabi.load().sd_verify_bild(
        p_probs.data_ptr(), p_probs.stride(0), p_probs.stride(1), _ptr(q_probs),
        q_probs.stride(0) if q_probs is not None else 0, q_probs.stride(1) if q_probs is not None else 0,
        draft_tok.data_ptr(), draft_tok.stride(0), _ptr(n_check), C, float(fallback_thres), float(rollback_thres),
        u_final.data_ptr(), B, V, n_accepted.data_ptr(), next_tok.data_ptr(), _ptr(nll), _ptr(n_drafted), _ptr(tokens),
        tokens.stride(0) if tokens is not None else 0, _ptr(seq_len), _ptr(limit), _ptr(active), _cref(q_compact),
        int(q_cmp_req_stride), int(eos_token_id), err.ptr(), _stream())
    _cabi.check(rc, "sd_verify_bild")
    return n_accepted, next_tok


def norm_sample_verify(logits: torch.Tensor, temperature: float, top_k: int, top_p: float, u: torch.Tensor,
                       probs_out: torch.Tensor, tok_out: torch.Tensor, compact, rows_per_request: int,
                       request_counters: torch.Tensor, p_probs: torch.Tensor, q_probs: torch.Tensor,
                       draft_tok: torch.Tensor, u_acc: torch.Tensor, u_final: torch.Tensor, n_accepted: torch.Tensor,
                       next_tok: torch.Tensor, p_compact, p_cmp_req_stride: int, q_compact, q_cmp_req_stride: int,
                       err: ErrFlag, strict: bool = False, ratios: Optional[torch.Tensor] = None,
                       tie_count: Optional[torch.Tensor] = None, tokens: Optional[torch.Tensor] = None,
                       seq_len: Optional[torch.Tensor] = None, active: Optional[torch.Tensor] = None,
                       stats: Optional[torch.Tensor] = None, pipeline: bool = True) -> None:
    """Kernels 1 + 2 in one launch (sd_norm_sample_verify): `norm_sample` over (B * rows_per_request, V) logits where
    request b owns rows b * rows_per_request ..; each request is verified (arguments as `verify`) as soon as its last
    row is normalised.  `request_counters`: (B,) int32 zeros, left zeroed.  Falls back to two launches inside the
    library where the persistent kernel does not apply."""
    _require_cuda(logits, "logits")
    x = _rows2d(logits)
    rows, V = x.shape
    B, g1, _ = p_probs.shape
    gamma = g1 - 1
    assert rows == B * rows_per_request and request_counters.dtype == torch.int32 and request_counters.numel() >= B
    assert u.is_cuda and u.dtype == torch.float32 and u.numel() == rows and u.is_contiguous()
    assert tok_out.dtype == torch.int64 and tok_out.numel() == rows and tok_out.is_contiguous()
    assert probs_out.dtype == torch.float32 and probs_out.shape == (rows, V) and probs_out.stride(1) == 1
    assert q_probs.shape == (B, gamma, V) and p_probs.stride(2) == 1 and q_probs.stride(2) == 1
    assert draft_tok.dtype == torch.int64 and draft_tok.shape == (B, gamma) and draft_tok.stride(1) == 1
    assert u_acc.dtype == torch.float32 and u_acc.shape == (B, gamma) and u_acc.stride(1) == 1
    assert u_final.dtype == torch.float32 and u_final.numel() == B and u_final.is_contiguous()
    assert err is not None and not err.shared, "the fused launch needs an ErrFlag of its own (scheduler workspace)"
    import ctypes
    va = _cabi.VerifyArgs(
        p_probs.data_ptr(), p_probs.stride(0), p_probs.stride(1), q_probs.data_ptr(), q_probs.stride(0), q_probs.stride(1),
        draft_tok.data_ptr(), draft_tok.stride(0), u_acc.data_ptr(), u_acc.stride(0), u_final.data_ptr(),
        B, gamma, V, 1 if strict else 0, n_accepted.data_ptr(), next_tok.data_ptr(), _ptr(ratios), _ptr(tie_count),
        _ptr(tokens), tokens.stride(0) if tokens is not None else 0, _ptr(seq_len), _ptr(active),
        ctypes.addressof(p_compact) if p_compact is not None else None, p_cmp_req_stride,
        ctypes.addressof(q_compact) if q_compact is not None else None, q_cmp_req_stride, _ptr(stats))
    flags = _norm_flags(False, pipeline)
    rc = _cabi.load().sd_norm_sample_verify(
        x.data_ptr(), _DT[x.dtype], rows, V, x.stride(0), float(temperature), int(top_k or 0), float(top_p or 0.0),
        probs_out.data_ptr(), probs_out.stride(0), u.data_ptr(), tok_out.data_ptr(), _cref(compact), ctypes.byref(va),
        int(rows_per_request), request_counters.data_ptr(), err.ptr(), flags, err.ws_ptr(), _stream())
    _cabi.check(rc, "sd_norm_sample_verify")


def kv_append(k_new: torch.Tensor, v_new: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor,
              write_pos: torch.Tensor) -> None:
    """cache[b, h, write_pos[b] + j] = new[b, h, j] for K and V.  new: (B, H, q, D) any strides with a
    contiguous last dim; caches: (B, H, S, D) contiguous."""
    B, H, q, D = k_new.shape
    S = k_cache.shape[2]
    assert k_cache.is_contiguous() and v_cache.is_contiguous() and k_cache.shape == (B, H, S, D)
    if k_new.stride(3) != 1 or v_new.stride() != k_new.stride():
        k_new, v_new = k_new.contiguous(), v_new.contiguous()
    if k_new.dtype != k_cache.dtype:
        k_new, v_new = k_new.to(k_cache.dtype), v_new.to(v_cache.dtype)
    rc = _cabi.load().sd_kv_append(k_new.data_ptr(), v_new.data_ptr(), k_new.stride(0), k_new.stride(1), k_new.stride(2),
                                   k_cache.data_ptr(), v_cache.data_ptr(), write_pos.data_ptr(), B, H, q, D, S,
                                   k_cache.element_size(), _stream())
    _cabi.check(rc, "sd_kv_append")


def kv_select(k_cache: torch.Tensor, v_cache: torch.Tensor, W: int, choice: torch.Tensor, start: torch.Tensor,
              start_stride: int, count: torch.Tensor, max_count: int, active: Optional[torch.Tensor] = None,
              active_stride: int = 1) -> None:
    """Multi-draft rollback on static caches (B*W, H, S, D): copy positions start .. start+count-1 of every request's
    winning row over its other rows (sd_kv_select)."""
    R, H, S, D = k_cache.shape
    assert R % W == 0 and k_cache.is_contiguous() and v_cache.is_contiguous()
    rc = _cabi.load().sd_kv_select(k_cache.data_ptr(), v_cache.data_ptr(), R // W, W, H, S, D, k_cache.element_size(),
                                   int(max_count), choice.data_ptr(), start.data_ptr(), int(start_stride), count.data_ptr(),
                                   _ptr(active), int(active_stride), _stream())
    _cabi.check(rc, "sd_kv_select")


class LayerTable:
    """Device arrays of the per-layer K / V cache pointers of one model (for sd_kv_select_layers)."""

    def __init__(self, k_caches, v_caches):
        self.k, self.v = list(k_caches), list(v_caches)            # keep the tensors alive
        dev = self.k[0].device
        self.shape, self.elem_size = tuple(self.k[0].shape), self.k[0].element_size()
        assert all(tuple(t.shape) == self.shape and t.is_contiguous() for t in self.k + self.v)
        self.k_ptrs = torch.tensor([t.data_ptr() for t in self.k], dtype=torch.int64, device=dev)
        self.v_ptrs = torch.tensor([t.data_ptr() for t in self.v], dtype=torch.int64, device=dev)


def kv_select_layers(table: LayerTable, W: int, choice: torch.Tensor, start: torch.Tensor, start_stride: int, count: torch.Tensor,
                     max_count: int, active: Optional[torch.Tensor] = None, active_stride: int = 1) -> None:
    """kv_select for every layer of a model in ONE launch (sd_kv_select_layers)."""
    R, H, S, D = table.shape
    assert R % W == 0
    rc = _cabi.load().sd_kv_select_layers(table.k_ptrs.data_ptr(), table.v_ptrs.data_ptr(), len(table.k), R // W, W, H, S, D,
                                          table.elem_size, int(max_count), choice.data_ptr(), start.data_ptr(), int(start_stride),
                                          count.data_ptr(), _ptr(active), int(active_stride), _stream())
    _cabi.check(rc, "sd_kv_select_layers")


def multi_commit(tokens: torch.Tensor, seq_len: torch.Tensor, W: int, choice: torch.Tensor, n_acc: torch.Tensor,
                 next_tok: torch.Tensor, active: Optional[torch.Tensor] = None) -> None:
    """Token append of the multi-draft loop for all W rows of every request (sd_multi_commit)."""
    R, S = tokens.shape
    rc = _cabi.load().sd_multi_commit(tokens.data_ptr(), tokens.stride(0), seq_len.data_ptr(), R // W, W, choice.data_ptr(),
                                      n_acc.data_ptr(), next_tok.data_ptr(), _ptr(active), S, _stream())
    _cabi.check(rc, "sd_multi_commit")


def build_step(tokens: torch.Tensor, seq_len: torch.Tensor, offset: int, q: int, prev_tok: Optional[torch.Tensor],
               S: int, input_ids: torch.Tensor, position_ids: torch.Tensor, write_pos: torch.Tensor,
               mask: Optional[torch.Tensor]) -> None:
    B = tokens.shape[0]
    rc = _cabi.load().sd_build_step(tokens.data_ptr(), tokens.stride(0), seq_len.data_ptr(), int(offset), int(q),
                                    _ptr(prev_tok), B, int(S), input_ids.data_ptr(), position_ids.data_ptr(),
                                    write_pos.data_ptr(), _ptr(mask), _stream())
    _cabi.check(rc, "sd_build_step")
