"""ctypes binding of include/specdec_b200.h — the only way the Python host code reaches the kernels.

There is deliberately no fallback: if the shared library is missing or a CUDA device is absent, the
product path raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_VARIANT = os.environ.get("SD_LIB_VARIANT", "")          # "prof": the copy with the ring kernel's timeline probes (build.py)
LIB_PATH = os.path.join(_HERE, "libspecdec_b200.so" if not _VARIANT else f"libspecdec_b200_{_VARIANT}.so")

i32, i64, f32, vp = C.c_int, C.c_int64, C.c_float, C.c_void_p

# name -> (restype, argtypes); must list every symbol include/specdec_b200.h declares
SIGNATURES = {
    "sd_version": (i32, []),
    "sd_last_error": (C.c_char_p, []),
    "sd_set_tuning": (None, [i32, i32, i32]),
    "sd_debug_set_prof": (None, [vp]),
    "sd_set_pdl": (None, [i32]),
    "sd_norm_probs": (i32, [vp, i32, i64, i64, i64, f32, i32, f32, vp, i64, vp, vp, i32, vp, vp]),
    "sd_norm_sample": (i32, [vp, i32, i64, i64, i64, f32, i32, f32, vp, i64, vp, vp, vp, vp, i32, vp, vp]),
    "sd_norm_sample_verify": (i32, [vp, i32, i64, i64, i64, f32, i32, f32, vp, i64, vp, vp, vp, vp, i32, vp, vp, i32, vp, vp]),
    "sd_verify_multi": (i32, [vp, i64, i64, i64, vp, i64, i64, i64, vp, i64, i64, vp, i64, vp, i32, i32, i32, i64, vp, vp, vp, vp, vp, vp]),
    "sd_verify_bild": (i32, [vp, i64, i64, vp, i64, i64, vp, i64, vp, i32, f32, f32, vp, i32, i64, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, i64, i64, vp, vp]),
    "sd_sample": (i32, [vp, i64, i64, i64, vp, vp, vp, vp]),
    "sd_verify": (i32, [vp, i64, i64, vp, i64, i64, vp, i64, vp, i64, vp, i32, i32, i64, i32,
                        vp, vp, vp, vp, vp, i64, vp, vp, vp, i64, vp, i64, vp, vp, vp]),
    "sd_max_fn": (i32, [vp, i64, i64, i64, vp, i64, vp]),
    "sd_kv_append": (i32, [vp, vp, i64, i64, i64, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "sd_kv_select": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, i32, vp, vp, i32, vp]),
    "sd_kv_select_layers": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, i32, vp, vp, i32, vp]),
    "sd_multi_commit": (i32, [vp, i64, vp, i32, i32, vp, vp, vp, vp, i32, vp]),
    "sd_build_step": (i32, [vp, i64, vp, i32, i32, vp, i32, i32, vp, vp, vp, vp, vp]),
}

class VerifyArgs(C.Structure):
    """struct sd_verify_args (include/specdec_b200.h): the arguments of sd_verify, for sd_norm_sample_verify"""
    _fields_ = [("p_probs", vp), ("p_req_stride", i64), ("p_row_stride", i64),
                ("q_probs", vp), ("q_req_stride", i64), ("q_row_stride", i64),
                ("draft_tok", vp), ("draft_stride", i64), ("u_acc", vp), ("u_acc_stride", i64), ("u_final", vp),
                ("B", C.c_int32), ("gamma", C.c_int32), ("V", i64), ("strict", C.c_int32),
                ("n_accepted", vp), ("next_tok", vp), ("ratios", vp), ("tie_count", vp),
                ("tokens", vp), ("tokens_stride", i64), ("seq_len", vp), ("active", vp),
                ("p_compact", vp), ("p_cmp_req_stride", i64), ("q_compact", vp), ("q_cmp_req_stride", i64),
                ("stats", vp)]


class Compact(C.Structure):
    """struct sd_compact (include/specdec_b200.h)"""
    _fields_ = [("cnt", vp), ("idx", vp), ("val", vp), ("cap", C.c_int32), ("row_stride", i64)]


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m llmspeculativesampling_b200.build` "
                "(there is no CPU or PyTorch fallback for the speculative-decoding kernels)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().sd_last_error().decode()
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")
