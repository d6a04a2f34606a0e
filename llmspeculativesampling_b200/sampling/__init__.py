"""Drop-in surface of the reference's `sampling` package (/root/reference/sampling/__init__.py:1-7)
for the draft-and-verify hot path, plus `multi_speculative_sampling(strategy='iid')` and `BiLD_sampling`
(SURVEY.md §8f rows N2, N3).  The beam / tree research variants of the reference are out of scope (SURVEY.md §2 rows
7-11) and raise NotImplementedError if called."""
from .speculative_sampling import speculative_sampling, speculative_sampling_v2
from .autoregressive_sampling import autoregressive_sampling
from .kvcache_model import KVCacheModel
from .bild import BiLD_sampling
from .multi import multi_speculative_sampling
from .utils import norm_logits, top_k_top_p_filter, sample, max_fn


def _out_of_scope(name):
    def f(*a, **k):
        raise NotImplementedError(f"{name} is outside the draft-and-verify hot path this package accelerates")
    f.__name__ = name
    return f


beam_speculative_sampling = _out_of_scope("beam_speculative_sampling")
beam_speculative_sampling_v2 = _out_of_scope("beam_speculative_sampling_v2")
mjsd_speculative_sampling = _out_of_scope("mjsd_speculative_sampling")
random_width_beam_sampling = _out_of_scope("random_width_beam_sampling")

__all__ = ["speculative_sampling", "speculative_sampling_v2", "autoregressive_sampling", "multi_speculative_sampling",
           "beam_speculative_sampling", "BiLD_sampling", "mjsd_speculative_sampling", "random_width_beam_sampling",
           "beam_speculative_sampling_v2", "KVCacheModel", "norm_logits", "top_k_top_p_filter", "sample", "max_fn"]
