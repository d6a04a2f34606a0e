"""llmspeculativesampling_b200 — B200-native (sm_100a) speculative-decoding draft-and-verify hot path.

Drop-in for the `sampling` package of ZongyueQin/LLMSpeculativeSampling:

    from llmspeculativesampling_b200.sampling import speculative_sampling, speculative_sampling_v2
    from llmspeculativesampling_b200.sampling.kvcache_model import KVCacheModel
    from llmspeculativesampling_b200.sampling.utils import norm_logits, top_k_top_p_filter, sample, max_fn

Host code is Python/PyTorch (device memory, streams, HF model forwards); everything between the
models' logits and the next step's input runs in hand-written CUDA kernels reached through the
C ABI of include/specdec_b200.h.  There is no CPU fallback.
"""
__version__ = "0.1.0"
