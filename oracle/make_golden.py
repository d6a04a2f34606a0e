"""Generate tests/golden/*.npz|json from the UNMODIFIED reference (dev container only).

    python -m oracle.make_golden

Fixtures (all produced by code under /root/reference, imported in place by oracle/ref_loader.py):
  norm_logits.npz   (logits, T, k, p) -> probs  by reference sampling/utils.py:norm_logits
  max_fn.npz        x -> max_fn(x)              by reference sampling/utils.py:max_fn
  multi_runs.json   end-to-end sampling.speculative_sampling.multi_speculative_sampling(strategy='iid')
  bild_runs.json    end-to-end sampling.speculative_sampling.BiLD_sampling on the replay models
  spec_runs.json    end-to-end sampling.speculative_sampling on the replay models with the
                    uniform tape: emitted token ids, acc_len per iteration, acc_rate
  v2_runs.json      end-to-end sampling.speculative_sampling.speculative_sampling_v2 (:2080-2194), same tape layout
  ar_runs.json      end-to-end sampling.autoregressive_sampling.autoregressive_sampling (:9-61), one uniform per token

    python -m oracle.make_golden [v2 ar ...]     only the named fixtures (default: all)
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import ref_loader, replay_model, tape

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

NORM_CASES = [  # (V, rows, T, top_k, top_p, scale, seed, dtype)
    (17, 3, 1.0, 0, 0.0, 2.0, 1, "f32"), (17, 3, 0.7, 5, 0.0, 2.0, 2, "f32"), (17, 3, 1.3, 0, 0.8, 2.0, 3, "f32"),
    (1000, 4, 0.8, 20, 0.9, 3.8, 4, "f32"), (1000, 4, 1.0, 1, 0.9, 3.0, 5, "f32"), (1000, 2, 1.0, 1000, 0.5, 3.0, 6, "f32"),
    (1000, 2, 0.5, 0, 0.9, 1.0, 7, "f32"), (1000, 2, 2.0, 50, 1.0, 3.0, 8, "f32"), (4099, 2, 0.8, 20, 0.9, 3.8, 9, "bf16"),
    (4099, 2, 1.0, 0, 0.0, 0.55, 10, "bf16"), (4099, 2, 1.0, 40, 0.95, 3.0, 11, "f16"),
    (32000, 2, 0.8, 20, 0.9, 3.8, 12, "f32"), (32000, 1, 1.0, 0, 0.0, 0.55, 13, "f32"),
    (50272, 1, 0.8, 20, 0.9, 3.8, 14, "bf16"), (32000, 1, 1.0, 0, 0.9, 2.0, 15, "f32"),
]


def _logits(V, rows, scale, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(rows, V, generator=g) * scale
    dt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[dtype]
    return x.to(dt)


BILD_CASES = [   # V, top_k, top_p, T, gamma, max_len, seed, noise, fallback_thres, rollback_thres
    (1000, 20, 0.9, 1.0, 4, 40, 31, 0.5, 0.6, 3.0), (32000, 20, 0.9, 0.8, 4, 32, 32, 0.5, 0.5, 2.0),
    (500, 0, 0.0, 1.0, 4, 40, 33, 0.5, 0.3, 2.0), (777, 0, 0.9, 1.3, 3, 32, 34, 0.3, 0.4, 4.0),
    (900, 5, 0.0, 0.7, 5, 40, 35, 0.8, 0.9, 1.0), (1000, 20, 0.9, 1.0, 8, 40, 36, 0.0, 0.2, 5.0),
]


def write_bild():
    """bild_runs.json: the reference's BiLD_sampling (speculative_sampling.py:1718-1873) on the replay models."""
    runs = []
    for (V, k, p, T, gamma, max_len, seed, noise, fb, rb) in BILD_CASES:
        d, t = replay_model.make_pair(V, seed=seed, noise=noise)
        prefix = torch.randint(3, V, (1, 7), generator=torch.Generator().manual_seed(seed))
        tp = tape.make_tape(seed, max_len + 1, gamma)
        out, det = ref_loader.run_reference_bild(prefix, d, t, max_len, gamma, fb, rb, T, k, p, tape=tp)
        runs.append(dict(V=V, top_k=k, top_p=p, temperature=T, gamma=gamma, max_len=max_len, seed=seed, noise=noise,
                         fallback_thres=fb, rollback_thres=rb, prefix=prefix[0].tolist(), tokens=out[0].tolist(),
                         acc_len=[int(a) for a in det["acc_len"]], target_call_times=int(det["target_call_times"]),
                         approx_call_times=int(det["approx_call_times"])))
        print("bild", V, k, p, T, gamma, fb, rb, "checks", det["target_call_times"], "draft tokens", det["approx_call_times"])
    with open(os.path.join(OUT, "bild_runs.json"), "w") as f:
        json.dump(runs, f)


BILD_EOS_CASES = [   # V, top_k, top_p, T, gamma, max_len, seed, noise, fallback_thres, rollback_thres, eos
    (48, 0, 0.0, 1.0, 4, 60, 71, 0.5, 0.02, 6.0, 5), (64, 20, 0.9, 1.0, 4, 60, 72, 0.5, 0.05, 5.0, 7), (40, 0, 0.0, 1.3, 6, 60, 73, 0.3, 0.02, 6.0, 3),
    (56, 10, 0.0, 1.0, 3, 60, 74, 0.8, 0.05, 4.0, 9),
]


def write_bild_eos():
    """bild_eos_runs.json: BiLD_sampling with an EOS id that the (small-vocabulary) draft emits now and then — the
    reference tests for EOS after every draft token (speculative_sampling.py:1826-1841), also between two checks."""
    runs = []
    for (V, k, p, T, gamma, max_len, seed, noise, fb, rb, eos) in BILD_EOS_CASES:
        d, t = replay_model.make_pair(V, seed=seed, noise=noise)
        prefix = torch.randint(10, V, (1, 7), generator=torch.Generator().manual_seed(seed))
        tp = tape.make_tape(seed, max_len + 1, gamma)
        out, det = ref_loader.run_reference_bild(prefix, d, t, max_len, gamma, fb, rb, T, k, p, tape=tp, eos_token_id=eos)
        runs.append(dict(V=V, top_k=k, top_p=p, temperature=T, gamma=gamma, max_len=max_len, seed=seed, noise=noise,
                         fallback_thres=fb, rollback_thres=rb, eos=eos, prefix=prefix[0].tolist(), tokens=out[0].tolist(),
                         acc_len=[int(a) for a in det["acc_len"]], target_call_times=int(det["target_call_times"]),
                         approx_call_times=int(det["approx_call_times"])))
        print("bild-eos", V, gamma, "generated", out.shape[1] - 7, "of", max_len, "checks", det["target_call_times"], "drafted", det["approx_call_times"])
    with open(os.path.join(OUT, "bild_eos_runs.json"), "w") as f:
        json.dump(runs, f)


MULTI_CASES = [   # V, top_k, top_p, T, gamma, width, max_len, seed, noise
    (1000, 20, 0.9, 1.0, 4, 3, 32, 41, 0.5), (32000, 20, 0.9, 0.8, 4, 4, 24, 42, 0.5), (500, 0, 0.0, 1.0, 4, 4, 32, 43, 0.8),
    (777, 0, 0.9, 1.3, 3, 2, 32, 44, 0.3), (900, 5, 0.0, 0.7, 5, 3, 32, 45, 1.5), (1000, 20, 0.9, 1.0, 2, 8, 24, 46, 2.0),
]


def multi_tape(seed: int, iterations: int, gamma: int, width: int) -> torch.Tensor:
    return torch.rand(iterations, ref_loader.multi_block(gamma, width), generator=torch.Generator().manual_seed(seed))


def write_multi():
    """multi_runs.json: the reference's multi_speculative_sampling(strategy='iid') (speculative_sampling.py:1379-1716)."""
    runs = []
    for (V, k, p, T, gamma, width, max_len, seed, noise) in MULTI_CASES:
        d, t = replay_model.make_pair(V, seed=seed, noise=noise)
        prefix = torch.randint(3, V, (1, 7), generator=torch.Generator().manual_seed(seed))
        tp = multi_tape(seed, max_len + 1, gamma, width)
        out, det = ref_loader.run_reference_multi(prefix, d, t, max_len, gamma, width, T, k, p, tp)
        runs.append(dict(V=V, top_k=k, top_p=p, temperature=T, gamma=gamma, width=width, max_len=max_len, seed=seed,
                         noise=noise, prefix=prefix[0].tolist(), tokens=out[0].tolist(),
                         acc_len=[int(a) for a in det["acc_len"]], acc_rate=float(det["acc_rate"])))
        print("multi", V, k, p, T, gamma, width, "mean acc len", np.mean(det["acc_len"]))
    with open(os.path.join(OUT, "multi_runs.json"), "w") as f:
        json.dump(runs, f)


V2_CASES = [   # V, top_k, top_p, T, gamma, max_len, seed, noise
    (1000, 20, 0.9, 0.8, 4, 40, 51, 0.5), (32000, 20, 0.9, 1.0, 4, 20, 52, 0.5), (500, 0, 0.0, 1.0, 4, 32, 53, 0.5),
    (777, 0, 0.9, 1.3, 3, 32, 54, 0.3), (900, 5, 0.0, 0.7, 5, 32, 55, 0.8), (2048, 20, 0.9, 1.0, 1, 20, 56, 0.5),
    (1000, 20, 0.9, 1.0, 8, 40, 57, 0.0), (1500, 10, 0.5, 1.0, 4, 32, 58, 2.0),
]


def write_v2():
    """v2_runs.json: the reference's speculative_sampling_v2 (speculative_sampling.py:2080-2194) on the replay models."""
    runs = []
    for (V, k, p, T, gamma, max_len, seed, noise) in V2_CASES:
        d, t = replay_model.make_pair(V, seed=seed, noise=noise)
        prefix = torch.randint(3, V, (1, 7), generator=torch.Generator().manual_seed(seed))
        tp = tape.make_tape(seed, max_len + 1, gamma)
        out, det = ref_loader.run_reference_v2(prefix, d, t, max_len, gamma, T, k, p, tape=tp)
        runs.append(dict(V=V, top_k=k, top_p=p, temperature=T, gamma=gamma, max_len=max_len, seed=seed, noise=noise,
                         prefix=prefix[0].tolist(), tokens=out[0].tolist(), acc_len=[int(a) for a in det["acc_len"]],
                         acc_rate=float(det["acc_rate"])))
        print("v2", V, k, p, T, gamma, "mean acc", np.mean(det["acc_len"]))
    with open(os.path.join(OUT, "v2_runs.json"), "w") as f:
        json.dump(runs, f)


AR_CASES = [   # V, top_k, top_p, T, N, seed, eos (None: never stops early)
    (1000, 20, 0.9, 0.8, 40, 61, None), (32000, 20, 0.9, 1.0, 24, 62, None), (500, 0, 0.0, 1.0, 40, 63, None),
    (777, 0, 0.9, 1.3, 32, 64, None), (900, 5, 0.0, 0.7, 32, 65, None), (300, 1, 0.0, 1.0, 16, 66, None),
    (64, 0, 0.0, 1.0, 200, 67, 5), (50272, 20, 0.9, 0.8, 12, 68, None),
]


def ar_uniforms(seed: int, N: int) -> torch.Tensor:
    return torch.rand(N, generator=torch.Generator().manual_seed(int(seed)))


def write_ar():
    """ar_runs.json: the reference's autoregressive_sampling (autoregressive_sampling.py:9-61) on the replay target."""
    runs = []
    for (V, k, p, T, N, seed, eos) in AR_CASES:
        _, t = replay_model.make_pair(V, seed=seed, noise=0.5)
        prefix = torch.randint(3, V, (1, 7), generator=torch.Generator().manual_seed(seed))
        out = ref_loader.run_reference_ar(prefix, t, N, T, k, p, ar_uniforms(seed, N), eos_token_id=-1 if eos is None else eos)
        runs.append(dict(V=V, top_k=k, top_p=p, temperature=T, N=N, seed=seed, eos=eos, prefix=prefix[0].tolist(),
                         tokens=out[0].tolist()))
        print("ar", V, k, p, T, N, "generated", out.shape[1] - 7)
    with open(os.path.join(OUT, "ar_runs.json"), "w") as f:
        json.dump(runs, f)


def main():
    import sys
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    only = set(sys.argv[1:])
    if only:
        for name, fn in (("v2", write_v2), ("ar", write_ar), ("bild", write_bild), ("multi", write_multi), ("bild_eos", write_bild_eos)):
            if name in only:
                fn()
        return
    utils = ref_loader.load_utils()
    blob = {}
    for ci, (V, rows, T, k, p, scale, seed, dtype) in enumerate(NORM_CASES):
        x = _logits(V, rows, scale, seed, dtype)
        probs = torch.cat([utils.norm_logits(x[i:i + 1].float(), T, k, p) for i in range(rows)], 0)
        # fixtures keep only the non-zero entries of the big rows (probs are sparse after top-k)
        nz = probs.nonzero()
        blob[f"c{ci}_meta"] = np.array([V, rows, T, k, p, scale, seed, {"f32": 0, "bf16": 1, "f16": 2}[dtype]], dtype=np.float64)
        blob[f"c{ci}_nz_idx"] = nz.numpy().astype(np.int32)
        blob[f"c{ci}_nz_val"] = probs[nz[:, 0], nz[:, 1]].numpy()
    np.savez_compressed(os.path.join(OUT, "norm_logits.npz"), **blob)

    g = torch.Generator().manual_seed(99)
    xs = torch.randn(6, 2000, generator=g) * 0.01
    xs[5] = -xs[5].abs()                                        # all-negative row -> all zeros
    np.savez_compressed(os.path.join(OUT, "max_fn.npz"), x=xs.numpy(), y=utils.max_fn(xs).numpy())

    runs = []
    for (V, k, p, T, gamma, max_len, seed, noise) in [
        (1000, 20, 0.9, 0.8, 4, 48, 11, 0.5), (32000, 20, 0.9, 1.0, 4, 40, 12, 0.5), (500, 0, 0.0, 1.0, 4, 40, 13, 0.5),
        (777, 0, 0.9, 1.3, 3, 40, 14, 0.3), (900, 5, 0.0, 0.7, 5, 40, 15, 0.8), (2048, 20, 0.9, 1.0, 1, 24, 16, 0.5),
        (1000, 20, 0.9, 1.0, 8, 48, 17, 0.0), (1500, 10, 0.5, 1.0, 4, 40, 18, 2.0),
    ]:
        d, t = replay_model.make_pair(V, seed=seed, noise=noise)
        prefix = torch.randint(3, V, (1, 7), generator=torch.Generator().manual_seed(seed))
        tp = tape.make_tape(seed, max_len + 1, gamma)
        out, det = ref_loader.run_reference(prefix, d, t, max_len, gamma, T, k, p, tape=tp)
        runs.append(dict(V=V, top_k=k, top_p=p, temperature=T, gamma=gamma, max_len=max_len, seed=seed, noise=noise,
                         prefix=prefix[0].tolist(), tokens=out[0].tolist(), acc_len=[int(a) for a in det["acc_len"]],
                         acc_rate=float(det["acc_rate"])))
        print("run", V, k, p, T, gamma, "mean acc", np.mean(det["acc_len"]))
    with open(os.path.join(OUT, "spec_runs.json"), "w") as f:
        json.dump(runs, f)
    write_bild()
    write_multi()
    write_v2()
    write_ar()
    write_bild_eos()
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
