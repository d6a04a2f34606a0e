"""GPU parity: kernel 1 (fused temperature/top-k/top-p/softmax [+sample]) vs the CPU oracle and the
golden vectors produced by the unmodified reference.  All calls go through the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_ops
from tests.helpers import compare_probs, make_logits, oracle_probs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

CASES = [  # V, rows, T, k, p, scale, dtype
    (32000, 6, 0.8, 20, 0.9, 3.8, torch.float32),
    (32000, 4, 1.0, 20, 0.9, 3.8, torch.bfloat16),
    (50272, 4, 0.8, 20, 0.9, 3.8, torch.bfloat16),
    (50272, 3, 1.0, 0, 0.0, 0.55, torch.float32),
    (32000, 3, 1.0, 0, 0.0, 0.55, torch.float16),
    (32000, 3, 0.7, 0, 0.9, 2.0, torch.float32),
    (32000, 2, 1.0, 1, 0.0, 3.0, torch.float32),
    (32000, 2, 1.3, 128, 0.95, 3.0, torch.float32),
    (32000, 2, 1.0, 500, 0.0, 3.0, torch.float32),
    (32000, 2, 1.0, 500, 0.9, 3.0, torch.bfloat16),
    (1000, 5, 0.8, 20, 0.9, 3.8, torch.float32),
    (17, 4, 1.0, 5, 0.8, 2.0, torch.float32),
    (50257, 3, 0.8, 20, 0.9, 3.8, torch.float32),      # odd V: no TMA path
    (50257, 2, 1.0, 0, 0.0, 1.0, torch.bfloat16),
    (128256, 2, 0.8, 20, 0.9, 3.8, torch.bfloat16),
    (262144, 2, 0.8, 20, 0.9, 3.8, torch.float32),
    (262144, 1, 1.0, 0, 0.0, 1.0, torch.float32),
    (151936, 2, 1.0, 50, 0.8, 3.0, torch.float32),
    # dense rows (top_k = 0, top_p = 0) on the ring kernel: temperature != 1, 16-bit logits, rows shorter than one chunk
    (32000, 3, 0.8, 0, 0.0, 3.8, torch.float32),
    (32000, 5, 1.3, 0, 0.0, 2.0, torch.bfloat16),
    (50272, 3, 0.7, 0, 0.0, 3.0, torch.bfloat16),
    (48, 3, 1.0, 0, 0.0, 2.0, torch.float32),
    (40, 1, 1.3, 0, 0.0, 2.0, torch.float32),
    (4096, 300, 0.9, 0, 0.0, 3.0, torch.float32),      # more rows than SMs: several rows per persistent CTA
    (50272, 160, 0.8, 0, 0.0, 2.0, torch.float32),     # 13 chunks = every ring slot: pass A of the next row lags pass B
    (53248, 3, 1.0, 0, 0.0, 1.0, torch.float32),
]


@pytest.mark.parametrize("V,rows,T,k,p,scale,dtype", CASES)
@pytest.mark.parametrize("path", ["pipeline", "cluster", "classic", "general"])
def test_norm_probs_matches_oracle(cuda_lib, V, rows, T, k, p, scale, dtype, path):
    """pipeline: the library's choice — the ring kernel (one persistent CTA per SM) where a row fits one CTA and the
    setting is top-k or dense, else the persistent cluster pipeline, else the one-cluster-per-row kernel; cluster: the
    same without the ring kernel; classic: one-cluster-per-row kernel; general: its sort-free threshold search."""
    from llmspeculativesampling_b200 import ops
    x = make_logits(rows, V, scale, seed=V + rows + k, dtype=dtype)
    want = oracle_probs(x, T, k, p)
    got = ops.norm_probs(x.cuda(), T, k, p, general=(path == "general"), pipeline={"pipeline": True, "cluster": "cluster"}.get(path, False))
    ops.default_flag("cuda").check()
    boundary = compare_probs(got, want, f"V={V} k={k} p={p} path={path}")
    assert boundary == 0, f"{boundary} rows with a different support (boundary ties) — none expected on these seeds"
    assert torch.allclose(got.sum(-1).cpu(), torch.ones(rows), atol=1e-5)


@pytest.mark.parametrize("cluster,threads", [(1, 256), (2, 256), (4, 256), (8, 256), (2, 512), (1, 1024), (4, 512)])
def test_norm_probs_all_launch_shapes(cuda_lib, cluster, threads):
    from llmspeculativesampling_b200 import ops
    x = make_logits(5, 32000, 3.8, seed=5, dtype=torch.float32)
    try:
        ops.set_tuning(cluster, threads, 0)
        for (T, k, p) in [(0.8, 20, 0.9), (1.0, 0, 0.0), (1.0, 0, 0.9), (1.0, 300, 0.5)]:
            if cluster == 1 and threads == 256:
                continue    # 128 KB slice + scratch does not fit three CTAs; covered by (1, 1024)
            want = oracle_probs(x, T, k, p)
            got = ops.norm_probs(x.cuda(), T, k, p)
            ops.default_flag("cuda").check()
            assert compare_probs(got, want, f"C={cluster} T={threads} k={k} p={p}") == 0
    finally:
        ops.set_tuning(0, 0, 0)


def test_norm_probs_golden_reference_vectors(cuda_lib):
    """Fixtures written by oracle/make_golden.py from the UNMODIFIED reference norm_logits."""
    from llmspeculativesampling_b200 import ops
    from oracle.make_golden import NORM_CASES, _logits
    blob = np.load(os.path.join(GOLD, "norm_logits.npz"))
    for ci, (V, rows, T, k, p, scale, seed, dtype) in enumerate(NORM_CASES):
        x = _logits(V, rows, scale, seed, dtype)
        want = torch.zeros(rows, V)
        nz = torch.from_numpy(blob[f"c{ci}_nz_idx"].astype(np.int64))
        want[nz[:, 0], nz[:, 1]] = torch.from_numpy(blob[f"c{ci}_nz_val"])
        got = ops.norm_probs(x.cuda(), T, k, p)
        ops.default_flag("cuda").check()
        assert compare_probs(got, want, f"golden case {ci}") == 0


def test_massive_ties_fall_back_to_general_path(cuda_lib):
    from llmspeculativesampling_b200 import ops
    V = 32000
    x = torch.zeros(4, V)
    x[1] = torch.randint(0, 3, (V,)).float()                 # three distinct values -> ~10k-way ties at the top
    x[2, ::2] = 1.0
    x[3] = make_logits(1, V, 3.0, 3).bfloat16().float().round()   # coarse grid
    for (T, k, p) in [(1.0, 20, 0.9), (0.8, 20, 0.0), (1.0, 5, 0.5)]:
        want = oracle_probs(x, T, k, p)
        for pipeline in (True, False):
            got = ops.norm_probs(x.cuda(), T, k, p, pipeline=pipeline)
            ops.default_flag("cuda").check()
            assert compare_probs(got, want, f"ties k={k} p={p} pipeline={pipeline}") == 0


def test_minus_inf_and_strided_rows(cuda_lib):
    from llmspeculativesampling_b200 import ops
    V = 4096
    big = make_logits(6, 2 * V, 3.0, 11)
    x = big[:, :V]                                           # row stride 2V (non-contiguous rows)
    x[:, 100:2000] = float("-inf")
    for (T, k, p) in [(1.0, 20, 0.9), (1.0, 0, 0.0), (0.9, 0, 0.7)]:
        want = oracle_probs(x.contiguous(), T, k, p)
        got = ops.norm_probs(big.cuda()[:, :V], T, k, p)
        ops.default_flag("cuda").check()
        assert compare_probs(got, want, f"-inf k={k} p={p}") == 0
        assert float(got[:, 100:2000].abs().sum()) == 0.0


def test_dense_16bit_rows_with_output_rows_not_32_byte_aligned(cuda_lib):
    """16-bit logits on the dense ring kernel write 8 floats per thread with one 32-byte store when the output row is
    32-byte aligned and with two 16-byte stores otherwise: both give the same rows."""
    from llmspeculativesampling_b200 import ops
    V, rows = 4096, 7
    x = make_logits(rows, V, 2.0, 3, dtype=torch.bfloat16).cuda()
    want = ops.norm_probs(x, 0.9, 0, 0.0)
    wide = torch.full((rows, V + 4), -1.0, device="cuda")        # row stride 4100 floats: rows 1, 3, 5 are only 16-byte aligned
    got = ops.norm_probs(x, 0.9, 0, 0.0, out=wide[:, :V])
    ops.default_flag("cuda").check()
    assert torch.equal(got, want) and bool((wide[:, V:] == -1.0).all())
    assert compare_probs(got, oracle_probs(x.cpu(), 0.9, 0, 0.0), "bf16 dense, strided output") == 0


def test_nan_logit_raises_like_reference(cuda_lib):
    from llmspeculativesampling_b200 import ops
    x = make_logits(2, 1000, 1.0, 1)
    x[1, 7] = float("nan")
    ops.norm_probs(x.cuda(), 1.0, 20, 0.9)
    with pytest.raises(RuntimeError, match="norm logits error"):
        ops.default_flag("cuda").check()
    ops.norm_probs(x.cuda(), 1.0, 0, 0.0)
    with pytest.raises(RuntimeError, match="norm logits error"):
        ops.default_flag("cuda").check()


@pytest.mark.parametrize("V,T,k,p,dtype", [(32000, 0.8, 20, 0.9, torch.float32), (32000, 1.0, 0, 0.0, torch.float32),
                                           (50272, 1.0, 0, 0.9, torch.bfloat16), (1000, 1.0, 300, 0.0, torch.float32),
                                           (50257, 1.0, 20, 0.9, torch.float16)])
def test_norm_sample_tokens_bit_exact(cuda_lib, V, T, k, p, dtype):
    """Sampled token must equal the oracle's inverse-CDF rule applied to the SAME probabilities."""
    from llmspeculativesampling_b200 import ops
    rows = 48
    x = make_logits(rows, V, 3.0 if k else 1.0, seed=V + k, dtype=dtype)
    u = torch.rand(rows, generator=torch.Generator().manual_seed(3))
    probs = torch.empty(rows, V, device="cuda")
    tok = ops.norm_sample(x.cuda(), T, k, p, u.cuda(), probs_out=probs)
    ops.default_flag("cuda").check()
    pc = probs.cpu()
    want = [ref_ops.icdf_sample(pc[i], float(u[i])) for i in range(rows)]
    assert tok.cpu().tolist() == want
    # and the token-only variant (no dense write) gives the same ids
    tok2 = ops.norm_sample(x.cuda(), T, k, p, u.cuda(), probs_out=None)
    assert torch.equal(tok, tok2)
    tok3 = ops.norm_sample(x.cuda(), T, k, p, u.cuda(), probs_out=None, pipeline=False)
    assert torch.equal(tok, tok3)
    # against the oracle's own probabilities the only permitted divergence is a boundary tie
    wp = oracle_probs(x, T, k, p)
    diff = 0
    for i in range(rows):
        t, margin = ref_ops.icdf_sample(wp[i], float(u[i]), return_margin=True)
        if t != want[i]:
            # CDF positions inherit the 1e-5 relative tolerance of the probabilities themselves
            assert margin < 1e-5, f"row {i}: token differs with margin {margin}"
            diff += 1
    assert diff <= 1


@pytest.mark.parametrize("V,dtype,rows", [(32000, torch.float32, 576), (50272, torch.bfloat16, 300), (32000, torch.bfloat16, 149),
                                          (128256, torch.bfloat16, 40), (262144, torch.float32, 9)])
@pytest.mark.parametrize("kernel", [True, "cluster"])
def test_pipeline_kernel_many_rows_with_sampling(cuda_lib, V, dtype, rows, kernel):
    """More rows than SMs: every persistent CTA walks several work items through all buffers / parities (kernel True:
    the ring kernel where a row fits one CTA; 'cluster': the persistent cluster pipeline)."""
    from llmspeculativesampling_b200 import ops
    x = make_logits(rows, V, 3.8, seed=rows, dtype=dtype).cuda()
    x[3] = 1.0                                               # an all-ties row must be flagged and served by the general path
    u = torch.rand(rows, generator=torch.Generator().manual_seed(1)).cuda()
    pa = torch.empty(rows, V, device="cuda"); pb = torch.empty(rows, V, device="cuda")
    ta = ops.norm_sample(x, 0.8, 20, 0.9, u, probs_out=pa, pipeline=kernel)
    tb = ops.norm_sample(x, 0.8, 20, 0.9, u, probs_out=pb, pipeline=False)
    ops.default_flag("cuda").check()
    assert torch.equal(pa, pb), "pipelined and classic kernels must agree bit for bit"
    assert torch.equal(ta, tb)
    # u < 0 switches sampling off for that row (its tok_out entry is left untouched)
    u2 = u.clone(); u2[1::2] = -1.0
    for pipeline in (kernel, False):
        tc = torch.full((rows,), -5, dtype=torch.int64, device="cuda")
        ops.norm_sample(x, 0.8, 20, 0.9, u2, probs_out=pb, tok_out=tc, pipeline=pipeline)
        assert torch.equal(tc[0::2], ta[0::2]) and bool((tc[1::2] == -5).all())
        assert torch.equal(pa, pb)
    sel = [0, 3, rows // 2, rows - 1]
    want = oracle_probs(x[sel].cpu(), 0.8, 20, 0.9)
    assert compare_probs(pa[sel], want, "pipeline vs oracle") == 0


@pytest.mark.parametrize("V,rows,n_tied", [(4096, 9000, 3), (32000, 2400, 40), (2048, 40000, 40000), (4096, 30000, 15000),
                                           # top-k rows longer than the ring (pass 2 from L2): tied rows go to the follow-up launch
                                           (131072, 400, 7), (131072, 300, 300), (262144, 150, 1)])
@pytest.mark.parametrize("kernel", [True, "cluster"])
def test_pipeline_kernel_dynamic_rows_and_deferred_rows(cuda_lib, V, rows, n_tied, kernel):
    """Rows are handed to the clusters through a ticket counter (tens of items per cluster: the row-index ring wraps), and
    rows the candidate path cannot serve (all ties) are deferred to the general path in bounded batches — with tens of
    thousands of them every cluster pauses, drains its list and resumes several times.  Two launches back to back check
    that the counter is re-armed."""
    from llmspeculativesampling_b200 import ops
    x = make_logits(rows, V, 3.8, seed=rows, dtype=torch.float32).cuda()
    g = torch.Generator().manual_seed(3)
    tied = torch.randperm(rows, generator=g)[:n_tied].cuda()
    x[tied] = 0.25
    u = torch.rand(rows, generator=g).cuda()
    pa = torch.empty(rows, V, device="cuda"); pb = torch.empty(rows, V, device="cuda")
    tb = ops.norm_sample(x, 0.8, 20, 0.9, u, probs_out=pb, pipeline=False)
    for _ in range(2):
        pa.fill_(-1.0)
        ta = ops.norm_sample(x, 0.8, 20, 0.9, u, probs_out=pa, pipeline=kernel)
        ops.default_flag("cuda").check()
        assert torch.equal(pa, pb), "pipelined and classic kernels must agree bit for bit"
        assert torch.equal(ta, tb)
    sel = [0, int(tied[0]), rows // 2, rows - 1]
    want = oracle_probs(x[sel].cpu(), 0.8, 20, 0.9)
    assert compare_probs(pa[sel], want, "pipeline vs oracle") == 0


@pytest.mark.parametrize("V,dtype,rows,T", [(32000, torch.float32, 576, 1.0), (50272, torch.bfloat16, 300, 0.8), (32000, torch.float16, 149, 1.3),
                                            (50272, torch.float32, 160, 1.0), (4096, torch.float32, 5000, 0.7), (1000, torch.bfloat16, 40, 1.0),
                                            # rows longer than the ring: streamed through it twice
                                            (65536, torch.float32, 200, 0.9), (131072, torch.bfloat16, 160, 1.0),
                                            (151936, torch.float32, 20, 1.0), (262144, torch.float32, 5, 0.8)])
def test_ring_kernel_dense_rows_with_sampling(cuda_lib, V, dtype, rows, T):
    """Dense rows (top_k = 0, top_p = 0 — the reference API's default, speculative_sampling.py:1879-1880) on the ring
    kernel: probabilities against the oracle, the sampled token bit-exact against the inverse-CDF rule applied to the
    SAME probabilities (the sampler's two-limb weight sums must be exact), rows with u < 0 left unsampled, a row that
    contains -inf entries, and the one-cluster-per-row kernel as a cross-check of the probabilities."""
    from llmspeculativesampling_b200 import ops
    scale = 0.55 if rows % 2 else 3.0                        # flat (random-init-like) and peaked distributions
    x = make_logits(rows, V, scale, seed=rows + V, dtype=dtype).cuda()
    x[1, ::3] = float("-inf")
    u = torch.rand(rows, generator=torch.Generator().manual_seed(2)).cuda()
    u[2] = 0.0
    u[3] = 1.0 - 2.0 ** -24
    pa = torch.empty(rows, V, device="cuda"); pb = torch.empty(rows, V, device="cuda")
    for rep in range(2):                                     # twice: the ticket counter must be re-armed
        pa.fill_(-1.0)
        ta = ops.norm_sample(x, T, 0, 0.0, u, probs_out=pa)
        ops.default_flag("cuda").check()
    ops.norm_probs(x, T, 0, 0.0, out=pb, pipeline=False)
    assert compare_probs(pa, pb, "ring dense vs one-cluster-per-row kernel") == 0
    pc = pa.cpu()
    assert torch.allclose(pc.sum(-1), torch.ones(rows), atol=2e-5)
    sel = sorted(set([0, 1, 2, 3, rows // 2, rows - 1]))
    assert compare_probs(pa[sel], oracle_probs(x[sel].cpu(), T, 0, 0.0), "ring dense vs oracle") == 0
    check = range(rows) if rows <= 600 else list(range(0, rows, 37)) + [rows - 1]
    for i in check:
        assert int(ta[i]) == ref_ops.icdf_sample(pc[i], float(u[i])), f"row {i}"
    # rows with u < 0 are not sampled; probabilities only (no uniform) give the same rows
    u2 = u.clone(); u2[1::2] = -1.0
    tc = torch.full((rows,), -5, dtype=torch.int64, device="cuda")
    ops.norm_sample(x, T, 0, 0.0, u2, probs_out=pb, tok_out=tc)
    assert torch.equal(tc[0::2], ta[0::2]) and bool((tc[1::2] == -5).all()) and torch.equal(pa, pb)
    assert torch.equal(ops.norm_probs(x, T, 0, 0.0), pa)


def test_dense_rows_against_float64_softmax(cuda_lib):
    """Accuracy of the dense kernels in absolute terms: on wide rows the reference's own fp32 log-softmax carries up to
    ~1e-5 of rounding bias (its probabilities sum to 1.00001 on some rows of this case), so a 1e-5 comparison with the
    oracle measures the oracle there; against the float64 softmax of the same fp32 quotients the kernels stay within 3e-6
    on every entry that matters, on the ring kernel (incl. rows that share a CTA) and on the one-cluster-per-row kernel."""
    from llmspeculativesampling_b200 import ops
    V, rows, T = 50272, 160, 0.8
    x = make_logits(rows, V, 3.0, seed=V + rows, dtype=torch.float32)
    p64 = torch.softmax((x / T).double(), dim=-1)
    big = p64 > 1e-6
    for pipeline in (True, False):
        got = ops.norm_probs(x.cuda(), T, 0, 0.0, pipeline=pipeline).cpu().double()
        ops.default_flag("cuda").check()
        rel = ((got - p64).abs() / p64)[big]
        assert float(rel.max()) < 3e-6, f"pipeline={pipeline}: max rel {float(rel.max()):.3e}"
        assert float((got - p64).abs()[~big].max()) < 1e-11
        assert torch.allclose(got.sum(-1), torch.ones(rows, dtype=torch.float64), atol=2e-6)


def test_fuzz_shapes_parameters_and_ties_against_oracle(cuda_lib):
    """Randomised sweep (fixed seed) over V, rows, dtype, T, top_k, top_p, tie density, -inf masks and row strides, on
    every kernel path; probabilities must match the oracle and the support must be identical."""
    from llmspeculativesampling_b200 import ops
    rng = np.random.default_rng(1234)
    n_cases, boundary_rows, total_rows = 60, 0, 0
    for case in range(n_cases):
        V = int(rng.choice([17, 64, 257, 1000, 1024, 4099, 8192, 12345, 32000]))
        rows = int(rng.integers(1, 7))
        dtype = [torch.float32, torch.bfloat16, torch.float16][int(rng.integers(0, 3))]
        T = float(rng.choice([1.0, 0.7, 0.8, 1.3, 2.0]))
        k = int(rng.choice([0, 1, 2, 5, 20, 50, 128, 129, 400]))
        p = float(rng.choice([0.0, 0.3, 0.9, 0.95, 1.0]))
        scale = float(rng.choice([0.3, 1.0, 3.8, 8.0]))
        g = torch.Generator().manual_seed(case)
        x = torch.randn(rows, V, generator=g) * scale
        mode = int(rng.integers(0, 4))
        if mode == 1:
            x = (x * 2).round() / 2                               # dense ties
        elif mode == 2:
            x[:, rng.integers(0, V, size=max(1, V // 3))] = float("-inf")
        x = x.to(dtype)
        pad = int(rng.choice([0, 0, 8, 24]))
        big = torch.zeros(rows, V + pad, dtype=dtype)
        big[:, :V] = x
        xd = big.cuda()[:, :V]                                    # row stride V + pad
        want = oracle_probs(x, T, k, p)
        for path in ("pipeline", "cluster", "classic", "general"):
            got = ops.norm_probs(xd, T, k, p, general=(path == "general"), pipeline={"pipeline": True, "cluster": "cluster"}.get(path, False))
            ops.default_flag("cuda").check()
            boundary_rows += compare_probs(got, want, f"case {case} V={V} rows={rows} {dtype} T={T} k={k} p={p} mode={mode} path={path}")
            total_rows += rows
    # a different support is only legitimate when the top-p cumulative sum lands within fp32 rounding of top_p
    assert boundary_rows <= max(2, total_rows // 100), f"{boundary_rows} boundary rows out of {total_rows}"
