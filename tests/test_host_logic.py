"""CPU tests of the host side: C-ABI surface, tape, sharding (incl. world_size-2 gloo), API shape."""
import inspect
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    from llmspeculativesampling_b200 import build, _cabi
    build.build()
    lib = _cabi.load()
    hdr = open(os.path.join(ROOT, "include", "specdec_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sd_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_cabi.SIGNATURES), (declared ^ set(_cabi.SIGNATURES))
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sd_version() == 210


def test_no_cpu_fallback():
    from llmspeculativesampling_b200 import ops
    from llmspeculativesampling_b200.sampling import norm_logits, speculative_sampling
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.norm_probs(torch.randn(2, 100), 1.0, 5, 0.9)
    with pytest.raises(RuntimeError, match="CUDA"):
        norm_logits(torch.randn(1, 100), 1.0, 0, 0.0)
    with pytest.raises(RuntimeError, match="CUDA|no CPU path"):
        speculative_sampling(torch.randint(0, 10, (1, 4)), None, None, None, None, 8)


def test_product_never_imports_oracle():
    import llmspeculativesampling_b200
    pkg = os.path.dirname(llmspeculativesampling_b200.__file__)
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)


def test_tape_layout_matches_oracle_tape():
    from llmspeculativesampling_b200 import uniform_tape
    from oracle import tape
    for g in (1, 4, 7):
        assert uniform_tape.block(g) == tape.block(g) == 2 * g + 2
    assert uniform_tape.seed_of(3, 9) == tape.seed_of(3, 9)
    assert torch.equal(uniform_tape.make_tape(42, 5, 4), tape.make_tape(42, 5, 4))
    bt = uniform_tape.batch_tape(7, [0, 5, 2], 6, 4)
    assert bt.shape == (6, 3, 10)
    assert torch.equal(bt[:, 1], tape.make_tape(tape.seed_of(7, 5), 6, 4))
    u = bt.flatten() * 2 ** 24
    assert torch.equal(u, u.round()) and float(bt.max()) < 1.0


def test_shard_requests_partition():
    from llmspeculativesampling_b200.sharding import shard_requests, batches
    for n, w in [(256, 8), (10, 4), (3, 8)]:
        parts = [shard_requests(n, w, r) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert batches(list(range(10)), 4) == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]]
    with pytest.raises(ValueError):
        shard_requests(4, 2, 2)


def test_drop_in_signatures_follow_the_reference():
    from llmspeculativesampling_b200 import sampling
    sig = list(inspect.signature(sampling.speculative_sampling).parameters)
    assert sig[:13] == ["prefix", "approx_model", "target_model", "eos_token_id", "pad_token_id", "max_len", "gamma",
                        "temperature", "top_k", "top_p", "verbose", "random_seed", "details"]   # reference :1877-1881
    sig2 = list(inspect.signature(sampling.speculative_sampling_v2).parameters)
    assert sig2[:10] == ["prefix", "approx_model", "target_model", "max_len", "gamma", "temperature", "top_k", "top_p",
                         "random_seed", "details"]                                                # reference :2080-2082
    assert list(inspect.signature(sampling.norm_logits).parameters) == ["logits", "temperature", "top_k", "top_p"]
    assert list(inspect.signature(sampling.top_k_top_p_filter).parameters) == ["logits", "top_k", "top_p"]
    assert list(inspect.signature(sampling.max_fn).parameters) == ["x"]
    assert list(inspect.signature(sampling.sample).parameters)[:2] == ["probs", "num_samples"]
    kv = list(inspect.signature(sampling.KVCacheModel.__init__).parameters)
    assert kv[:5] == ["self", "model", "temperature", "top_k", "top_p"]                          # reference kvcache_model.py:24
    ar = list(inspect.signature(sampling.autoregressive_sampling).parameters)
    assert ar[:8] == ["x", "model", "N", "eos_token_id", "temperature", "top_k", "top_p", "pad_token_id"]
    for name in ["random_width_beam_sampling", "mjsd_speculative_sampling", "beam_speculative_sampling"]:
        with pytest.raises(NotImplementedError):
            getattr(sampling, name)()


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from llmspeculativesampling_b200.sharding import shard_requests, reduce_stats
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
ids = shard_requests(11, 2, rank)
st = reduce_stats({"requests": len(ids), "tokens": 10.0 * sum(ids), "elapsed_max": 1.0 + rank})
assert st == {"requests": 11.0, "tokens": 550.0, "elapsed_max": 2.0}, st
dist.destroy_process_group()
print("ok", rank)
"""


def test_sharded_stats_reduce_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_batching_queue_groups_concurrent_requests():
    """serving front-end: concurrent submits are served in batches, results go back to their callers,
    a failing batch fails its waiters only."""
    import threading
    from llmspeculativesampling_b200.serving import BatchingQueue
    calls = []

    def handler(reqs):
        calls.append(len(reqs))
        if any(r.get("boom") for r in reqs):
            raise ValueError("bad batch")
        return [r["x"] * 2 for r in reqs]

    q = BatchingQueue(handler, max_batch=4, max_wait_s=0.2)
    futs = [None] * 10

    def client(i):
        futs[i] = q.submit({"x": i})

    ts = [threading.Thread(target=client, args=(i,)) for i in range(10)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert [f.result(timeout=10) for f in futs] == [2 * i for i in range(10)]
    assert sum(calls) == 10 and max(calls) <= 4 and len(calls) <= 5
    bad = q.submit({"x": 1, "boom": True})
    with pytest.raises(ValueError):
        bad.result(timeout=10)
    assert q.submit({"x": 21}).result(timeout=10) == 42
    q.close()
    with pytest.raises(RuntimeError):
        q.submit({"x": 0})
