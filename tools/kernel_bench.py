"""Micro-benchmarks of kernels 2 and 3 and the small utilities: achieved HBM GB/s of
every launch that replaces /root/reference/sampling/speculative_sampling.py:1966-2027 (verify), utils.py:213-245 (sample,
max_fn) and kvcache_model.py:360-436 (rollback / append), with the algorithmic bytes computed from the ACTUAL outcomes.

    python tools/kernel_bench.py [--mode all|verify_dense|verify_sparse|verify_multi|verify_bild|sample|max_fn|kv_append|
                                  kv_select|build_step] [--V 32000] [--B 64] [--gamma 4] [--iters 200]

Method: the launch is captured once per rotating input set into ONE CUDA graph (sets total > 2x the 126 MB L2 wherever
the kernel's working set allows), the graph is replayed back to back and timed with CUDA events on the launching stream
after warm-up; ms = per launch.  One JSON line per point.  `algorithmic_bytes` follows SURVEY.md §8(d):
  verify      per request 2*gamma*4 B of gathers + 2*V*4 B if a rejection occurred (rows p_n, q_n) or V*4 B if all accepted
              (row p_last) + 16 B out — from the accept counts the run produced
  sample      V*4 B per row + 8 B        max_fn   2*V*4 B per row
  kv_append   K and V of every appended token read + written: 2 * 2*H*D*s B per (request, new row) and layer
  kv_select   per request count*H*D*s*2 read + (W-1)*count*H*D*s*2 written
The sparse verify reads only the compact lists ((2*gamma+1) * (4 + 8*cnt) B per request): it is latency-bound by design
and its line carries both figures (`algorithmic_bytes` of the dense definition and `bytes_touched`).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llmspeculativesampling_b200 import build, ops  # noqa: E402

L2 = 126e6
ONCE = False


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def graph_time(launches, iters):
    """launches: list of callables (one per rotating input set).  Returns ms per launch."""
    for f in launches:
        f()
    torch.cuda.synchronize()
    if ONCE:                                   # under ncu: one eager launch per input set is all the profiler needs
        return float("nan")
    side = torch.cuda.Stream()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(gr, stream=side):
            for f in launches:
                f()
    torch.cuda.synchronize()
    reps = max(3, iters // len(launches))
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * len(launches))


def emit(kernel, ms, alg_bytes, extra):
    peak, src = peak_gbs()
    gbs = alg_bytes / ms / 1e6
    d = dict(kernel=kernel, ms=round(ms, 5), algorithmic_bytes=int(alg_bytes), GBs=round(gbs, 1), frac_of_measured=round(gbs / peak, 4),
             frac_of_8TBs=round(gbs / 8000.0, 4), peak_source=src)
    d.update(extra)
    print(json.dumps(d), flush=True)
    return d


def synth_probs(B, gamma, V, T, k, p, seed, dtype=torch.float32, width=1, noise=0.5):
    """Config-2 style logits (shared base + noise) -> q rows, drafted tokens, p rows, compact lists, uniforms."""
    g = torch.Generator(device="cuda").manual_seed(1234 + seed)
    R = B * width
    z = 3.0 * torch.randn(R, gamma + 1, V, generator=g, device="cuda")
    target = (z + noise * torch.randn(R, gamma + 1, V, generator=g, device="cuda")).to(dtype)
    draft = (z[:, :gamma] + noise * torch.randn(R, gamma, V, generator=g, device="cuda")).to(dtype)
    del z
    u = torch.rand(R, 2 * gamma + 2, generator=g, device="cuda")
    q = torch.empty(R, gamma, V, device="cuda")
    pp = torch.empty(R, gamma + 1, V, device="cuda")
    use_cmp = 0 < k <= 128
    qc = ops.CompactRows(R * gamma, "cuda") if use_cmp else None
    pc = ops.CompactRows(R * (gamma + 1), "cuda") if use_cmp else None
    err = ops.ErrFlag("cuda")
    tok = ops.norm_sample(draft.view(R * gamma, V), T, k, p, u[:, :gamma].contiguous().view(-1), probs_out=q.view(R * gamma, V),
                          err=err, compact=qc.view() if use_cmp else None).view(R, gamma)
    ops.norm_probs(target.view(R * (gamma + 1), V), T, k, p, out=pp.view(R * (gamma + 1), V), err=err,
                   compact=pc.view() if use_cmp else None)
    torch.cuda.synchronize()
    err.check()
    return dict(q=q, p=pp, tok=tok, u_acc=u[:, gamma + 1:2 * gamma + 1].contiguous(), u_fin=u[:, 2 * gamma + 1].contiguous(),
                qc=qc, pc=pc, u=u)


def verify_bytes(n_acc, gamma, V):
    n_acc = n_acc.long()
    rej = (n_acc < gamma).sum().item()
    full = (n_acc >= gamma).sum().item()
    B = n_acc.numel()
    return B * (2 * gamma * 4 + 16) + rej * 2 * V * 4 + full * V * 4


def bench_verify(a, sparse):
    B, g, V = a.B, a.gamma, a.V
    T, k, p = (a.T, 20, 0.9) if sparse else (1.0, 0, 0.0)
    per_set = B * (2 * g + 1) * V * 4
    n_sets = max(2, min(8, int(2.2 * L2 / per_set) + 1))
    sets = [synth_probs(B, g, V, T, k, p, s) for s in range(n_sets)]
    n_acc = [torch.zeros(B, dtype=torch.int32, device="cuda") for _ in sets]
    nxt = torch.zeros(B, dtype=torch.int64, device="cuda")
    err = ops.ErrFlag("cuda")

    def mk(i):
        s = sets[i]
        kw = dict(p_compact=s["pc"].view(), p_cmp_req_stride=g + 1, q_compact=s["qc"].view(), q_cmp_req_stride=g) if sparse else {}
        return lambda: ops.verify(s["p"], s["q"], s["tok"], s["u_acc"], s["u_fin"], n_accepted=n_acc[i], next_tok=nxt, err=err, **kw)
    ms = graph_time([mk(i) for i in range(n_sets)], a.iters)
    err.check()
    alg = sum(verify_bytes(n, g, V) for n in n_acc) / n_sets
    extra = dict(B=B, gamma=g, V=V, top_k=k, top_p=p, mean_accepted=float(torch.stack(n_acc).float().mean()), input_sets=n_sets)
    if sparse:
        cnt = sum(int(s["pc"].cnt.clamp_min(0).sum()) + int(s["qc"].cnt.clamp_min(0).sum()) for s in sets) / n_sets
        extra["bytes_touched"] = int(B * (2 * g + 1) * 4 + cnt * 8 + B * (2 * g * 4 + 8 * g + 16))
        extra["note"] = "reads only the compact lists: latency-bound, the dense-definition GB/s is nominal"
    return emit("sd_verify (sparse path, verify_sparse_kernel)" if sparse else "sd_verify (dense path, verify_kernel)", ms, alg, extra)


def bench_verify_multi(a):
    B, g, V, W = a.B // 4 or 1, a.gamma, a.V, 4
    sets = [synth_probs(B, g, V, 1.0, 20, 0.9, s, width=W) for s in range(2)]
    err = ops.ErrFlag("cuda")
    outs = []

    def mk(s):
        u_acc = torch.rand(B, W * g, device="cuda")
        u_fin = s["u_fin"].view(B, W)[:, 0].contiguous()
        return lambda: outs.append(ops.verify_multi(s["p"].view(B, W, g + 1, V), s["q"].view(B, W, g, V), s["tok"].view(B, W, g), u_acc,
                                                    u_fin, err=err))
    ms = graph_time([mk(s) for s in sets], a.iters)
    err.check()
    na = outs[-1][1]
    alg = B * W * g * 2 * 4 + verify_bytes(na, g, V)
    return emit("sd_verify_multi (verify_multi_kernel)", ms, alg, dict(B=B, width=W, gamma=g, V=V, mean_accepted=float(na.float().mean())))


def bench_verify_bild(a):
    B, g, V = a.B, a.gamma, a.V
    sets = [synth_probs(B, g, V, 1.0, 20, 0.9, s) for s in range(2)]
    err = ops.ErrFlag("cuda")
    res = {}

    def mk(s, engine_mode):
        n_acc = torch.zeros(B, dtype=torch.int32, device="cuda")
        nt = torch.zeros(B, dtype=torch.int64, device="cuda")
        nd = torch.zeros(B, dtype=torch.int32, device="cuda")
        res[engine_mode] = (n_acc, nd)
        if engine_mode:
            return lambda: ops.verify_bild(s["p"], s["tok"], 3.0, s["u_fin"], q_probs=s["q"], fallback_thres=0.3, n_drafted=nd,
                                           n_accepted=n_acc, next_tok=nt, err=err)
        return lambda: ops.verify_bild(s["p"], s["tok"], 3.0, s["u_fin"], n_accepted=n_acc, next_tok=nt, err=err)
    out = []
    for engine_mode in (False, True):
        ms = graph_time([mk(s, engine_mode) for s in sets], a.iters)
        err.check()
        n_acc, nd = res[engine_mode]
        # every request samples one dense p row; engine mode also scans the q rows up to the fallback point for max q
        alg = B * (V * 4 + g * 4 + 16) + (int(nd.sum()) * V * 4 if engine_mode else 0)
        out.append(emit("sd_verify_bild (%s)" % ("engine mode: max q over dense q rows" if engine_mode else "check + sample"), ms, alg,
                        dict(B=B, gamma=g, V=V, mean_kept=float(n_acc.float().mean()))))
    return out


def bench_sample(a):
    rows, V = a.B * (2 * a.gamma + 1), a.V
    n_sets = max(2, min(8, int(2.2 * L2 / (rows * V * 4)) + 1))
    probs = [torch.softmax(torch.randn(rows, V, device="cuda") * 3.0, -1) for _ in range(n_sets)]
    u = torch.rand(rows, device="cuda")
    tok = torch.zeros(rows, dtype=torch.int64, device="cuda")
    err = ops.ErrFlag("cuda")
    ms = graph_time([(lambda pr=pr: ops.sample_rows(pr, u, tok_out=tok, err=err)) for pr in probs], a.iters)
    err.check()
    return emit("sd_sample", ms, rows * (V * 4 + 8), dict(rows=rows, V=V, input_sets=n_sets))


def bench_max_fn(a):
    rows, V = a.B * (2 * a.gamma + 1), a.V
    n_sets = max(2, min(8, int(2.2 * L2 / (rows * V * 8)) + 1))
    xs = [torch.randn(rows, V, device="cuda") * 0.01 for _ in range(n_sets)]
    lib_call = lambda x: ops.max_fn(x)          # noqa: E731  (allocates its output: measured including torch.empty)
    ms = graph_time([(lambda x=x: lib_call(x)) for x in xs], a.iters)
    return emit("sd_max_fn", ms, rows * V * 8, dict(rows=rows, V=V, input_sets=n_sets))


def bench_kv(a):
    out = []
    for name, layers, H, D, dt in (("llama-68m", 2, 12, 64, torch.float32), ("llama-2-13b", 40, 40, 128, torch.bfloat16)):
        B, S = a.B, 256
        es = torch.tensor([], dtype=dt).element_size()
        for q in (1, a.gamma + 1):
            k_new = torch.randn(B, H, q, D, device="cuda").to(dt)
            v_new = torch.randn(B, H, q, D, device="cuda").to(dt)
            n_l = min(layers, 8)
            kc = [torch.zeros(B, H, S, D, dtype=dt, device="cuda") for _ in range(n_l)]
            vc = [torch.zeros(B, H, S, D, dtype=dt, device="cuda") for _ in range(n_l)]
            pos = torch.randint(8, S - q - 1, (B,), device="cuda", dtype=torch.int32)
            ms = graph_time([(lambda i=i: ops.kv_append(k_new, v_new, kc[i], vc[i], pos)) for i in range(n_l)], a.iters)
            out.append(emit("sd_kv_append", ms, 2 * 2 * B * H * q * D * es, dict(model=name, B=B, H=H, D=D, new_rows=q, dtype=str(dt).split(".")[-1],
                                                                                   note="per layer; K and V, read + written")))
        W, Bq = 4, max(1, a.B // 4)
        kc = torch.randn(Bq * W, H, S, D, device="cuda").to(dt)
        vc = torch.randn(Bq * W, H, S, D, device="cuda").to(dt)
        choice = torch.randint(0, W, (Bq,), device="cuda", dtype=torch.int32)
        start = torch.randint(8, S - 8, (Bq * W,), device="cuda", dtype=torch.int32)
        count = torch.randint(0, a.gamma + 1, (Bq,), device="cuda", dtype=torch.int32)
        ms = graph_time([lambda: ops.kv_select(kc, vc, W, choice, start, W, count, a.gamma)] * 2, a.iters)
        nb = int(count.sum()) * H * D * es * 2
        out.append(emit("sd_kv_select", ms, nb + (W - 1) * nb, dict(model=name, requests=Bq, width=W, H=H, D=D, dtype=str(dt).split(".")[-1],
                                                                    note="per layer; rollback(end_pos, choice)")))
    return out


def bench_build_step(a):
    B, S, q = a.B, 256, a.gamma + 1
    tokens = torch.randint(3, 32000, (B, S), device="cuda")
    seq = torch.randint(8, S - 2 * q, (B,), device="cuda", dtype=torch.int32)
    ids = torch.zeros(B, q, dtype=torch.int64, device="cuda")
    pos = torch.zeros(B, q, dtype=torch.int64, device="cuda")
    wp = torch.zeros(B, dtype=torch.int32, device="cuda")
    mask = torch.zeros(B, 1, q, S, dtype=torch.uint8, device="cuda")
    cur = torch.randint(3, 32000, (B,), device="cuda")
    ms = graph_time([lambda: ops.build_step(tokens, seq, -1, q, cur, S, ids, pos, wp, mask)] * 2, a.iters)
    return emit("sd_build_step", ms, B * (q * S + q * 16 + 4 + 8), dict(B=B, S=S, q=q, note="mask (B,1,q,S) u8 + ids + positions"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="all")
    ap.add_argument("--V", type=int, default=32000)
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--gamma", type=int, default=4)
    ap.add_argument("--T", type=float, default=0.8)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--once", action="store_true", help="one eager launch per input set, no timing (for ncu)")
    ap.add_argument("--verify-cluster", type=int, default=0, help="cluster size of the dense verify kernel (0 = heuristic)")
    a = ap.parse_args()
    global ONCE
    ONCE = a.once
    build.build()
    ops.set_tuning(0, 0, a.verify_cluster)
    modes = {"verify_dense": lambda: bench_verify(a, False), "verify_sparse": lambda: bench_verify(a, True),
             "verify_multi": lambda: bench_verify_multi(a), "verify_bild": lambda: bench_verify_bild(a),
             "sample": lambda: bench_sample(a), "max_fn": lambda: bench_max_fn(a), "kv_append": lambda: bench_kv(a),
             "kv_select": lambda: bench_kv(a), "build_step": lambda: bench_build_step(a)}
    todo = [m for m in modes if m != "kv_select"] if a.mode == "all" else [a.mode]
    for m in todo:
        try:
            modes[m]()
        except Exception as e:  # noqa: BLE001
            print(json.dumps(dict(kernel=m, error=f"{type(e).__name__}: {e}"[:300])), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
