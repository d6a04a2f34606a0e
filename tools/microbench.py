"""Kernel micro-benchmarks (B200): HBM GB/s of the norm and verify kernels on synthetic logits.

    python tools/microbench.py [--V 32000] [--rows 576] [--dtype f32] [--sweep]

Inputs rotate over enough distinct buffers to exceed 2x the 126 MB L2, timing is CUDA events on
the launching stream after warm-up.
"""
import argparse
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmspeculativesampling_b200 import ops, build  # noqa: E402


PIPELINE = True
SCALE = 3.8


def time_norm(rows, V, dtype, T, k, p, iters=40, sample=False, write=True):
    es = torch.tensor([], dtype=dtype).element_size()
    per_set = rows * V * (es + (4 if write else 0))
    n_sets = max(2, int(2.2 * 126e6 / per_set) + 1)
    g = torch.Generator(device="cuda").manual_seed(1)
    ins = [(torch.randn(rows, V, device="cuda", generator=g) * SCALE).to(dtype) for _ in range(n_sets)]
    outs = [torch.empty(rows, V, device="cuda") for _ in range(n_sets)] if write else [None] * n_sets
    u = torch.rand(rows, device="cuda")
    tok = torch.empty(rows, dtype=torch.int64, device="cuda")

    def run(i):
        if sample:
            ops.norm_sample(ins[i % n_sets], T, k, p, u, probs_out=outs[i % n_sets], tok_out=tok, pipeline=PIPELINE)
        else:
            ops.norm_probs(ins[i % n_sets], T, k, p, out=outs[i % n_sets], pipeline=PIPELINE)
    for i in range(5):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    ops.default_flag("cuda").check()
    ms = e0.elapsed_time(e1) / iters
    return ms, per_set / ms / 1e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--V", type=int, default=32000)
    ap.add_argument("--rows", type=int, default=576)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--mode", default="")
    ap.add_argument("--prof", action="store_true")
    ap.add_argument("--scale", type=float, default=3.8, help="standard deviation of the synthetic logits")
    ap.add_argument("--classic", action="store_true")
    ap.add_argument("--no-ring", action="store_true", help="skip the ring kernel: persistent cluster pipeline where it applies")
    ap.add_argument("--sample", action="store_true", help="time sd_norm_sample (one token per row) instead of sd_norm_probs")
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--k", type=int, default=-1, help="override the mode's top_k")
    ap.add_argument("--p", type=float, default=-1.0, help="override the mode's top_p")
    ap.add_argument("--cluster", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    global PIPELINE, SCALE
    PIPELINE = False if a.classic else ("cluster" if a.no_ring else True)
    SCALE = a.scale
    build.build()
    dt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[a.dtype]
    modes_all = {"topk": ("topk20_p0.9", 0.8, 20, 0.9), "dense": ("dense", 1.0, 0, 0.0), "topp": ("top_p_only", 1.0, 0, 0.9)}
    if a.mode:
        ops.set_tuning(a.cluster, a.threads, 0)
        name, T, k, p = modes_all[a.mode]
        k = a.k if a.k >= 0 else k
        p = a.p if a.p >= 0 else p
        ms, gbs = time_norm(a.rows, a.V, dt, T, k, p, iters=a.iters, sample=a.sample)
        print(json.dumps(dict(kernel="norm_sample" if a.sample else "norm", mode=name, rows=a.rows, V=a.V, dtype=a.dtype, ms=round(ms, 4), GBs=round(gbs, 1))))
        if a.prof:
            from llmspeculativesampling_b200 import _cabi
            buf = torch.zeros(a.rows * 8 + 8192, 16, dtype=torch.int64, device="cuda")
            x = (torch.randn(a.rows, a.V, device="cuda") * SCALE).to(dt)
            out = torch.empty(a.rows, a.V, device="cuda")
            uu = torch.rand(a.rows, device="cuda")
            call = (lambda: ops.norm_sample(x, T, k, p, uu, probs_out=out, pipeline=PIPELINE)) if a.sample else \
                   (lambda: ops.norm_probs(x, T, k, p, out=out, pipeline=PIPELINE))
            call()
            torch.cuda.synchronize()
            _cabi.load().sd_debug_set_prof(buf.data_ptr())
            call()
            torch.cuda.synchronize()
            _cabi.load().sd_debug_set_prof(None)
            b = buf.cpu()
            if PIPELINE and a.mode == "topk":
                t = b[: (b.shape[0] // 32) * 32].view(-1, 32, 16)
                t = t[t[:, 0, 0] != 0]
                print("persistent CTAs", t.shape[0])
                ent, ext, g0, g1 = t[:, 0, 15], t[:, 1, 15], t[:, 2, 15], t[:, 3, 15]
                print(f"  kernel entry -> first item start: {(t[:, 0, 0] - ent).double().mean() / 1000:.2f} kcyc; "
                      f"CTA lifetime (clock64): {(ext - ent).double().mean() / 1000:.2f} kcyc; "
                      f"globaltimer: first entry -> last exit {(g1.max() - g0.min()).item() / 1000:.2f} us, "
                      f"entry spread {(g0.max() - g0.min()).item() / 1000:.2f} us, exit spread {(g1.max() - g1.min()).item() / 1000:.2f} us")
                t0 = t[:, 0, 0].clone()
                last_done = t[:, :, 11].max(dim=1).values
                tail = (ext - last_done).double() / 1000
                life = (ext - ent).double() / 1000
                n_it = (t[:, :, 11] != 0).sum(dim=1)
                print(f"  exit - last item done: mean {tail.mean():.2f} kcyc, max {tail.max():.2f}; lifetime p10/p50/p90/max "
                      f"{life.quantile(0.1):.1f}/{life.quantile(0.5):.1f}/{life.quantile(0.9):.1f}/{life.max():.1f} kcyc; "
                      f"items per CTA {int(n_it.min())}..{int(n_it.max())}; lifetime of CTAs with max items {life[n_it == n_it.max()].mean():.1f}, others {life[n_it != n_it.max()].mean() if (n_it != n_it.max()).any() else float('nan'):.1f}")
                names = ["mem: item start", "mem: buffer free", "mem: load issued + zero-filled", "grp: wait for data", "grp: data landed",
                         "grp: pass1 done", "grp: pivot done", "grp: rescan done", "grp: buffer released", "grp: peers landed",
                         "grp: sorted", "grp: item done", "grp: compacted", "grp: ranks done", "grp: hot list"]
                for it in range(0, 32):
                    row = []
                    for s_ in [0, 1, 2, 3, 4, 5, 6, 14, 7, 8, 9, 12, 13, 10, 11]:
                        v = t[:, it, s_]
                        m = v != 0
                        row.append(f"{((v - t0)[m]).double().mean() / 1000:6.2f}" if m.any() else "   -  ")
                    if (t[:, it, 0] != 0).any() or it == 0:
                        print(f"  item {it:2d}: " + " ".join(row))
                order = [0, 1, 2, 3, 4, 5, 6, 14, 7, 8, 9, 12, 13, 10, 11]
                print("  columns (kcycles since kernel start, mean over CTAs): " + " | ".join(names[o] if o < 12 else names[o] for o in [0,1,2,3,4,5,6] ) + " | hot list | rescan done | released | peers landed | compacted | ranks done | sorted | item done")
                return
            b = b[b[:, 0] != 0]
            names = {1: "setup+issue", 2: "first chunk landed", 3: "pass1 done", 11: "warp sort+sync", 12: "rank+sync", 13: "hot list+sync", 4: "rescan done", 5: "cluster sync", 6: "merge+sort", 7: "select+scatter", 8: "dense sums", 9: "dense write", 10: "exit"}
            prev = 0
            print("CTAs", b.shape[0])
            for s_ in [1, 2, 3, 11, 12, 13, 4, 5, 6, 7, 8, 9, 10]:
                m = b[:, s_] != 0
                if m.any():
                    d = (b[m, s_] - b[m, prev]).float()
                    print(f"  slot {s_:2d} {names[s_]:20s} +{d.mean():9.0f} cyc (p10 {d.quantile(0.1):7.0f}, p90 {d.quantile(0.9):7.0f})")
                    prev = s_
            if (b[:, 14] != 0).any():
                print(f"  merged candidates per row: mean {b[:, 14].float().mean():.1f}, max {int(b[:, 14].max())}")
            tot = (b[:, 10] - b[:, 0]).float()
            print(f"  CTA lifetime mean {tot.mean():.0f} cyc, p90 {tot.quantile(0.9):.0f}")
        return
    # reference point: a plain device copy of the same number of bytes
    n = a.rows * a.V
    src = [torch.randn(n, device="cuda") for _ in range(4)]
    dst = [torch.empty(n, device="cuda") for _ in range(4)]
    for i in range(4):
        dst[i].copy_(src[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        dst[i % 4].copy_(src[i % 4])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(json.dumps(dict(kernel="torch_copy_same_bytes", ms=round(ms, 4), GBs=round(n * 8 / ms / 1e6, 1))), flush=True)
    del src, dst
    res = []
    tunings = [(0, 0)] + ([(1, 512), (1, 1024), (2, 256), (2, 512), (4, 256), (8, 256)] if a.sweep else [])
    modes = [("topk20_p0.9", 0.8, 20, 0.9), ("dense", 1.0, 0, 0.0), ("top_p_only", 1.0, 0, 0.9), ("dense_T0.8", 0.8, 0, 0.0)]
    for (c, t) in tunings:
        ops.set_tuning(c, t, 0)
        for name, T, k, p in modes:
            try:
                ms, gbs = time_norm(a.rows, a.V, dt, T, k, p)
                res.append(dict(kernel="norm", mode=name, cluster=c, threads=t, rows=a.rows, V=a.V, dtype=a.dtype, ms=round(ms, 4), GBs=round(gbs, 1)))
            except Exception as e:  # noqa: BLE001
                res.append(dict(kernel="norm", mode=name, cluster=c, threads=t, error=str(e)[:80]))
            print(json.dumps(res[-1]), flush=True)
    ops.set_tuning(0, 0, 0)
    ms, gbs = time_norm(a.rows, a.V, dt, 0.8, 20, 0.9, sample=True)
    print(json.dumps(dict(kernel="norm_sample", ms=round(ms, 4), GBs=round(gbs, 1))))
    ms, gbs = time_norm(a.rows, a.V, dt, 0.8, 20, 0.9, sample=True, write=False)
    print(json.dumps(dict(kernel="norm_sample_tokens_only", ms=round(ms, 4), GBs_read=round(gbs, 1))))


if __name__ == "__main__":
    main()
