"""BASELINE.json config 5: sweep of V x gamma x batch on synthetic logits, HBM roofline fraction per point.

    python tools/sweep.py [--full] > sweep.jsonl
    torchrun --nproc-per-node 8 tools/sweep.py --subset headline > sweep_n8.jsonl     (replicas: slowest rank per point)

Each point runs kernel 1 (sd_norm_sample: filter + softmax + draft token) over the B*(2*gamma+1) rows of one batch,
T=0.8, top_k=20, top_p=0.9 (mode topk) or T=1, top_k=0, top_p=0 (mode dense, the API default), inputs/outputs rotating over > 2x L2, CUDA-graph replays timed with CUDA events.
Algorithmic bytes = rows * V * (sizeof(logit) + 4).  Peak = MEASURED_PEAKS.json hbm_gbs.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llmspeculativesampling_b200 import build, ops  # noqa: E402


PIPELINE = True


def point(V, gamma, B, dtype, peak, mode):
    rows = B * (2 * gamma + 1)
    es = torch.tensor([], dtype=dtype).element_size()
    per_set = rows * V * (es + 4)
    n_sets = max(2, min(16, int(2.2 * 126e6 / per_set) + 1))
    if n_sets * per_set > 40e9:
        return None
    T, k, p = (0.8, 20, 0.9) if mode == "topk" else (1.0, 0, 0.0)
    g = torch.Generator(device="cuda").manual_seed(1)
    ins = [(torch.randn(rows, V, device="cuda", generator=g) * 3.8).to(dtype) for _ in range(n_sets)]
    outs = [torch.empty(rows, V, device="cuda") for _ in range(n_sets)]
    u = torch.rand(rows, device="cuda")
    tok = torch.empty(rows, dtype=torch.int64, device="cuda")
    cmp_rows = ops.CompactRows(rows, "cuda")
    c = cmp_rows.view()
    err = ops.ErrFlag("cuda")                                # own scheduler workspace: usable inside the graph capture
    for i in range(n_sets):
        ops.norm_sample(ins[i], T, k, p, u, probs_out=outs[i], tok_out=tok, compact=c, err=err, pipeline=PIPELINE)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(gr, stream=side):
            for i in range(n_sets):
                ops.norm_sample(ins[i], T, k, p, u, probs_out=outs[i], tok_out=tok, compact=c, err=err, pipeline=PIPELINE)
    torch.cuda.synchronize()
    reps = max(3, int(0.02 / max(per_set * n_sets / 5e12, 1e-6)))
    reps = min(reps, 200)
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    err.check()
    ms = e0.elapsed_time(e1) / (reps * n_sets)
    gbs = per_set / ms / 1e6
    return dict(V=V, gamma=gamma, batch=B, rows=rows, dtype=str(dtype).split(".")[-1], mode=mode, ms=round(ms, 4),
                GBs=round(gbs, 1), frac_of_measured=round(gbs / peak, 3), frac_of_8TBs=round(gbs / 8000, 3))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--subset", default="all", choices=["all", "headline", "draft"],
                    help="headline: the B >= 64 points of V = 32000 / 50272 only (the multi-GPU run); draft: the launch shape of "
                         "an engine's draft step (one row per request: gamma = 0)")
    ap.add_argument("--no-ring", action="store_true", help="cluster pipeline / one-cluster-per-row kernels instead of the ring kernel")
    a = ap.parse_args()
    global PIPELINE
    PIPELINE = "cluster" if a.no_ring else True
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:                                             # N GPUs: every rank runs the same points (replicas, no collective
        import torch.distributed as dist                      # in the path); a point reports its slowest rank
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    build.build()
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except OSError:
        peak = 6650.0
    pts = []
    if a.subset == "draft":
        for dtype in (torch.float32, torch.bfloat16):
            for mode in ("topk", "dense"):
                for V in (32000, 50272, 131072):
                    for B in (16, 32, 64, 128):
                        pts.append((V, 0, B, dtype, mode))
    elif a.subset == "headline":
        for dtype in (torch.float32, torch.bfloat16):
            for mode in ("topk", "dense"):
                for V in (32000, 50272):
                    for B in (64, 256):
                        pts.append((V, 4, B, dtype, mode))
    else:
        for dtype in (torch.float32, torch.bfloat16):
            for V in (32000, 50272, 65536, 131072, 262144):
                for B in (1, 8, 64, 256, 512):
                    pts.append((V, 4, B, dtype, "topk"))
            for V in (32000, 50272):
                for gamma in (1, 2, 8, 16):
                    for B in ((1, 8, 64, 256, 512) if a.full else (64, 256)):
                        pts.append((V, gamma, B, dtype, "topk"))
            for V in (32000, 50272, 131072):
                for B in (8, 64, 256):
                    pts.append((V, 4, B, dtype, "dense"))
            for gamma in (1, 16):
                pts.append((32000, gamma, 64, dtype, "dense"))
    for pt in pts:
        try:
            r = point(*pt[:4], peak, pt[4])
        except Exception as e:  # noqa: BLE001
            r = dict(V=pt[0], gamma=pt[1], batch=pt[2], dtype=str(pt[3]), mode=pt[4], error=str(e)[:100])
        if world > 1:
            t = torch.tensor([r["ms"] if r is not None and "ms" in r else -1.0], dtype=torch.float64, device="cuda")
            allt = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            if r is not None and "ms" in r:
                ms_all = [float(x) for x in allt]
                worst = max(ms_all)
                per_set = r["rows"] * r["V"] * ((4 if r["dtype"] == "float32" else 2) + 4)
                r.update(n_gpus=world, ms_per_rank=[round(x, 4) for x in ms_all], ms=round(worst, 4), GBs=round(per_set / worst / 1e6, 1),
                         frac_of_measured=round(per_set / worst / 1e6 / peak, 3), frac_of_8TBs=round(per_set / worst / 1e6 / 8000, 3),
                         aggregate_GBs=round(world * per_set / worst / 1e6, 1))
        if r is not None and rank == 0:
            print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
