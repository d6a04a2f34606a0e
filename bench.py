#!/usr/bin/env python
"""Headline benchmark of the speculative-decoding draft-and-verify hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): synthetic-logits
verify microbench, V=32000, gamma=4, batch=64 requests per GPU, T=0.8, top_k=20, top_p=0.9, fp32
logits.  ONE STEP = one pass of the hot path over one batch:
    kernel 1/1b  sd_norm_sample  ONE launch over the B*(2*gamma+1) rows of the batch (per request: gamma draft rows
                                 that also draw the drafted token from their uniform, then gamma+1 target rows whose
                                 uniform is -1 = "no sample"): filter + softmax (+ draft token)
    kernel 2     sd_verify       accept / first reject / residual / inverse-CDF sample / append
Metric: accepted tokens/s (whole job, all GPUs); emitted tokens/s and the acceptance are reported too.

value     inputs resident in HBM, steps replayed from CUDA graphs, timed with CUDA events around exactly K steps (barrier +
          synchronize on both sides, max over ranks).  Inputs rotate over distinct sets whose total exceeds 2x the 126 MB
          L2.  By default the steps are software-pipelined (--pipeline 2): a 16-step graph keeps independent batches in
          flight, as a serving loop with more than one batch would run them — kernel 2 of batch i runs on a second stream
          beside kernel 1 of batch i+1, and kernel 1 of consecutive batches alternates between two streams (own scheduler
          workspace each), so the persistent CTAs of batch i+1 take over every SM as soon as batch i's CTA on it has exited:
          one launch's drain (selection latency of its last rows, slowest CTA, grid completion) overlaps the next
          launch's ramp instead of leaving HBM idle.  `serial_ms_per_step` is the same step with its kernels strictly one
          after the other (the latency of ONE batch); --pipeline 1 / 0 select the older schedules.
e2e       same steps through the public tensor API with HOST buffers: every step copies its logits and uniforms from
          pinned host memory to the device (double-buffered on a copy stream) and reads accept counts and tokens back.
roofline  dominant kernel (norm, 1 launch per step): algorithmic bytes (rows * V * (4 read + 4 written)) / its average
          launch duration from back-to-back graph replays ON ONE STREAM (>= 240 launches whatever --steps is); peak =
          MEASURED_PEAKS.json hbm_gbs.  `roofline.overlapped` is the time per launch when the same launches alternate
          between two streams, `roofline.step` the whole step (both kernels) over the timed ms_per_step,
          `roofline.verify_dense` kernel 2's dense path (top_k = 0) at V = 32000 / 50272.
gpu_aten_baseline   the reference's ATen op chain on the same B200, batch 1 as the reference runs (oracle/aten_gpu.py).
cpu_baseline / --impl reference   the oracle port of the reference's CPU path (oracle/ref_ops.py: the same
          ATen op chain, one row at a time, host syncs included) on a bounded sample of the same workload.
config1   BASELINE.json configs[0] (reduced): the oracle loop on one host core vs the drop-in on cuda:0, tokens compared.

N > 1 (torchrun): requests are independent, every rank runs the same per-GPU batch (identical synthetic sets)
with no data-path collective ("scaling": "weak"); NCCL is used for the barrier and the final
statistics reduction only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

V, GAMMA, BATCH = 32000, 4, 64
TEMP, TOP_K, TOP_P = 0.8, 20, 0.9
WORKLOAD = f"synthetic-logits verify microbench V={V} gamma={GAMMA} batch={BATCH}/GPU T={TEMP} top_k={TOP_K} top_p={TOP_P} fp32 logits"


def synth_set(seed: int, device, batch: int = BATCH):
    """shared base z, draft = z + noise, target = z + noise'."""
    g = torch.Generator(device=device).manual_seed(1234 + seed)
    z = 3.0 * torch.randn(batch, GAMMA + 1, V, generator=g, device=device)
    target = z + 0.5 * torch.randn(batch, GAMMA + 1, V, generator=g, device=device)
    draft = z[:, :GAMMA] + 0.5 * torch.randn(batch, GAMMA, V, generator=g, device=device)
    u = torch.rand(batch, 2 * GAMMA + 2, generator=g, device=device)
    return draft.contiguous(), target.contiguous(), u


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6 or not f[0].isdigit():
                continue
            sm.append(int(f[0])); mx = int(f[1])
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------- reference arm
def oracle_step(draft, target, u, requests):
    """The reference's CPU path for `requests` of one batch (oracle port; rows one at a time)."""
    from oracle import ref_ops
    acc = emitted = 0
    for b in requests:
        q_rows = torch.cat([ref_ops.norm_probs(draft[b, i:i + 1], TEMP, TOP_K, TOP_P) for i in range(GAMMA)], 0)
        toks = torch.tensor([ref_ops.icdf_sample(q_rows[i], float(u[b, i])) for i in range(GAMMA)])
        p_rows = torch.cat([ref_ops.norm_probs(target[b, i:i + 1], TEMP, TOP_K, TOP_P) for i in range(GAMMA + 1)], 0)
        _ = ref_ops.icdf_sample(p_rows[GAMMA], float(u[b, GAMMA]))            # the sample the reference discards
        n_acc, _, _, _ = ref_ops.verify_request(p_rows, q_rows, toks, u[b, GAMMA + 1:2 * GAMMA + 1].numpy(),
                                                float(u[b, 2 * GAMMA + 1]), residual="normalised")
        acc += n_acc
        emitted += n_acc + 1
    return acc, emitted


def time_oracle(steps: int, warmup: int, per_step: int, threads: int):
    torch.set_num_threads(threads)
    draft, target, u = synth_set(0, "cpu")
    acc = emitted = 0
    for s in range(warmup):
        oracle_step(draft, target, u, [(s * per_step + j) % BATCH for j in range(per_step)])
    t0 = time.perf_counter()
    for s in range(steps):
        a, e = oracle_step(draft, target, u, [((warmup + s) * per_step + j) % BATCH for j in range(per_step)])
        acc += a; emitted += e
    dt = time.perf_counter() - t0
    return acc, emitted, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    per_step = 4
    cores = os.cpu_count() or 1
    # the reference is single-stream Python; intra-op threads are the only host parallelism it can use
    best = None
    for th in sorted({1, cores}):
        a, e, dt = time_oracle(max(1, min(args.steps, 3)), 1, per_step, th)
        if best is None or a / dt > best[0]:
            best = (a / dt, th)
    threads = best[1]
    acc, emitted, dt = time_oracle(args.steps, args.warmup, per_step, threads)
    val = acc / dt
    line = {
        "impl": "reference", "metric": "accepted_tokens_per_s", "value": val, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{per_step} of the {BATCH} requests per step"},
        "emitted_tokens_per_s": emitted / dt, "mean_accepted_per_iteration": acc / (args.steps * per_step),
        "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps x {per_step} requests (gamma+gamma+1 rows each), oracle port of the "
                                   f"reference CPU path, torch threads={threads} of {cores} host cores"},
        "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------- config 1 (side report)
def run_config1(args):
    """BASELINE.json configs[0]: llama-68m-shape draft + target (random init), gamma=4, batch=1, 128 new tokens, fp32.
    The reference's own CPU-runnable case: the oracle port of its KV-cached loop is timed on the host cores and the
    drop-in speculative_sampling on cuda:0, same weights, same prompt, same uniform tape.  Reported for T=1 with
    (top_k=20, top_p=0.9) and (0, 0), and for identical / independent weights (acceptance 1.0 / ~0)."""
    from transformers import LlamaConfig, LlamaForCausalLM
    from llmspeculativesampling_b200 import build, uniform_tape
    from llmspeculativesampling_b200.sampling import speculative_sampling
    from oracle import spec_loop
    build.build()
    cfg = LlamaConfig(vocab_size=32000, hidden_size=768, intermediate_size=3072, num_hidden_layers=2,
                      num_attention_heads=12, num_key_value_heads=12, max_position_embeddings=2048)
    prompt = torch.randint(3, 32000, (1, 16), generator=torch.Generator().manual_seed(7))
    out = []
    for pair in ("identical", "independent"):
        torch.manual_seed(0); draft = LlamaForCausalLM(cfg).eval()
        torch.manual_seed(0 if pair == "identical" else 1); target = LlamaForCausalLM(cfg).eval()
        for (k, p) in ((20, 0.9), (0, 0.0)):
            tape = uniform_tape.batch_tape(3, [0], 129, GAMMA)
            torch.set_num_threads(1)
            t0 = time.perf_counter()
            ref_tok, ref_d = spec_loop.speculative_sampling(prompt, draft, target, 128, GAMMA, 1.0, k, p, tape=tape[:, 0])
            cpu_s = time.perf_counter() - t0
            dg, tg = draft.cuda(), target.cuda()
            speculative_sampling(prompt.cuda(), dg, tg, None, None, 128, GAMMA, 1.0, k, p, uniforms=tape)   # warm-up + graph
            torch.cuda.synchronize(); t0 = time.perf_counter()
            speculative_sampling(prompt.cuda(), dg, tg, None, None, 128, GAMMA, 1.0, k, p, uniforms=tape)
            torch.cuda.synchronize(); gpu_s = time.perf_counter() - t0
            tok, d = speculative_sampling(prompt.cuda(), dg, tg, None, None, 128, GAMMA, 1.0, k, p, uniforms=tape, details=True)   # (untimed: statistics)
            n_ref, n_gpu = ref_tok.shape[1] - 16, tok.shape[1] - 16
            same = min(n_ref, n_gpu)
            agree = int((ref_tok[0, 16:16 + same] == tok[0, 16:16 + same].cpu()).long().cumprod(0).sum())
            out.append({"weights": pair, "top_k": k, "top_p": p,
                        "cpu_oracle": {"seconds": cpu_s, "emitted_tokens_per_s": n_ref / cpu_s, "accepted_tokens_per_s": sum(ref_d["acc_len"]) / cpu_s,
                                       "iterations": ref_d["iterations"], "mean_accepted": sum(ref_d["acc_len"]) / max(1, len(ref_d["acc_len"])), "threads": 1},
                        "b200": {"seconds": gpu_s, "emitted_tokens_per_s": n_gpu / gpu_s, "accepted_tokens_per_s": sum(d["acc_len"]) / gpu_s,
                                 "iterations": d["iterations"], "mean_accepted": sum(d["acc_len"]) / max(1, len(d["acc_len"])), "cuda_graph": d["cuda_graph"]},
                        "leading_tokens_identical_to_cpu_run": agree, "of": same})
            draft.cpu(); target.cpu()
    # per-op micro-timings: one row, T=0.8 top_k=20 top_p=0.9, CPU oracle (1 thread) vs the kernels
    from oracle import ref_ops
    from llmspeculativesampling_b200 import ops
    micro = []
    for V_ in (32000, 50272):
        x = torch.randn(1, V_, generator=torch.Generator().manual_seed(V_)) * 3.8
        pr = ref_ops.norm_probs(x, 0.8, 20, 0.9)
        dense = torch.softmax(x, -1)
        cpu = {}
        for name, fn in (("norm_logits", lambda: ref_ops.norm_probs(x, 0.8, 20, 0.9)), ("sample", lambda: ref_ops.icdf_sample(dense[0], 0.37)),
                         ("max_fn", lambda: ref_ops.max_fn(dense - pr))):
            fn(); t0 = time.perf_counter()
            for _ in range(20): fn()
            cpu[name] = (time.perf_counter() - t0) / 20 * 1e3
        xg, dg, pg, ug = x.cuda(), dense.cuda(), pr.cuda(), torch.tensor([0.37], device="cuda")
        gpu = {}
        for name, fn in (("norm_logits", lambda: ops.norm_probs(xg, 0.8, 20, 0.9)), ("sample", lambda: ops.sample_rows(dg, ug)),
                         ("max_fn", lambda: ops.max_fn(dg - pg))):
            for _ in range(5): fn()
            torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50): fn()
            e1.record(); torch.cuda.synchronize()
            gpu[name] = e0.elapsed_time(e1) / 50
        micro.append({"V": V_, "cpu_oracle_ms_per_row_1thread": cpu, "b200_ms_per_call_batch1": gpu})
    print(json.dumps({"workload": "config 1: llama-68m-shape draft + target, gamma=4, batch=1, 128 new tokens, fp32", "results": out,
                      "per_op_one_row": micro}))
    return 0


# ------------------------------------------------------------------------------------------- next rows (side report)
def run_variants(args):
    """SURVEY 8f rows N2 / N3 on the config-1 models (llama-68m-shape draft + target, fp32, batch 1, 64 new tokens):
    multi_speculative_sampling(strategy='iid', width=4) and BiLD_sampling — the oracle port of the reference's loop timed
    on one host core next to the drop-in on cuda:0, same weights, prompt and uniform tape; leading tokens compared."""
    from transformers import LlamaConfig, LlamaForCausalLM
    from llmspeculativesampling_b200 import build, uniform_tape
    from llmspeculativesampling_b200.sampling import multi_speculative_sampling, BiLD_sampling
    from oracle import spec_loop, ref_loader
    build.build()
    cfg = LlamaConfig(vocab_size=32000, hidden_size=768, intermediate_size=3072, num_hidden_layers=2,
                      num_attention_heads=12, num_key_value_heads=12, max_position_embeddings=2048)
    prompt = torch.randint(3, 32000, (1, 16), generator=torch.Generator().manual_seed(7))
    N, W, k, p = 64, 4, 20, 0.9
    out = []
    torch.manual_seed(0); draft = LlamaForCausalLM(cfg).eval()
    torch.manual_seed(0); target = LlamaForCausalLM(cfg).eval()          # identical weights: non-trivial acceptance
    torch.set_num_threads(1)

    def agree(a, b):
        n = min(a.shape[1], b.shape[1]) - 16
        return int((a[0, 16:16 + n] == b[0, 16:16 + n].cpu()).long().cumprod(0).sum()), n

    # ---- N2: multi-draft, iid
    tape = torch.rand(N + 1, spec_loop.multi_block(GAMMA, W), generator=torch.Generator().manual_seed(3))
    da, ta = ref_loader.LegacyCacheAdapter(draft), ref_loader.LegacyCacheAdapter(target)
    t0 = time.perf_counter()
    ref_tok, ref_d = spec_loop.multi_speculative_sampling(prompt, da, ta, N, GAMMA, W, 1.0, k, p, tape=tape)
    cpu_s = time.perf_counter() - t0
    dg, tg = draft.cuda(), target.cuda()
    multi_speculative_sampling(prompt.cuda(), dg, tg, None, None, N, GAMMA, W, None, "iid", None, 0.4, 1.0, k, p, uniforms=tape)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tok, d = multi_speculative_sampling(prompt.cuda(), dg, tg, None, None, N, GAMMA, W, None, "iid", None, 0.4, 1.0, k, p,
                                        uniforms=tape, details=True)
    torch.cuda.synchronize(); gpu_s = time.perf_counter() - t0
    a, n = agree(ref_tok, tok)
    out.append({"row": "N2 multi_speculative_sampling(strategy='iid')", "width": W, "gamma": GAMMA, "top_k": k, "top_p": p,
                "cpu_oracle": {"seconds": cpu_s, "emitted_tokens_per_s": (ref_tok.shape[1] - 16) / cpu_s, "mean_accepted": float(np.mean(ref_d["acc_len"])), "threads": 1},
                "b200": {"seconds": gpu_s, "emitted_tokens_per_s": (tok.shape[1] - 16) / gpu_s, "mean_accepted": float(np.mean(d["acc_len"]))},
                "leading_tokens_identical_to_cpu_run": a, "of": n})
    draft.cpu(); target.cpu()
    # ---- N3: BiLD
    fb, rb = 0.08, 3.0
    tape = uniform_tape.make_tape(5, N + 1, GAMMA)
    t0 = time.perf_counter()
    ref_tok, ref_d = spec_loop.bild_sampling(prompt, draft, target, N, GAMMA, fb, rb, 1.0, k, p, tape=tape)
    cpu_s = time.perf_counter() - t0
    dg, tg = draft.cuda(), target.cuda()
    BiLD_sampling(prompt.cuda(), dg, tg, GAMMA, None, None, fb, rb, N, 1.0, k, p, uniforms=tape)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tok, d = BiLD_sampling(prompt.cuda(), dg, tg, GAMMA, None, None, fb, rb, N, 1.0, k, p, uniforms=tape, details=True)
    torch.cuda.synchronize(); gpu_s = time.perf_counter() - t0
    a, n = agree(ref_tok, tok)
    out.append({"row": "N3 BiLD_sampling", "gamma": GAMMA, "fallback_thres": fb, "rollback_thres": rb, "top_k": k, "top_p": p,
                "cpu_oracle": {"seconds": cpu_s, "emitted_tokens_per_s": (ref_tok.shape[1] - 16) / cpu_s, "target_calls": ref_d["target_call_times"],
                               "draft_tokens": ref_d["approx_call_times"], "threads": 1},
                "b200": {"seconds": gpu_s, "emitted_tokens_per_s": (tok.shape[1] - 16) / gpu_s, "target_calls": d["target_call_times"],
                         "draft_tokens": d["approx_call_times"]},
                "leading_tokens_identical_to_cpu_run": a, "of": n})
    # ---- the same two rows with 16 and 64 requests in flight (ragged prompts, one CUDA graph per iteration / check cycle)
    batched = []
    for Bq in (16, 64):
        gq = torch.Generator().manual_seed(11)
        prompts = [torch.randint(3, 32000, (int(n),), generator=gq).cuda() for n in torch.randint(8, 24, (Bq,), generator=gq)]
        tape_m = torch.rand(N + 1, Bq, spec_loop.multi_block(GAMMA, W), generator=gq)
        tape_b = uniform_tape.batch_tape(5, list(range(Bq)), N + 1, GAMMA)
        for name, fn in (("N2 multi_speculative_sampling(strategy='iid'), width 4", lambda: multi_speculative_sampling(
                              prompts, dg, tg, None, None, N, GAMMA, W, None, "iid", None, 0.4, 1.0, k, p, uniforms=tape_m, details=True)),
                         ("N3 BiLD_sampling", lambda: BiLD_sampling(prompts, dg, tg, GAMMA, None, None, fb, rb, N, 1.0, k, p, uniforms=tape_b,
                                                                    details=True))):
            fn()                                                   # engine construction + graph capture
            torch.cuda.synchronize(); t0 = time.perf_counter()
            outs, det = fn()
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            emitted = sum(o.shape[1] - pr.numel() for o, pr in zip(outs, prompts))
            its = det.get("iterations", det.get("target_call_times"))
            its = max(its) if isinstance(its, (list, tuple)) else its
            batched.append({"row": name, "requests_in_flight": Bq, "seconds": dt, "emitted_tokens_per_s": emitted / dt,
                            "graph_iterations": its, "ms_per_iteration": dt / max(int(its or 1), 1) * 1e3})
    print(json.dumps({"workload": "SURVEY 8f next rows on the config-1 models (llama-68m shapes, identical weights, fp32, batch 1, 64 new tokens)",
                      "note": "both run on their batched CUDA-graph engines (multi_engine.MultiDraftEngine, bild_engine.BiLDEngine), batch 1 here",
                      "results": out, "batched": batched}))
    return 0


# ------------------------------------------------------------------------------------------- B200 arm
PIPE_STEPS = 16          # steps captured in the software-pipelined graph


def gpu_aten_baseline(logits_set, dev, n_req=8):
    """SURVEY §2.2 / BASELINE.md §3: the reference's ATen op chain on the SAME B200 (oracle/aten_gpu.py restates it launch
    for launch, host syncs included), one request at a time as the reference runs, CUDA-event timed."""
    from oracle import aten_gpu
    g = GAMMA
    lg = logits_set[:n_req]                                  # (n_req, 2g+1, V): gamma draft rows then gamma+1 target rows
    torch.manual_seed(0)
    for b in range(2):
        aten_gpu.iteration(lg[b, :g], lg[b, g:], TEMP, TOP_K, TOP_P)
    torch.cuda.synchronize()
    launches = None
    try:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            aten_gpu.iteration(lg[0, :g], lg[0, g:], TEMP, TOP_K, TOP_P)
            torch.cuda.synchronize()
        launches = sum(1 for e in prof.events() if getattr(e, "device_type", None) is not None and "cuda" in str(e.device_type).lower())
    except Exception:                                        # noqa: BLE001  (profiler unavailable: the count stays None)
        launches = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc = 0
    e0.record()
    for b in range(n_req):
        n, _ = aten_gpu.iteration(lg[b, :g], lg[b, g:], TEMP, TOP_K, TOP_P)
        acc += n
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"accepted_tokens_per_s": acc / (ms * 1e-3), "ms_per_request_iteration": ms / n_req, "requests": n_req,
            "kernel_launches_per_request_iteration": launches,
            "what": "oracle/aten_gpu.py: the reference's per-row ATen op chain (utils.py:152-245, speculative_sampling.py:1966-2027) on "
                    "CUDA tensors of the same workload, batch 1 as the reference runs, torch.multinomial / torch.rand on the device"}


def config1_side_report():
    """BASELINE.json configs[0] (the one case the reference runs as-is), reduced so that it fits the default run:
    llama-68m-shape draft + target with IDENTICAL random weights (acceptance 1.0), gamma=4, batch=1, 64 new tokens, fp32,
    T=1 top_k=20 top_p=0.9: the oracle port of the reference loop on one host core vs the drop-in on cuda:0 — same weights,
    prompt and uniform tape; leading tokens compared.  The full report is `bench.py --workload config1`."""
    from transformers import LlamaConfig, LlamaForCausalLM
    from llmspeculativesampling_b200 import uniform_tape
    from llmspeculativesampling_b200.sampling import speculative_sampling
    from oracle import spec_loop
    cfg = LlamaConfig(vocab_size=32000, hidden_size=768, intermediate_size=3072, num_hidden_layers=2,
                      num_attention_heads=12, num_key_value_heads=12, max_position_embeddings=2048)
    prompt = torch.randint(3, 32000, (1, 16), generator=torch.Generator().manual_seed(7))
    torch.manual_seed(0); draft = LlamaForCausalLM(cfg).eval()
    torch.manual_seed(0); target = LlamaForCausalLM(cfg).eval()
    N = 64
    tape = uniform_tape.batch_tape(3, [0], N + 1, GAMMA)
    torch.set_num_threads(1)
    t0 = time.perf_counter()
    ref_tok, ref_d = spec_loop.speculative_sampling(prompt, draft, target, N, GAMMA, 1.0, 20, 0.9, tape=tape[:, 0])
    cpu_s = time.perf_counter() - t0
    dg, tg = draft.cuda(), target.cuda()
    speculative_sampling(prompt.cuda(), dg, tg, None, None, N, GAMMA, 1.0, 20, 0.9, uniforms=tape)            # warm-up + graph capture
    torch.cuda.synchronize(); t0 = time.perf_counter()
    speculative_sampling(prompt.cuda(), dg, tg, None, None, N, GAMMA, 1.0, 20, 0.9, uniforms=tape)
    torch.cuda.synchronize(); gpu_s = time.perf_counter() - t0
    # (statistics from a second, untimed run: details=True times every 8th iteration eagerly with CUDA events)
    tok, d = speculative_sampling(prompt.cuda(), dg, tg, None, None, N, GAMMA, 1.0, 20, 0.9, uniforms=tape, details=True)
    same = min(ref_tok.shape[1], tok.shape[1]) - 16
    agree = int((ref_tok[0, 16:16 + same] == tok[0, 16:16 + same].cpu()).long().cumprod(0).sum())
    return {"workload": "llama-68m-shape draft + target (identical random weights), gamma=4, batch=1, 64 new tokens, fp32, T=1 top_k=20 top_p=0.9",
            "cpu_oracle_1_thread": {"seconds": cpu_s, "emitted_tokens_per_s": (ref_tok.shape[1] - 16) / cpu_s,
                                    "accepted_tokens_per_s": sum(ref_d["acc_len"]) / cpu_s},
            "b200_drop_in": {"seconds": gpu_s, "emitted_tokens_per_s": (tok.shape[1] - 16) / gpu_s, "accepted_tokens_per_s": sum(d["acc_len"]) / gpu_s,
                             "cuda_graph": d["cuda_graph"]},
            "leading_tokens_identical_to_cpu_run": agree, "of": same}


def pin_to_gpu_numa_node(local: int):
    """The e2e leg streams 73.7 MB of pinned host memory per step to the GPU; with N ranks on one box the pinned buffers of
    a rank should live on the NUMA node its GPU hangs off (first touch), or all ranks pull through one socket's memory
    controllers.  Restricts this process to the CPUs local to its GPU (sysfs `local_cpulist` of the PCI device); returns
    a short description for the JSON line, None if the topology is not visible."""
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev_id = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0"
        cpus = open(os.path.join(path, "local_cpulist")).read().strip()
        node = open(os.path.join(path, "numa_node")).read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
            return {"numa_node": int(node), "cpus": cpus, "bound": len(ids)}
    except (OSError, ValueError, AttributeError, RuntimeError):
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-reports", action="store_true", help="skip gpu_aten_baseline / config1 / dense-verify side reports")
    ap.add_argument("--workload", default="config2", choices=["config2", "config1", "variants"],
                    help="config2 = headline synthetic verify microbench; config1 = side report on the reference's CPU-runnable case")
    ap.add_argument("--fused", type=int, default=0,
                    help="0: sd_norm_sample + sd_verify (two launches per step); "
                         "1: sd_norm_sample_verify (one launch per step, requests verified inside the cluster-pipeline norm kernel)")
    ap.add_argument("--pipeline", type=int, default=2,
                    help="1: software-pipelined steps (kernel 2 of batch i runs on a second stream while kernel 1 of batch i+1 streams); "
                         "2 (default): kernel 1 of consecutive steps additionally alternates between two streams, so one launch's drain "
                         "overlaps the next launch's ramp; 0: the two kernels of every step strictly one after the other")
    ap.add_argument("--pdl", type=int, default=int(os.environ.get("SD_PDL", "1")), help="programmatic dependent launch on/off")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "config1":
        return run_config1(args)
    if args.workload == "variants":
        return run_variants(args)

    import torch.distributed as dist
    from llmspeculativesampling_b200 import build, ops
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (the B200 arm has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # (N > 1 only: at N = 1 the CPU baseline leg of this process must keep every host core)
    numa = pin_to_gpu_numa_node(local) if world > 1 else None  # before any pinned allocation: first touch decides the node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    build.build()
    ops.set_pdl(bool(args.pdl))
    clocks = ClockSampler(local)
    clocks.start()                                          # (nvidia-smi needs a moment to come up: started before the set-up)

    B, g = BATCH, GAMMA
    R = 2 * g + 1                                           # rows per request: gamma draft + gamma+1 target
    n_sets = 4                                              # 4 x (73.7 MB logits + 73.7 MB probs) >> 2 x 126 MB L2
    # every rank runs the SAME synthetic sets (weak scaling: the per-GPU work is identical by construction)
    logits, probs, u_rows, u_acc, u_fin, tok_rows, cmp_rows, n_acc, next_tok = [], [], [], [], [], [], [], [], []
    for i in range(n_sets):
        d, t, u = synth_set(i, dev)
        logits.append(torch.cat([d, t], dim=1).contiguous())            # (B, 2g+1, V)
        probs.append(torch.empty(B, R, V, device=dev))
        ur = torch.full((B, R), -1.0, device=dev)
        ur[:, :g] = u[:, :g]                                            # draft rows sample, target rows do not
        u_rows.append(ur.view(-1).contiguous())
        u_acc.append(u[:, g + 1:2 * g + 1].contiguous())
        u_fin.append(u[:, 2 * g + 1].contiguous())
        tok_rows.append(torch.zeros(B, R, dtype=torch.int64, device=dev))
        cmp_rows.append(ops.CompactRows(B * R, dev))                    # compact top-k lists written by kernel 1, read by kernel 2
        n_acc.append(torch.zeros(B, dtype=torch.int32, device=dev))
        next_tok.append(torch.zeros(B, dtype=torch.int64, device=dev))
        del d, t
    acc_total = torch.zeros(2, dtype=torch.int64, device=dev)     # [accepted tokens, requests verified], updated by kernel 2
    err = ops.ErrFlag(dev)
    req_cnt = torch.zeros(B, dtype=torch.int32, device=dev)       # fused launch: finished-row counters (left zeroed)

    err_b = ops.ErrFlag(dev)                                      # second error flag + scheduler workspace (second kernel-1 stream)

    def norm(i: int, lg=None, ur=None, pr=None, ef=None):
        ops.norm_sample((logits[i] if lg is None else lg).view(B * R, V), TEMP, TOP_K, TOP_P, u_rows[i] if ur is None else ur,
                        probs_out=(probs[i] if pr is None else pr).view(B * R, V), tok_out=tok_rows[i].view(-1),
                        err=err if ef is None else ef, compact=cmp_rows[i].view())

    def verify(i: int, count: bool = True, pr=None, ua=None, uf=None):
        p_ = probs[i] if pr is None else pr
        ops.verify(p_[:, g:], p_[:, :g], tok_rows[i][:, :g], u_acc[i] if ua is None else ua, u_fin[i] if uf is None else uf,
                   n_accepted=n_acc[i], next_tok=next_tok[i], err=err, p_compact=cmp_rows[i].view(g, 1), p_cmp_req_stride=R,
                   q_compact=cmp_rows[i].view(0, 1), q_cmp_req_stride=R, stats=acc_total if count else None)

    def step(i: int, count: bool = True):
        if args.fused:
            ops.norm_sample_verify(logits[i].view(B * R, V), TEMP, TOP_K, TOP_P, u_rows[i], probs[i].view(B * R, V),
                                   tok_rows[i].view(-1), cmp_rows[i].view(), R, req_cnt, probs[i][:, g:], probs[i][:, :g],
                                   tok_rows[i][:, :g], u_acc[i], u_fin[i], n_acc[i], next_tok[i], cmp_rows[i].view(g, 1), R,
                                   cmp_rows[i].view(0, 1), R, err, stats=acc_total if count else None)
            return
        norm(i)
        verify(i, count)

    for i in range(n_sets):
        step(i)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side2 = torch.cuda.Stream()

    def capture(fn):
        gr = torch.cuda.CUDAGraph()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(gr, stream=side):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        return gr

    g_serial = [capture(lambda i=i: step(i)) for i in range(n_sets)]       # one step, its kernels strictly in order

    side3 = torch.cuda.Stream()

    def pipelined(n_steps: int = PIPE_STEPS):
        """n_steps steps (PIPE_STEPS by default), independent batches in flight.  --pipeline 1: kernel 1 of step j+1 is launched right behind
        kernel 1 of step j on one stream, kernel 2 of step j runs beside it on a second stream (it needs a few KB of
        compact lists and a handful of thread blocks).  --pipeline 2 (default): kernel 1 of consecutive steps additionally
        alternates between TWO streams (each with its own scheduler workspace): the persistent CTAs of step j+1 take over
        every SM as soon as step j's CTA on it has exited, so the drain of one launch (selection latency of the last rows,
        slowest CTA, grid completion) overlaps the ramp of the next instead of leaving HBM idle.
        A step's buffers are reused n_sets steps later: kernel 1 of step j waits for kernel 2 of step j - n_sets."""
        main_s = torch.cuda.current_stream()
        two = args.pipeline >= 2
        if two:
            side3.wait_stream(main_s)
        ev_v = {}
        for j in range(n_steps):
            i = j % n_sets
            s_n = side3 if (two and j % 2 == 1) else main_s
            if j - n_sets in ev_v:
                s_n.wait_event(ev_v[j - n_sets])
            with torch.cuda.stream(s_n):
                norm(i, ef=err_b if (two and j % 2 == 1) else None)
                ev_n = torch.cuda.Event()
                ev_n.record(s_n)
            if args.pipeline >= 3:                                  # two independent serial chains: kernel 2 stays on its kernel 1's stream
                with torch.cuda.stream(s_n):
                    verify(i)
                    ev_v[j] = torch.cuda.Event()
                    ev_v[j].record(s_n)
                continue
            side2.wait_event(ev_n)
            with torch.cuda.stream(side2):
                verify(i)
                ev_v[j] = torch.cuda.Event()
                ev_v[j].record(side2)
        for j in range(max(0, n_steps - n_sets), n_steps):
            main_s.wait_event(ev_v[j])
        if two:
            main_s.wait_stream(side3)

    use_pipe = bool(args.pipeline) and not args.fused
    g_pipe = capture(pipelined) if use_pipe else None
    # the steps that do not fill a PIPE_STEPS graph run from a shorter pipelined graph (PIPE_STEPS is a multiple of n_sets, so
    # its first step continues the rotation of the input sets), not one by one
    n_rem = args.steps % PIPE_STEPS if use_pipe else 0
    g_pipe_rem = capture(lambda: pipelined(n_rem)) if n_rem >= 2 else None
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(n: int, first: int = 0):
        s = 0
        if g_pipe is not None:
            while n - s >= PIPE_STEPS:
                g_pipe.replay()
                s += PIPE_STEPS
            if g_pipe_rem is not None and n - s == n_rem:
                g_pipe_rem.replay()
                s += n_rem
        while s < n:
            g_serial[(first + s) % n_sets].replay()
            s += 1

    run_steps(max(args.warmup, PIPE_STEPS if use_pipe else 0))
    if g_pipe_rem is not None:
        g_pipe_rem.replay()                                   # (warm-up of the remainder graph as well)
    barrier()
    acc_total.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    run_steps(args.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    err.check()
    accepted = int(acc_total[0].item())
    assert int(acc_total[1].item()) == args.steps * B, "kernel 2 must have verified every request of every step"
    pipelined_steps = ((args.steps // PIPE_STEPS) * PIPE_STEPS + (n_rem if g_pipe_rem is not None else 0)) if use_pipe else 0

    # ---- sub-measurements (self-sized: >= 200 launches whatever --steps is), all from CUDA graphs replayed back to back:
    #      the strictly serial step (latency of one batch), kernel 1 alone, kernel 2 alone
    def timed(graph, launches_per_replay, min_launches=240):
        reps = max(3, (min_launches + launches_per_replay - 1) // launches_per_replay)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            graph.replay()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / (reps * launches_per_replay)

    serial_ms = timed(capture(lambda: [step(i, False) for i in range(n_sets)]), n_sets)
    norm_ms = timed(capture(lambda: [norm(i) for i in range(n_sets)]), n_sets)

    def norm_two_streams():                                 # kernel 1 only, launches alternating between two streams
        main_s = torch.cuda.current_stream()
        side3.wait_stream(main_s)
        for j in range(2 * n_sets):
            with torch.cuda.stream(side3 if j % 2 else main_s):
                norm(j % n_sets, ef=err_b if j % 2 else None)
        main_s.wait_stream(side3)
    norm_overlap_ms = timed(capture(norm_two_streams), 2 * n_sets)
    verify_ms = timed(capture(lambda: [verify(i, False) for i in range(n_sets)]), n_sets)
    clock_info = clocks.stop()
    norm_bytes = B * R * V * 8                              # logits read once (4 B) + probs written once (4 B)
    achieved = norm_bytes / (norm_ms * 1e-3) / 1e9
    # kernel 2, SURVEY 8(d): per request 2*gamma*4 B of gathers + 2*V*4 B if a rejection occurred, V*4 B if all accepted, 16 B out
    n_all = torch.stack(n_acc).long()
    verify_bytes = float((n_all.numel() * (2 * g * 4 + 16) + int((n_all < g).sum()) * 2 * V * 4 + int((n_all >= g).sum()) * V * 4) / n_sets)

    # ---- end to end: host buffers; the H2D copy of step s+1 runs on a copy stream while step s computes (two device
    #      input buffers), accept counts / tokens come back with an event, no device-wide synchronize inside the loop
    host_sets = [(logits[i].cpu().pin_memory(), torch.cat([u_rows[i].view(B, R), u_acc[i], u_fin[i].view(B, 1)], 1).cpu().pin_memory())
                 for i in range(n_sets)]
    nbuf = 2
    l_dev = [torch.empty(B, R, V, device=dev) for _ in range(nbuf)]
    u_dev = [torch.empty(B, R + g + 1, device=dev) for _ in range(nbuf)]
    pr_dev = torch.empty(B, R, V, device=dev)
    h_acc = [torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(nbuf)]
    h_tok = [torch.empty(B, dtype=torch.int64).pin_memory() for _ in range(nbuf)]
    copy_s = torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(nbuf)]
    ev_free = [torch.cuda.Event() for _ in range(nbuf)]
    ev_out = [torch.cuda.Event() for _ in range(nbuf)]
    cur = torch.cuda.current_stream()

    def e2e_upload(s: int):
        b = s % nbuf
        hl, hu = host_sets[s % n_sets]
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(ev_free[b])                   # the step that last read this buffer has finished
            l_dev[b].copy_(hl, non_blocking=True)
            u_dev[b].copy_(hu, non_blocking=True)
            ev_in[b].record(copy_s)

    def e2e_compute(s: int):
        b, i = s % nbuf, s % n_sets
        cur.wait_event(ev_in[b])
        ub = u_dev[b]
        norm(i, lg=l_dev[b], ur=ub[:, :R].reshape(-1), pr=pr_dev)
        verify(i, False, pr=pr_dev, ua=ub[:, R:R + g].contiguous(), uf=ub[:, R + g].contiguous())
        ev_free[b].record(cur)
        h_acc[b].copy_(n_acc[i], non_blocking=True)
        h_tok[b].copy_(next_tok[i], non_blocking=True)
        ev_out[b].record(cur)

    def e2e_run(n: int) -> int:
        got = 0
        e2e_upload(0)
        for s in range(n):
            if s + 1 < n:
                e2e_upload(s + 1)
            e2e_compute(s)
            if s >= 1:                                      # the caller reads step s-1's tokens while step s runs
                ev_out[(s - 1) % nbuf].synchronize()
                got += int(h_acc[(s - 1) % nbuf].sum())
        ev_out[(n - 1) % nbuf].synchronize()
        return got + int(h_acc[(n - 1) % nbuf].sum())

    for ev in ev_free:
        ev.record(cur)
    e2e_steps = 60                                          # (own, fixed length: ~80 ms; a short --steps would mostly time the fill of the two-buffer pipeline)
    e2e_run(4)
    barrier()
    t0 = time.perf_counter()
    e2e_acc = e2e_run(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    h2d = (B * R * V + B * (R + g + 1)) * 4
    d2h = B * (4 + 8)

    # ---- reduce over ranks (max time, summed tokens)
    stats = torch.tensor([ms, e2e_s, float(accepted), float(e2e_acc)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats[:2].clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats[2:].clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_s, accepted, e2e_acc = float(mx[0]), float(mx[1]), float(sm[0]), float(sm[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "norm_traffic.json"))).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    secs = ms * 1e-3
    iters_total = args.steps * B * world
    step_ms = ms / args.steps
    step_bytes = norm_bytes + verify_bytes
    line = {
        "metric": "accepted_tokens_per_s", "value": accepted / secs, "unit": "tokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "l2_policy": f"inputs and outputs rotate over {n_sets} sets "
                   f"({n_sets * norm_bytes / 1e6:.0f} MB > 2 x 126 MB L2), identical on every rank", "cuda_graph": True,
                   "pipeline": (f"software-pipelined (--pipeline {args.pipeline}): {pipelined_steps} of the {args.steps} timed steps ran from a "
                                f"{PIPE_STEPS}-step CUDA graph in which kernel 2 of batch i runs on a second stream beside kernel 1 of batch i+1"
                                + (", and kernel 1 of consecutive batches alternates between two streams (the next launch's persistent CTAs "
                                   "take over each SM as the previous launch's CTA on it exits)" if args.pipeline >= 2 else "")
                                + " — independent batches in flight, results identical; the rest strictly serial") if use_pipe else "strictly serial steps",
                   "kernels_per_step": (["sd_norm_sample_verify (ONE launch: kernel 1 over the B*(2*gamma+1) rows — dense probs + "
                                         "compact lists — and kernel 2's verify of each request inside it)"] if args.fused else
                                        ["sd_norm_sample (ring kernel: B*(2*gamma+1) rows, one launch; dense probs + compact lists)",
                                         "sd_verify (sparse path on the compact lists)"])},
        "emitted_tokens_per_s": (accepted + iters_total) / secs,
        "mean_accepted_per_iteration": accepted / iters_total,
        "request_iterations_per_s": iters_total / secs,
        "serial_ms_per_step": serial_ms,
        "clocks": clock_info,
        "gpu_launches": (1 if args.fused else 2) * args.steps,
        "e2e": {"value": e2e_acc / e2e_s, "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3, "h2d_GBs_per_gpu": h2d / (e2e_s / e2e_steps) / 1e9,
                "host_affinity": numa,
                "note": "double-buffered pinned-host -> device copies on a copy stream, results read back with events; "
                        "bound by the host link (73.7 MB of fp32 logits per step)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "norm_ring_kernel<float, top-k> (1 launch per step, 576 rows)",
                     "algorithmic_bytes_per_step": norm_bytes, "ms_per_step": norm_ms,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                     "frac_of_nominal_8TBs": achieved / 8000.0,
                     "timing": "CUDA events around back-to-back graph replays of the launch (4 rotating input sets, >= 240 launches)",
                     "kernel_ms_graph_back_to_back": {"norm": norm_ms, "verify": verify_ms},
                     "overlapped": {"ms_per_launch": norm_overlap_ms, "achieved": norm_bytes / (norm_overlap_ms * 1e-3) / 1e9,
                                    "frac": norm_bytes / (norm_overlap_ms * 1e-3) / 1e9 / peak,
                                    "note": "the same launches alternating between two streams (independent batches): time per launch when one "
                                            "launch's drain overlaps the next launch's ramp; `achieved` / `frac` above are the strict "
                                            "single-stream figures"},
                     "step": {"algorithmic_bytes": step_bytes, "ms": step_ms, "achieved": step_bytes / (step_ms * 1e-3) / 1e9,
                              "frac": step_bytes / (step_ms * 1e-3) / 1e9 / peak, "serial_ms": serial_ms,
                              "serial_frac": step_bytes / (serial_ms * 1e-3) / 1e9 / peak,
                              "note": "whole step (kernel 1 + kernel 2, SURVEY 8d bytes from the actual accept counts) over the timed ms_per_step"}},
    }
    side_reports = not args.no_side_reports and world == 1    # (side reports and the CPU baseline: rank 0 at N = 1 only)
    if side_reports:
        try:
            line["roofline"]["verify_dense"] = dense_verify_report(dev, ops)
        except Exception as e:                                # noqa: BLE001
            line["roofline"]["verify_dense"] = {"error": f"{type(e).__name__}: {e}"[:200]}
        try:
            line["gpu_aten_baseline"] = gpu_aten_baseline(logits[0], dev)
        except Exception as e:                                # noqa: BLE001
            line["gpu_aten_baseline"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    if not args.no_cpu_baseline and world == 1:
        per = 16
        cores = os.cpu_count() or 1
        a1, _, dt1 = time_oracle(1, 1, per, 1)
        best_threads, best_rate = 1, a1 / dt1
        if cores > 1:
            a2, _, dt2 = time_oracle(1, 1, per, cores)
            if a2 / dt2 > best_rate:
                best_threads, best_rate = cores, a2 / dt2
        n_steps = 60 if best_threads > 1 else 24          # ~10 s of CPU work
        acc_c, _, dt_c = time_oracle(n_steps, 0, per, best_threads)
        line["cpu_baseline"] = {"value": acc_c / dt_c, "unit": "tokens/s", "cores": best_threads, "kind": "port",
                                "sample": f"{n_steps} steps x {per} requests of the same workload ({n_steps * per * (2 * GAMMA + 1)} rows "
                                          f"+ verify) through oracle/ref_ops.py (the reference's ATen op chain, row by row), "
                                          f"torch threads={best_threads} of {cores} host cores; {dt_c:.1f} s of CPU work"}
        if side_reports:
            try:
                line["config1"] = config1_side_report()
            except Exception as e:                            # noqa: BLE001
                line["config1"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def dense_verify_report(dev, ops):
    """Kernel 2's dense path (no compact lists: top_k = 0, the reference API's default) at B = 64, gamma = 4: achieved HBM
    GB/s on the SURVEY 8(d) bytes computed from the actual accept counts, CUDA-graph replays back to back."""
    out = {}
    B, g = BATCH, GAMMA
    for V_ in (32000, 50272):
        n_sets = 3
        sets = []
        for s in range(n_sets):
            gen = torch.Generator(device=dev).manual_seed(77 + s)
            z = 3.0 * torch.randn(B, g + 1, V_, generator=gen, device=dev)
            tl = z + 0.5 * torch.randn(B, g + 1, V_, generator=gen, device=dev)
            dl = z[:, :g] + 0.5 * torch.randn(B, g, V_, generator=gen, device=dev)
            u = torch.rand(B, 2 * g + 2, generator=gen, device=dev)
            q = torch.empty(B, g, V_, device=dev)
            p = torch.empty(B, g + 1, V_, device=dev)
            e = ops.ErrFlag(dev)
            tok = ops.norm_sample(dl.view(B * g, V_), 1.0, 0, 0.0, u[:, :g].contiguous().view(-1), probs_out=q.view(B * g, V_), err=e).view(B, g)
            ops.norm_probs(tl.view(B * (g + 1), V_), 1.0, 0, 0.0, out=p.view(B * (g + 1), V_), err=e)
            sets.append((p, q, tok, u[:, g + 1:2 * g + 1].contiguous(), u[:, 2 * g + 1].contiguous(), torch.zeros(B, dtype=torch.int32, device=dev)))
            del z, tl, dl
        nxt = torch.zeros(B, dtype=torch.int64, device=dev)
        e = ops.ErrFlag(dev)

        def launch_all():
            for (p, q, tok, ua, uf, na) in sets:
                ops.verify(p, q, tok, ua, uf, n_accepted=na, next_tok=nxt, err=e)
        launch_all()
        torch.cuda.synchronize()
        st = torch.cuda.Stream()
        gr = torch.cuda.CUDAGraph()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            with torch.cuda.graph(gr, stream=st):
                launch_all()
        torch.cuda.current_stream().wait_stream(st)
        for _ in range(3):
            gr.replay()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 80
        t0.record()
        for _ in range(reps):
            gr.replay()
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / (reps * n_sets)
        na = torch.stack([s[5] for s in sets]).long()
        nbytes = (na.numel() * (2 * g * 4 + 16) + int((na < g).sum()) * 2 * V_ * 4 + int((na >= g).sum()) * V_ * 4) / n_sets
        out[f"V{V_}"] = {"ms": ms, "algorithmic_bytes": nbytes, "GBs": nbytes / ms / 1e6, "mean_accepted": float(na.float().mean())}
        del sets
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    sys.exit(main())
