"""Model-pair benchmarks (BASELINE.json configs 3 and 4): stock Hugging Face modules with random-init weights of
the named shapes behind the batched B200 engine (static KV caches, one CUDA graph per iteration).

    python tools/bench_models.py --pair opt   --batch 32 --new 128            # config 3
    torchrun --nproc-per-node N tools/bench_models.py --pair llama --requests 256 --batch 64   # config 4

Reports emitted / accepted tokens per second, iterations/s, mean accepted length and acceptance, for the reference's
sampling setting (T=1, top_k=20, top_p=0.9: independent random-init pairs accept ~0) and for the full
softmax (k=0, p=0).  `--tie-draft` builds the draft from the target's first layers + its embeddings/head so that the two
distributions correlate (acceptance > 0) without any checkpoint.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build_pair(pair: str, dtype, device, small: bool):
    from transformers import LlamaConfig, LlamaForCausalLM, OPTConfig, OPTForCausalLM
    if pair == "opt":
        dcfg = OPTConfig(vocab_size=50272, hidden_size=768, ffn_dim=3072, num_hidden_layers=12, num_attention_heads=12,
                         max_position_embeddings=2048, word_embed_proj_dim=768)
        tcfg = OPTConfig(vocab_size=50272, hidden_size=5120, ffn_dim=20480, num_hidden_layers=40, num_attention_heads=40,
                         max_position_embeddings=2048, word_embed_proj_dim=5120)
        if small:
            tcfg = OPTConfig(vocab_size=50272, hidden_size=1024, ffn_dim=4096, num_hidden_layers=4, num_attention_heads=16,
                             max_position_embeddings=2048, word_embed_proj_dim=1024)
        mk = OPTForCausalLM
    else:
        dcfg = LlamaConfig(vocab_size=32000, hidden_size=768, intermediate_size=3072, num_hidden_layers=2,
                           num_attention_heads=12, num_key_value_heads=12, max_position_embeddings=2048)
        tcfg = LlamaConfig(vocab_size=32000, hidden_size=5120, intermediate_size=13824, num_hidden_layers=40,
                           num_attention_heads=40, num_key_value_heads=40, max_position_embeddings=4096)
        if small:
            tcfg = LlamaConfig(vocab_size=32000, hidden_size=1024, intermediate_size=2816, num_hidden_layers=4,
                               num_attention_heads=16, num_key_value_heads=16, max_position_embeddings=4096)
        mk = LlamaForCausalLM
    torch.manual_seed(0)
    with torch.device(device):
        torch.set_default_dtype(dtype)
        try:
            draft = mk(dcfg).eval()
            torch.manual_seed(1)
            target = mk(tcfg).eval()
        finally:
            torch.set_default_dtype(torch.float32)
    return draft.to(dtype), target.to(dtype)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pair", default="opt", choices=["opt", "llama"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--requests", type=int, default=0, help="total requests over all ranks (default: batch per rank)")
    ap.add_argument("--prompt", type=int, default=64)
    ap.add_argument("--new", type=int, default=128)
    ap.add_argument("--gamma", type=int, default=4)
    ap.add_argument("--small", action="store_true", help="small target (smoke test of the harness)")
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--settings", default="k20p0.9,full")
    a = ap.parse_args()

    import torch.distributed as dist
    from llmspeculativesampling_b200 import build, uniform_tape
    from llmspeculativesampling_b200.engine import SpecDecEngine
    from llmspeculativesampling_b200.sharding import shard_requests, batches
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    build.build()
    dtype = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[a.dtype]
    t0 = time.time()
    draft, target = build_pair(a.pair, dtype, dev, a.small)
    V = target.config.vocab_size
    init_s = time.time() - t0
    n_req = a.requests or a.batch * world
    mine = shard_requests(n_req, world, rank)
    a.batch = max(1, min(a.batch, len(mine)))                          # strong scaling: a shard smaller than the batch runs as one batch
    results = []
    for setting in a.settings.split(","):
        top_k, top_p = (20, 0.9) if setting == "k20p0.9" else (0, 0.0)
        eng = SpecDecEngine(draft, target, a.batch, a.prompt + a.new, a.gamma, 1.0, top_k, top_p, dev)
        tot_emit = tot_acc = tot_iter_req = 0
        checksum = 0
        elapsed = 0.0
        captured = False
        for bi, ids in enumerate(batches(mine, a.batch)):
            while len(ids) < a.batch:                                  # pad the last batch (padding is not counted)
                ids = ids + [ids[-1]]
            g = torch.Generator().manual_seed(7)
            prompts = [torch.randint(3, V, (a.prompt,), generator=torch.Generator().manual_seed(1000 + r)) for r in ids]
            tape = uniform_tape.batch_tape(5, ids, a.new + 1, a.gamma).to(dev)
            eng.load_prompts(prompts, a.new)
            if bi == 0:                                                # graph capture + warm-up outside the timing
                eng.run(tape[:2])
                eng.load_prompts(prompts, a.new)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            iters = eng.run(tape)
            torch.cuda.synchronize()
            elapsed += time.perf_counter() - t1
            captured = eng.graph_captured
            acc = eng.acc_hist[:iters].cpu()
            live = acc >= 0
            real = len(set(ids))
            tot_acc += int(acc[:, :real][live[:, :real]].sum())
            tot_iter_req += int(live[:, :real].sum())
            tot_emit += int((eng.seq_len - eng.prompt_len)[:real].sum())
            # order-independent checksum of (request id, emitted tokens): equal sums over N = 1, 2, 4, 8 mean the union of
            # the shards reproduced the single-GPU tokens (exact for bit-reproducible logits; GEMMs of a real model may
            # pick different algorithms at different batch sizes)
            toks, lens, plen = eng.tokens[:real].cpu(), eng.seq_len[:real].cpu(), eng.prompt_len[:real].cpu()
            for j in range(real):
                row = toks[j, int(plen[j]):int(lens[j])].tolist()
                h = 1469598103934665603
                for t in [ids[j]] + row:
                    h = ((h ^ int(t)) * 1099511628211) & 0xFFFFFFFFFFFF
                checksum = (checksum + (h & 0xFFFFFFFFFF)) % (1 << 48)      # (< 2^48: the float64 all-reduce below stays exact)
        stats = torch.tensor([elapsed, float(tot_emit), float(tot_acc), float(tot_iter_req), float(checksum)], dtype=torch.float64, device=dev)
        if world > 1:
            mx = stats[:1].clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = stats[1:].clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            stats = torch.cat([mx, sm])
        el, emit, acc_n, itreq, csum = [float(x) for x in stats]
        results.append({"setting": setting, "top_k": top_k, "top_p": top_p, "emitted_tokens_per_s": emit / el,
                        "accepted_tokens_per_s": acc_n / el, "mean_accepted_per_iteration": acc_n / max(itreq, 1),
                        "request_iterations_per_s": itreq / el, "seconds": el, "cuda_graph": captured,
                        "emitted_tokens": int(emit), "tokens_checksum": int(csum) % (1 << 48)})
        del eng
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps({"pair": a.pair, "small_target": a.small, "dtype": a.dtype, "n_gpus": world, "requests": n_req,
                          "batch_per_gpu": a.batch, "prompt": a.prompt, "new_tokens": a.new, "gamma": a.gamma,
                          "model_init_s": round(init_s, 1), "results": results}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
