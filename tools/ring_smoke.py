"""Quick correctness + timing probe of the ring kernel (run under a short timeout before the full test suite)."""
import sys, os, json, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmspeculativesampling_b200 import ops, build
from oracle import ref_ops
build.build()
torch.manual_seed(0)
for (V, rows, dt, T, k, p) in [(32000, 7, torch.float32, 0.8, 20, 0.9), (32000, 300, torch.float32, 0.8, 20, 0.9), (50272, 200, torch.bfloat16, 0.8, 20, 0.9),
                               (32000, 7, torch.float32, 1.0, 0, 0.0), (32000, 300, torch.float32, 0.8, 0, 0.0), (50272, 200, torch.bfloat16, 1.0, 0, 0.0)]:
    x = (torch.randn(rows, V) * 3.0).to(dt).cuda()
    u = torch.rand(rows).cuda()
    pa = torch.empty(rows, V, device="cuda"); pb = torch.empty(rows, V, device="cuda")
    t0 = time.time()
    ta = ops.norm_sample(x, T, k, p, u, probs_out=pa)
    torch.cuda.synchronize()
    tb = ops.norm_sample(x, T, k, p, u, probs_out=pb, pipeline=False)
    torch.cuda.synchronize()
    ops.default_flag("cuda").check()
    err = ((pa - pb).abs() / pb.clamp_min(1e-30)).max().item()
    sup = bool(((pa > 0) == (pb > 0)).all())
    pc = pa.cpu()
    tok_ok = all(int(ta[i]) == ref_ops.icdf_sample(pc[i], float(u[i])) for i in range(min(rows, 40)))
    print(json.dumps(dict(V=V, rows=rows, dtype=str(dt), k=k, max_rel_err_vs_classic=err, same_support=sup, tokens_match_icdf=tok_ok,
                          tokens_equal_classic=bool(torch.equal(ta, tb)), secs=round(time.time() - t0, 2))), flush=True)
