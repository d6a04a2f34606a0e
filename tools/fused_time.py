"""Timing driver: the bench step (B=64, gamma=4, V=32000 fp32) as one fused launch vs norm + verify, CUDA-graph replays."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmspeculativesampling_b200 import ops, build

build.build()
B, g, V = 64, 4, 32000
R = 2 * g + 1
T, k, p = 0.8, 20, 0.9
n_sets = 4
gen = torch.Generator(device="cuda").manual_seed(1)
sets = []
for i in range(n_sets):
    z = 3.0 * torch.randn(B, g + 1, V, device="cuda", generator=gen)
    d = z[:, :g] + 0.5 * torch.randn(B, g, V, device="cuda", generator=gen)
    t = z + 0.5 * torch.randn(B, g + 1, V, device="cuda", generator=gen)
    u = torch.rand(B, 2 * g + 2, device="cuda", generator=gen)
    ur = torch.full((B, R), -1.0, device="cuda"); ur[:, :g] = u[:, :g]
    sets.append((torch.cat([d, t], 1).contiguous(), torch.empty(B, R, V, device="cuda"), ur.view(-1).contiguous(),
                 u[:, g + 1:2 * g + 1].contiguous(), u[:, 2 * g + 1].contiguous()))
tok = torch.zeros(B, R, dtype=torch.int64, device="cuda")
cmp_rows = ops.CompactRows(B * R, "cuda")
c_all, c_q, c_p = cmp_rows.view(), cmp_rows.view(0, 1), cmp_rows.view(g, 1)
n_acc = torch.zeros(B, dtype=torch.int32, device="cuda"); nxt = torch.zeros(B, dtype=torch.int64, device="cuda")
cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
err = ops.ErrFlag("cuda")

def step(i, mode):
    lg, pr, ur, ua, uf = sets[i]
    if mode == "fused":
        ops.norm_sample_verify(lg.view(B * R, V), T, k, p, ur, pr.view(B * R, V), tok.view(-1), c_all, R, cnt, pr[:, g:], pr[:, :g],
                               tok[:, :g], ua, uf, n_acc, nxt, c_p, R, c_q, R, err)
    else:
        ops.norm_sample(lg.view(B * R, V), T, k, p, ur, probs_out=pr.view(B * R, V), tok_out=tok.view(-1), err=err, compact=c_all)
        if mode == "two":
            ops.verify(pr[:, g:], pr[:, :g], tok[:, :g], ua, uf, n_accepted=n_acc, next_tok=nxt, err=err, p_compact=c_p,
                       p_cmp_req_stride=R, q_compact=c_q, q_cmp_req_stride=R)

if os.environ.get("PROF"):
    from llmspeculativesampling_b200 import _cabi
    buf = torch.zeros(148 * 32 + 64, 16, dtype=torch.int64, device="cuda")
    step(0, "fused"); step(1, "fused"); torch.cuda.synchronize()
    _cabi.load().sd_debug_set_prof(buf.data_ptr())
    step(2, "fused"); torch.cuda.synchronize()
    _cabi.load().sd_debug_set_prof(None)
    t = buf.cpu()[:148 * 32].view(148, 32, 16)
    dur = t[:, 31, 0:4].flatten(); dur = dur[dur != 0].double()
    ent = t[:, 0, 15]; ext = t[:, 1, 15]
    vend = t[:, 31, 4:8].max(dim=1).values
    last_item = t[:, :31, 11].max(dim=1).values
    print(f"in-kernel verifies: {dur.numel()}, duration mean {dur.mean() / 1000:.1f} kcyc, max {dur.max() / 1000:.1f}; "
          f"CTA lifetime mean {(ext - ent).double().mean() / 1000:.1f} kcyc, max {(ext - ent).double().max() / 1000:.1f}; "
          f"last item done -> exit: max {((ext - last_item).double() / 1000).max():.1f} kcyc; "
          f"verify end -> exit: min {((ext - vend)[vend != 0].double() / 1000).min():.1f} kcyc")
    sys.exit(0)
for mode in sys.argv[1:] or ["norm", "two", "fused"]:
    for i in range(n_sets):
        step(i, mode)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(gr, stream=side):
            for i in range(n_sets):
                step(i, mode)
    torch.cuda.synchronize()
    for _ in range(5):
        gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    print(mode, "us per step:", round(e0.elapsed_time(e1) / (200 * n_sets) * 1000, 2), "acc", int(n_acc.sum()), flush=True)
