"""GPU parity: kernel 2 (verify), sample, max_fn, KV append and the step builder."""
import numpy as np
import pytest
import torch

from oracle import ref_ops
from tests.helpers import make_logits, oracle_probs

pytestmark = pytest.mark.gpu


def _case(B, gamma, V, seed, k=20, p=0.9, T=0.8, noise=0.5):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(B, gamma + 1, V, generator=g) * 3.0
    tl = z + noise * torch.randn(B, gamma + 1, V, generator=g)
    dl = z[:, :gamma] + noise * torch.randn(B, gamma, V, generator=g)
    pp = oracle_probs(tl.reshape(-1, V), T, k, p).reshape(B, gamma + 1, V)
    qq = oracle_probs(dl.reshape(-1, V), T, k, p).reshape(B, gamma, V)
    u = torch.rand(B, 2 * gamma + 2, generator=g)
    draft = torch.tensor([[ref_ops.icdf_sample(qq[b, i], float(u[b, i])) for i in range(gamma)] for b in range(B)])
    return pp, qq, draft, u[:, gamma + 1:2 * gamma + 1].contiguous(), u[:, 2 * gamma + 1].contiguous()


@pytest.mark.parametrize("B,gamma,V,k,p", [(16, 4, 32000, 20, 0.9), (8, 4, 50272, 0, 0.0), (12, 1, 1000, 20, 0.9),
                                           (6, 8, 4099, 5, 0.0), (4, 16, 2048, 0, 0.9), (3, 4, 262144, 20, 0.9),
                                           # up to SMs / 2 requests: two CTAs per request (cluster of 2); more: one CTA per request
                                           (74, 4, 32000, 0, 0.0), (80, 3, 32000, 0, 0.0), (5, 4, 9000, 0, 0.0)])
@pytest.mark.parametrize("strict", [False, True])
def test_verify_bit_exact(cuda_lib, B, gamma, V, k, p, strict):
    from llmspeculativesampling_b200 import ops
    pp, qq, draft, u_acc, u_fin = _case(B, gamma, V, seed=B * 7 + gamma, k=k, p=p)
    ratios = torch.zeros(B, gamma, device="cuda")
    ties = torch.zeros(1, dtype=torch.int32, device="cuda")
    n_acc, tok = ops.verify(pp.cuda(), qq.cuda(), draft.cuda(), u_acc.cuda(), u_fin.cuda(), strict=strict,
                            ratios=ratios, tie_count=ties)
    ops.default_flag("cuda").check()
    for b in range(B):
        wn, wt, wr, _ = ref_ops.verify_request(pp[b], qq[b], draft[b], u_acc[b].numpy(), float(u_fin[b]), strict=strict)
        assert int(n_acc[b]) == wn and int(tok[b]) == wt, f"request {b}: got ({int(n_acc[b])},{int(tok[b])}) want ({wn},{wt})"
        assert np.array_equal(ratios[b].cpu().numpy(), wr)
    assert 0 < float(n_acc.float().mean()) or k == 0


def test_verify_all_accept_all_reject_and_fallback(cuda_lib):
    from llmspeculativesampling_b200 import ops
    B, gamma, V = 6, 4, 32000
    pp, qq, draft, u_acc, u_fin = _case(B, gamma, V, seed=5)
    # request 0/1: identical distributions -> every token accepted (ratio 1, u < 1), bonus from row gamma
    qq[0] = pp[0, :gamma]; qq[1] = pp[1, :gamma]
    draft[0] = torch.tensor([int(pp[0, i].argmax()) for i in range(gamma)])
    draft[1] = torch.tensor([int(pp[1, i].argmax()) for i in range(gamma)])
    # request 2: q >= p everywhere at the rejected row -> empty residual -> falls back to sample(p_n)
    # (reference speculative_sampling.py:2009-2010)
    lo = int((pp[2, 0] > 0).nonzero()[-1])
    qq[2, 0] = pp[2, 0]
    qq[2, 0, lo] = pp[2, 0, lo] * 4
    draft[2, 0] = lo
    u_acc[2, 0] = 0.999999
    # request 3: drafted token has zero target probability -> rejected at position 0
    zero_tok = int((pp[3, 0] == 0).nonzero()[0])
    qq[3, 0, zero_tok] = 0.25
    draft[3, 0] = zero_tok
    u_acc[3, 0] = 0.5
    tokens = torch.zeros(B, 64, dtype=torch.int64, device="cuda")
    seq = torch.full((B,), 10, dtype=torch.int32, device="cuda")
    tokens[:, 10:10 + gamma] = draft.cuda()
    n_acc, tok = ops.verify(pp.cuda(), qq.cuda(), draft.cuda(), u_acc.cuda(), u_fin.cuda(), tokens=tokens, seq_len=seq)
    ops.default_flag("cuda").check()
    for b in range(B):
        wn, wt, _, _ = ref_ops.verify_request(pp[b], qq[b], draft[b], u_acc[b].numpy(), float(u_fin[b]))
        assert (int(n_acc[b]), int(tok[b])) == (wn, wt), b
        assert int(seq[b]) == 10 + wn + 1
        assert int(tokens[b, 10 + wn]) == wt
        assert tokens[b, 10:10 + wn].cpu().tolist() == draft[b, :wn].tolist()
    assert int(n_acc[0]) == gamma and int(n_acc[1]) == gamma and int(n_acc[3]) == 0


def test_verify_active_mask_and_errors(cuda_lib):
    from llmspeculativesampling_b200 import ops
    B, gamma, V = 4, 4, 2048
    pp, qq, draft, u_acc, u_fin = _case(B, gamma, V, seed=9)
    active = torch.tensor([1, 0, 1, 0], dtype=torch.int32, device="cuda")
    n_acc = torch.full((B,), -7, dtype=torch.int32, device="cuda")
    tok = torch.full((B,), -7, dtype=torch.int64, device="cuda")
    ops.verify(pp.cuda(), qq.cuda(), draft.cuda(), u_acc.cuda(), u_fin.cuda(), n_accepted=n_acc, next_tok=tok, active=active)
    ops.default_flag("cuda").check()
    assert int(n_acc[1]) == -7 and int(tok[3]) == -7 and int(n_acc[0]) >= 0 and int(tok[2]) >= 0
    qq2 = qq.clone()
    qq2[0, 0, int(draft[0, 0])] = 0.0                        # ZeroDivisionError in the reference -> RuntimeError('s')
    ops.verify(pp.cuda(), qq2.cuda(), draft.cuda(), u_acc.cuda(), u_fin.cuda())
    with pytest.raises(RuntimeError, match="^s$"):
        ops.default_flag("cuda").check()


@pytest.mark.parametrize("V", [17, 1000, 32000, 50257, 262144])
def test_sample_rows_bit_exact(cuda_lib, V):
    from llmspeculativesampling_b200 import ops
    rows = 40
    g = torch.Generator().manual_seed(V)
    probs = torch.rand(rows, V, generator=g) ** 8
    if V >= 1000:
        probs[::3] *= (torch.rand(rows, V, generator=g) > 0.99)[::3]     # sparse rows
    probs[1] = 0; probs[1, V // 2] = 3.5                             # single-support, un-normalised
    probs = probs.float()
    u = torch.rand(rows, generator=g)
    u[2] = 0.0; u[3] = 1.0 - 2 ** -24
    tok = ops.sample_rows(probs.cuda(), u.cuda())
    ops.default_flag("cuda").check()
    want = [ref_ops.icdf_sample(probs[i], float(u[i])) for i in range(rows)]
    assert tok.cpu().tolist() == want
    z = torch.zeros(2, V)
    ops.sample_rows(z.cuda(), u[:2].cuda())
    with pytest.raises(RuntimeError, match="prob error"):
        ops.default_flag("cuda").check()


def test_max_fn_matches_reference_golden(cuda_lib):
    import os
    from llmspeculativesampling_b200 import ops
    blob = np.load(os.path.join(os.path.dirname(__file__), "golden", "max_fn.npz"))
    x, y = torch.from_numpy(blob["x"]), torch.from_numpy(blob["y"])
    got = ops.max_fn(x.cuda()).cpu()
    assert torch.allclose(got, y, rtol=1e-5, atol=1e-12)
    assert float(got[5].abs().sum()) == 0.0
    assert torch.allclose(ops.max_fn(x[0].cuda()).cpu(), y[0], rtol=1e-5, atol=1e-12)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_kv_append_and_build_step(cuda_lib, dtype):
    from llmspeculativesampling_b200 import ops
    B, H, S, D, q = 5, 3, 40, 64, 3
    g = torch.Generator().manual_seed(1)
    kc = torch.randn(B, H, S, D, generator=g).to(dtype).cuda(); vc = torch.randn(B, H, S, D, generator=g).to(dtype).cuda()
    kn = torch.randn(B, q, H, D, generator=g).to(dtype).cuda().transpose(1, 2)      # HF layout: strided (B,H,q,D) view
    vn = torch.randn(B, q, H, D, generator=g).to(dtype).cuda().transpose(1, 2)
    tokens = torch.randint(0, 1000, (B, S), generator=g).cuda()
    seq = torch.tensor([5, 9, 2, 30, 17], dtype=torch.int32).cuda()
    prev = torch.tensor([901, 902, 903, 904, 905]).cuda()
    ids = torch.zeros(B, q, dtype=torch.int64).cuda(); pos = torch.zeros(B, q, dtype=torch.int64).cuda()
    wp = torch.zeros(B, dtype=torch.int32).cuda(); mask = torch.zeros(B, 1, q, S, dtype=torch.uint8).cuda()
    want_tokens = tokens.clone()
    ops.build_step(tokens, seq, -2, q, prev, S, ids, pos, wp, mask)
    start = (seq - 2).clamp_min(0).long()
    for b in range(B):
        want_tokens[b, start[b] + q - 1] = prev[b]
        assert ids[b].tolist() == want_tokens[b, start[b]:start[b] + q].tolist()
        assert pos[b].tolist() == list(range(int(start[b]), int(start[b]) + q))
        for j in range(q):
            assert mask[b, 0, j].bool().tolist() == [s <= int(start[b]) + j for s in range(S)]
    assert torch.equal(tokens, want_tokens) and torch.equal(wp.long(), start)
    k_ref, v_ref = kc.clone(), vc.clone()
    for b in range(B):
        k_ref[b, :, start[b]:start[b] + q] = kn[b]
        v_ref[b, :, start[b]:start[b] + q] = vn[b]
    ops.kv_append(kn, vn, kc, vc, wp)
    assert torch.equal(kc, k_ref) and torch.equal(vc, v_ref)


@pytest.mark.parametrize("V,dtype", [(32000, torch.float32), (50272, torch.bfloat16)])
@pytest.mark.parametrize("pipeline", [True, False])
def test_compact_lists_and_sparse_verify_equal_dense(cuda_lib, V, dtype, pipeline):
    """Kernel 1 emits the compact (index, prob) lists next to the dense rows; kernel 2 fed with them (sparse path)
    must return exactly what the dense path returns, including for requests whose lists are unavailable."""
    from llmspeculativesampling_b200 import ops
    B, gamma, T, k, p = 24, 4, 0.8, 20, 0.9
    R = 2 * gamma + 1
    g = torch.Generator().manual_seed(V)
    z = torch.randn(B, gamma + 1, V, generator=g) * 3.0
    tl = (z + 0.5 * torch.randn(B, gamma + 1, V, generator=g)).to(dtype)
    dl = (z[:, :gamma] + 0.5 * torch.randn(B, gamma, V, generator=g)).to(dtype)
    tl[5, 2] = 1.0                                            # all-ties target row: no compact list (-1) -> dense fallback
    dl[7, 0] = dl[7, 0].round()                               # heavy ties in a draft row
    logits = torch.cat([dl, tl], 1).contiguous().cuda()
    u = torch.rand(B, 2 * gamma + 2, generator=g)
    ur = torch.full((B, R), -1.0); ur[:, :gamma] = u[:, :gamma]
    probs = torch.empty(B, R, V, device="cuda")
    tok = torch.zeros(B, R, dtype=torch.int64, device="cuda")
    cmp_rows = ops.CompactRows(B * R, "cuda")
    ops.norm_sample(logits.view(B * R, V), T, k, p, ur.view(-1).cuda(), probs_out=probs.view(B * R, V), tok_out=tok.view(-1),
                    pipeline=pipeline, compact=cmp_rows.view())
    ops.default_flag("cuda").check()
    dense_from_lists = cmp_rows.to_dense(V).view(B, R, V)
    have = ~torch.isnan(dense_from_lists[:, :, 0])
    assert bool(have.float().mean() > 0.9) and not bool(have[5, gamma + 2])
    assert torch.equal(dense_from_lists[have], probs[have])
    u_acc = u[:, gamma + 1:2 * gamma + 1].contiguous().cuda(); u_fin = u[:, 2 * gamma + 1].contiguous().cuda()
    u_acc[5, :] = 0.0                                         # request 5 accepts up to the tie row ...
    u_acc[5, 2] = 1.0 - 2 ** -24                              # ... and rejects there: its residual needs the dense fallback
    a_d, t_d = ops.verify(probs[:, gamma:], probs[:, :gamma], tok[:, :gamma], u_acc, u_fin)
    ratios = torch.zeros(B, gamma, device="cuda")
    a_s, t_s = ops.verify(probs[:, gamma:], probs[:, :gamma], tok[:, :gamma], u_acc, u_fin, ratios=ratios,
                          p_compact=cmp_rows.view(gamma, 1), p_cmp_req_stride=R, q_compact=cmp_rows.view(0, 1), q_cmp_req_stride=R)
    ops.default_flag("cuda").check()
    assert torch.equal(a_d, a_s) and torch.equal(t_d, t_s)
    assert int(a_s[5]) == 2
    pc, qc = probs.cpu(), None
    for b in range(0, B, 5):
        wn, wt, wr, _ = ref_ops.verify_request(pc[b, gamma:], pc[b, :gamma], tok[b, :gamma].cpu(), u_acc[b].cpu().numpy(), float(u_fin[b]))
        assert (int(a_s[b]), int(t_s[b])) == (wn, wt)
        assert np.array_equal(ratios[b].cpu().numpy(), wr)


@pytest.mark.parametrize("V,dtype,B", [(32000, torch.float32, 64), (50272, torch.bfloat16, 40), (32000, torch.bfloat16, 300),
                                       (4096, torch.float32, 2000), (32000, torch.float32, 1200)])
@pytest.mark.parametrize("strict", [False, True])
def test_fused_norm_sample_verify_equals_two_launches(cuda_lib, V, dtype, B, strict):
    """sd_norm_sample_verify (verify inside the persistent norm kernel, by the group that finishes a request's last row)
    must return exactly what sd_norm_sample followed by sd_verify returns: probabilities, drafted tokens, accept counts,
    next tokens, appended tokens / lengths, ratios, statistics — including requests with tied rows (deferred to the
    general path at the end of the kernel) and requests whose residual row has no compact list (dense scan)."""
    from llmspeculativesampling_b200 import ops
    gamma, T, k, p = 4, 0.8, 20, 0.9
    R = 2 * gamma + 1
    g = torch.Generator().manual_seed(V + B)
    z = torch.randn(B, gamma + 1, V, generator=g) * 3.0
    tl = (z + 0.5 * torch.randn(B, gamma + 1, V, generator=g)).to(dtype)
    dl = (z[:, :gamma] + 0.5 * torch.randn(B, gamma, V, generator=g)).to(dtype)
    tl[5, 2] = 1.0                                            # all-ties target row: deferred row, no list, dense residual
    dl[7, 0] = dl[7, 0].round()                               # heavy ties in a draft row
    tl[B - 1, gamma] = 0.5                                    # the very last row of the launch is a deferred one
    for b in range(11, B, 97):
        tl[b, 1] = -2.0
    logits = torch.cat([dl, tl], 1).contiguous().cuda()
    u = torch.rand(B, 2 * gamma + 2, generator=g)
    ur = torch.full((B, R), -1.0); ur[:, :gamma] = u[:, :gamma]
    ur = ur.view(-1).cuda()
    u_acc = u[:, gamma + 1:2 * gamma + 1].contiguous().cuda(); u_fin = u[:, 2 * gamma + 1].contiguous().cuda()
    u_acc[5, :] = 0.0; u_acc[5, 2] = 1.0 - 2 ** -24           # request 5 rejects at its tied row
    S = 16
    res = []
    for fused in (False, True, "two-launch fallback"):     # the last: no persistent kernel -> the library launches norm + verify itself
        probs = torch.empty(B, R, V, device="cuda")
        tok = torch.zeros(B, R, dtype=torch.int64, device="cuda")
        cmp_rows = ops.CompactRows(B * R, "cuda")
        n_acc = torch.full((B,), -7, dtype=torch.int32, device="cuda"); nxt = torch.full((B,), -7, dtype=torch.int64, device="cuda")
        ratios = torch.zeros(B, gamma, device="cuda"); stats = torch.zeros(2, dtype=torch.int64, device="cuda")
        tokens = torch.zeros(B, S, dtype=torch.int64, device="cuda"); seq_len = torch.full((B,), 3, dtype=torch.int32, device="cuda")
        err = ops.ErrFlag("cuda")
        kw = dict(p_compact=cmp_rows.view(gamma, 1), p_cmp_req_stride=R, q_compact=cmp_rows.view(0, 1), q_cmp_req_stride=R)
        for rep in range(2 if fused else 1):                   # twice: the counters must be left re-armed
            if fused:
                cnt = torch.zeros(B, dtype=torch.int32, device="cuda") if rep == 0 else cnt
                seq_len.fill_(3); stats.zero_(); n_acc.fill_(-7); probs.fill_(-1.0)
                ops.norm_sample_verify(logits.view(B * R, V), T, k, p, ur, probs.view(B * R, V), tok.view(-1), cmp_rows.view(), R, cnt,
                                       probs[:, gamma:], probs[:, :gamma], tok[:, :gamma], u_acc, u_fin, n_acc, nxt, err=err,
                                       strict=strict, ratios=ratios, tokens=tokens, seq_len=seq_len, stats=stats,
                                       pipeline=fused is True, **kw)
                assert int(cnt.abs().sum()) == 0
            else:
                ops.norm_sample(logits.view(B * R, V), T, k, p, ur, probs_out=probs.view(B * R, V), tok_out=tok.view(-1), err=err,
                                compact=cmp_rows.view())
                ops.verify(probs[:, gamma:], probs[:, :gamma], tok[:, :gamma], u_acc, u_fin, strict=strict, n_accepted=n_acc,
                           next_tok=nxt, ratios=ratios, tokens=tokens, seq_len=seq_len, err=err, stats=stats, **kw)
        torch.cuda.synchronize()
        err.check()
        res.append((probs, tok, n_acc, nxt, ratios, tokens, seq_len, stats))
    names = ["probs", "tok", "n_acc", "next", "ratios", "tokens", "seq_len", "stats"]
    for other in (1, 2):
        for nm, a, b_ in zip(names, res[0], res[other]):
            assert torch.equal(a, b_), f"{nm} differs between the fused entry point (arm {other}) and norm + verify"
    assert int(res[1][7][1]) == B and int(res[1][2][5]) == 2


@pytest.mark.parametrize("B,W,gamma,V,k,p,noise", [(6, 3, 4, 32000, 20, 0.9, 0.5), (5, 8, 2, 1000, 0, 0.0, 1.0), (4, 2, 5, 4099, 5, 0.0, 0.2),
                                                   (3, 4, 4, 2048, 20, 0.9, 3.0)])
def test_verify_multi_bit_exact(cuda_lib, B, W, gamma, V, k, p, noise):
    """sd_verify_multi (W drafts per request, lazily drawn accept uniforms, first longest run wins) vs the oracle's
    restatement of speculative_sampling.py:1612-1667: winning draft, accepted run, next token and ratios."""
    from llmspeculativesampling_b200 import ops
    g = torch.Generator().manual_seed(B * 100 + W)
    z = torch.randn(B, 1, gamma + 1, V, generator=g) * 3.0
    tl = z + noise * torch.randn(B, W, gamma + 1, V, generator=g)
    dl = z[:, :, :gamma] + noise * torch.randn(B, W, gamma, V, generator=g)
    pp = oracle_probs(tl.reshape(-1, V), 0.9, k, p).reshape(B, W, gamma + 1, V)
    qq = oracle_probs(dl.reshape(-1, V), 0.9, k, p).reshape(B, W, gamma, V)
    u_d = torch.rand(B, W, gamma, generator=g)
    draft = torch.tensor([[[ref_ops.icdf_sample(qq[b, w, i], float(u_d[b, w, i])) for i in range(gamma)] for w in range(W)]
                          for b in range(B)])
    u_seq = torch.rand(B, W * gamma + 3, generator=g)
    u_seq[0, :] = 0.0                                          # request 0: every token with p > 0 is accepted (bonus row)
    if B > 1:
        u_seq[1, :] = 1.0 - 2 ** -24                           # request 1: only ratios >= 1 are accepted
    u_fin = torch.rand(B, generator=g)
    ratios = torch.zeros(B, W, gamma, device="cuda")
    ch, na, nt = ops.verify_multi(pp.cuda(), qq.cuda(), draft.cuda(), u_seq.cuda(), u_fin.cuda(), ratios=ratios)
    ops.default_flag("cuda").check()
    for b in range(B):
        wc, wn, wt, wr = ref_ops.verify_multi_request(pp[b], qq[b], draft[b], u_seq[b].numpy(), float(u_fin[b]))
        assert (int(ch[b]), int(na[b]), int(nt[b])) == (wc, wn, wt), f"request {b}"
        assert np.array_equal(ratios[b].cpu().numpy(), wr, equal_nan=True)


@pytest.mark.parametrize("B,C,V,k,p", [(16, 4, 32000, 20, 0.9), (9, 7, 1000, 0, 0.0), (5, 1, 4099, 5, 0.0)])
def test_verify_bild_bit_exact(cuda_lib, B, C, V, k, p):
    """sd_verify_bild (keep tokens while -log p[token] <= rollback_thres, then sample the target's own token) vs the
    reference rule of speculative_sampling.py:1797-1812 restated with torch on the CPU."""
    from llmspeculativesampling_b200 import ops
    g = torch.Generator().manual_seed(B + C)
    pp = oracle_probs(torch.randn(B * (C + 1), V, generator=g) * 2.0, 1.0, k, p).reshape(B, C + 1, V)
    draft = torch.stack([torch.stack([torch.multinomial(pp[b, i] + 1e-4 / V, 1, generator=g)[0] for i in range(C)]) for b in range(B)])
    draft[0, 0] = int((pp[0, 0] == 0).nonzero()[0]) if bool((pp[0, 0] == 0).any()) else draft[0, 0]   # p = 0: -log = inf fails
    n_check = torch.randint(1, C + 1, (B,), generator=g, dtype=torch.int32)
    u_fin = torch.rand(B, generator=g)
    thr = 3.0
    nll = torch.zeros(B, C, device="cuda")
    kept, nxt = ops.verify_bild(pp.cuda(), draft.cuda(), thr, u_fin.cuda(), n_check=n_check.cuda(), nll=nll)
    ops.default_flag("cuda").check()
    for b in range(B):
        n = int(n_check[b])
        for i in range(int(n_check[b])):
            if float(-pp[b, i, draft[b, i]].log()) > thr:
                n = i
                break
        assert int(kept[b]) == n, f"request {b}"
        assert int(nxt[b]) == ref_ops.icdf_sample(pp[b, n], float(u_fin[b]))
        want = (-pp[b, torch.arange(int(n_check[b])), draft[b, :int(n_check[b])]].log()).numpy()
        assert np.allclose(nll[b, :int(n_check[b])].cpu().numpy(), want, rtol=1e-6, atol=1e-7)


def test_distribution_preservation_chi_square(cuda_lib):
    """The reference authors' manual two-token check (kvcache_model.py:73-76, speculative_sampling.py:227-229: force
    p = {1: .4, 12: .6}, q = {1: .6, 12: .4} and count emitted tokens) as a statistical test of the whole step:
    draft from q (sd_sample), verify against p (sd_verify); the first emitted token must follow p whatever q is."""
    from llmspeculativesampling_b200 import ops
    N, V = 200000, 16
    g = torch.Generator().manual_seed(11)
    for strict in (False, True):
        for (pd, qd) in [({1: 0.4, 12: 0.6}, {1: 0.6, 12: 0.4}), ({0: 0.1, 3: 0.2, 7: 0.3, 15: 0.4}, {0: 0.4, 3: 0.3, 7: 0.2, 9: 0.1})]:
            p_row = torch.zeros(V); q_row = torch.zeros(V)
            for k_, v_ in pd.items(): p_row[k_] = v_
            for k_, v_ in qd.items(): q_row[k_] = v_
            q = q_row.repeat(N, 1, 1).cuda()                      # (N, gamma=1, V)
            p = p_row.repeat(N, 2, 1).cuda()                      # (N, gamma+1, V)
            u = torch.rand(N, 4, generator=g).cuda()
            draft = ops.sample_rows(q[:, 0].contiguous(), u[:, 0].contiguous()).view(N, 1)
            n_acc, nxt = ops.verify(p, q, draft, u[:, 2:3].contiguous(), u[:, 3].contiguous(), strict=strict)
            ops.default_flag("cuda").check()
            first = torch.where(n_acc > 0, draft[:, 0], nxt)      # accepted draft token, else the resampled one
            counts = torch.bincount(first.cpu(), minlength=V).double()
            expect = p_row.double() * N
            support = expect > 0
            assert float(counts[~support].sum()) == 0.0
            chi2 = float(((counts[support] - expect[support]) ** 2 / expect[support]).sum())
            dof = int(support.sum()) - 1
            assert chi2 < 30.0, f"chi2={chi2:.1f} (dof {dof}) strict={strict}: emitted tokens do not follow the target"
            # acceptance rate = sum_i min(p_i, q_i)
            want_acc = float(torch.minimum(p_row, q_row).sum())
            assert abs(float(n_acc.float().mean()) - want_acc) < 0.01
