import sys, torch
sys.path.insert(0, '/root/repo')
from llmspeculativesampling_b200 import ops
torch.manual_seed(0)
for (rows, V, dt) in [(300, 32000, torch.float32), (40, 50272, torch.bfloat16), (9, 4099, torch.float32)]:
    x = (torch.randn(rows, V, device='cuda') * 3.8).to(dt)
    x[3] = 1.0
    u = torch.rand(rows, device='cuda')
    p = torch.empty(rows, V, device='cuda')
    c = ops.CompactRows(rows, 'cuda')
    for pipeline in (True, False):
        ops.norm_sample(x, 0.8, 20, 0.9, u, probs_out=p, pipeline=pipeline, compact=c.view())
        ops.norm_probs(x, 1.0, 0, 0.0, out=p, pipeline=pipeline)
        ops.norm_probs(x[:8], 1.0, 0, 0.9, out=p[:8], pipeline=pipeline)
B, g, V = 8, 4, 32000
pp = torch.softmax(torch.randn(B, g + 1, V, device='cuda') * 3, -1); qq = torch.softmax(torch.randn(B, g, V, device='cuda') * 3, -1)
d = torch.randint(0, V, (B, g), device='cuda'); ua = torch.rand(B, g, device='cuda'); uf = torch.rand(B, device='cuda')
ops.verify(pp, qq, d, ua, uf)
ops.sample_rows(pp[:, 0].contiguous(), uf)
torch.cuda.synchronize(); ops.default_flag('cuda').check(); print('sanitizer script ok')
