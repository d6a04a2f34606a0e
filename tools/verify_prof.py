"""In-kernel timeline of the row-staged dense verify (debug): mean clock64 deltas between its phases.

    python tools/verify_prof.py [--B 64] [--V 32000]
"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmspeculativesampling_b200 import ops, build, _cabi
from tools.kernel_bench import synth_probs

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=64); ap.add_argument("--V", type=int, default=32000); ap.add_argument("--gamma", type=int, default=4)
a = ap.parse_args()
build.build()
B, g, V = a.B, a.gamma, a.V
sets = [synth_probs(B, g, V, 1.0, 0, 0.0, s) for s in range(4)]
n_acc = torch.zeros(B, dtype=torch.int32, device="cuda"); nxt = torch.zeros(B, dtype=torch.int64, device="cuda")
err = ops.ErrFlag("cuda")
for s in sets:
    ops.verify(s["p"], s["q"], s["tok"], s["u_acc"], s["u_fin"], n_accepted=n_acc, next_tok=nxt, err=err)
torch.cuda.synchronize()
buf = torch.zeros(B, 8, dtype=torch.int64, device="cuda")
_cabi.load().sd_debug_set_prof(buf.data_ptr())
s = sets[1]
ops.verify(s["p"], s["q"], s["tok"], s["u_acc"], s["u_fin"], n_accepted=n_acc, next_tok=nxt, err=err)
torch.cuda.synchronize()
_cabi.load().sd_debug_set_prof(None)
b = buf.cpu().double()
names = ["start (after dependency wait)", "accept scan done", "row loads issued", "pass 1 done (residual, maximum)", "pass 2 done (exact sums)", "token found"]
for i in range(1, 6):
    m = (b[:, i] != 0) & (b[:, 0] != 0)
    d = (b[m, i] - b[m, 0])
    print(f"{names[i]:36s} +{d.mean():8.0f} cycles since start (min {d.min():.0f}, max {d.max():.0f}, n {int(m.sum())})")
print("mean accepted", float(n_acc.float().mean()))
