"""In-kernel timeline of the ring kernel (debug): mean clock64 deltas between the phases of a row, per item index.

    python tools/ring_prof.py [--mode topk|dense] [--rows 576] [--V 32000] [--dtype f32]
"""
import argparse, os, sys
os.environ["SD_LIB_VARIANT"] = "prof"          # the library copy with the probes compiled in (built on first use)
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmspeculativesampling_b200 import ops, build, _cabi

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="topk"); ap.add_argument("--rows", type=int, default=576)
ap.add_argument("--V", type=int, default=32000); ap.add_argument("--dtype", default="f32")
a = ap.parse_args()
build.build()
dt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[a.dtype]
T, k, p = (0.8, 20, 0.9) if a.mode == "topk" else (1.0, 0, 0.0)
x = (torch.randn(a.rows, a.V, device="cuda") * 3.8).to(dt)
out = torch.empty(a.rows, a.V, device="cuda")
u = torch.rand(a.rows, device="cuda")
for _ in range(3):
    ops.norm_sample(x, T, k, p, u, probs_out=out)
torch.cuda.synchronize()
NL = 4                                                    # back-to-back launches, each with its own timeline buffer
xs = [(torch.randn(a.rows, a.V, device="cuda") * 3.8).to(dt) for _ in range(NL)]
outs = [torch.empty(a.rows, a.V, device="cuda") for _ in range(NL)]
bufs = [torch.zeros(148 * 8 * 16 + 148 * 4, dtype=torch.int64, device="cuda") for _ in range(NL)]
torch.cuda.synchronize()
for j in range(NL):
    _cabi.load().sd_debug_set_prof(bufs[j].data_ptr())
    ops.norm_sample(xs[j], T, k, p, u, probs_out=outs[j])
torch.cuda.synchronize()
_cabi.load().sd_debug_set_prof(None)
cta = [bb.cpu()[148 * 8 * 16:].view(148, 4).double() for bb in bufs]
items = [bb.cpu()[:148 * 8 * 16].view(148, 8, 16).double() for bb in bufs]
for j in range(1, NL):
    prev_exit = cta[j - 1][:, 2].max()
    live = cta[j][:, 0] != 0
    ent, wt, ex = cta[j][live, 0], cta[j][live, 1], cta[j][live, 2]
    row0 = ent + (items[j][live, 0, 0] - cta[j][live, 3]) / 1.965          # first row start on the globaltimer axis (ns)
    print(f"launch {j}: period {(ex.max() - prev_exit) / 1000:.2f} us | since the previous launch's last exit: entry {(ent.min() - prev_exit) / 1000:.2f}..{(ent.max() - prev_exit) / 1000:.2f}, "
          f"dependency wait passed {(wt.min() - prev_exit) / 1000:.2f}..{(wt.max() - prev_exit) / 1000:.2f}, first row start {(row0.min() - prev_exit) / 1000:.2f}..{(row0.max() - prev_exit) / 1000:.2f}, "
          f"exit {(ex.min() - prev_exit) / 1000:.2f}..{(ex.max() - prev_exit) / 1000:.2f} (mean {(ex.mean() - prev_exit) / 1000:.2f}) us")
buf = bufs[-1][:148 * 8 * 16]
b = buf.cpu().view(148, 8, 16).double()
names = {"topk": ["row start", "chunks scanned", "tau barrier passed", "quad barrier passed", "search done", "sorted", "bar2 passed", "handed over", "aux: list received", "aux: finished", "aux: scattered", "-", "pass2 barrier passed"],
         "dense": ["row start", "pass A done", "combined", "-", "-", "-", "-", "pass B done"]}[a.mode]
t0 = b[:, 0, 0:1]
for it in range(8):
    live = b[:, it, 0] != 0
    if not live.any():
        break
    row = b[live, it]
    base = b[live, 0, 0]
    cells = []
    for s in range(13):
        m = row[:, s] != 0
        cells.append(f"{((row[m, s] - base[m]).mean() / 1000):7.2f}" if m.any() else "      -")
    print(f"item {it} ({int(live.sum())} CTAs): " + " ".join(cells))
print("columns (kcycles since the CTA's first row start): " + " | ".join(names))
