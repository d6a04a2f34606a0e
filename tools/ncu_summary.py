"""Summarises `ncu --set full` reports (.ncu-rep) as one markdown table per file: per launch the duration, DRAM bytes and
GB/s (dram__bytes_read + dram__bytes_write over gpu__time_duration — cold-cache, serialised launches: compare SHARES and
traffic, not absolute times), DRAM / issue utilisation, registers, shared memory and the top warp-stall reasons.

    python tools/ncu_summary.py gpurun_out/x/a.ncu-rep [b.ncu-rep ...] > profiles/r02_ncu_kernels.md      (runs without a GPU)
"""
import csv
import io
import os
import subprocess
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def to_bytes(v, unit):
    return num(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    return num(v) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


def main():
    for path in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            print(f"## {os.path.basename(path)}\n\n(no launches)\n")
            continue
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        print(f"## {os.path.basename(path)}\n")
        print("| kernel | grid x block | regs | dyn smem KB | us | DRAM read MB | DRAM write MB | DRAM GB/s | DRAM % of peak | issue active % | top stalls (warps per issue) |")
        print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---|")
        for r in rows[2:]:
            def g(k):
                return r[col[k]] if k in col else "nan"
            def u(k):
                return units[col[k]] if k in col else ""
            t = to_us(g("gpu__time_duration.sum"), u("gpu__time_duration.sum"))
            rd = to_bytes(g("dram__bytes_read.sum"), u("dram__bytes_read.sum"))
            wr = to_bytes(g("dram__bytes_write.sum"), u("dram__bytes_write.sum"))
            stalls = sorted(((num(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                             for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
                             and "not_issued" not in h), reverse=True)[:4]
            name = g("Kernel Name").replace("sd::", "").replace("void ", "")
            name = name[:name.index("(")] if "(" in name else name
            grid = f'{g("launch__grid_size")} x {g("launch__block_size")}'
            print(f'| `{name}` | {grid} | {g("launch__registers_per_thread")} | {num(g("launch__shared_mem_per_block_dynamic")) / 1e3 if u("launch__shared_mem_per_block_dynamic") == "byte" else g("launch__shared_mem_per_block_dynamic")} | {t:.2f} | {rd / 1e6:.2f} | {wr / 1e6:.2f} | '
                  f'{(rd + wr) / t / 1e3:.0f} | {num(g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")):.1f} | {num(g("smsp__issue_active.avg.pct_of_peak_sustained_active")):.1f} | '
                  + ", ".join(f"{n} {v:.2f}" for v, n in stalls) + " |")
        print()


if __name__ == "__main__":
    main()
