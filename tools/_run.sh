mkdir -p gpurun_out/r3l
o=gpurun_out/r3l
python bench.py --steps 1600 --no-side-reports --no-cpu-baseline > $o/bench.json 2> $o/bench.err
python - <<PY
import json
d=json.loads([l for l in open("$o/bench.json") if l.startswith("{")][-1])
r=d["roofline"]
print("ms_per_step", round(d["ms_per_step"]*1e3,2), "value", round(d["value"]), "norm", round(r["ms_per_step"]*1e3,2), "frac", round(r["frac"],3), "overlap", round(r["overlapped"]["ms_per_launch"]*1e3,2), "step frac", round(r["step"]["frac"],3), "serial", round(r["step"]["serial_ms"]*1e3,2))
PY
for dt in f32 bf16; do python tools/microbench.py --mode dense --rows 576 --dtype $dt --iters 200 >> $o/mb.log 2>&1; python tools/microbench.py --mode dense --rows 576 --dtype $dt --iters 200 --sample >> $o/mb.log 2>&1; python tools/microbench.py --mode topk --rows 576 --dtype $dt --iters 200 --sample >> $o/mb.log 2>&1;  done; cat $o/mb.log
