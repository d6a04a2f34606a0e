mkdir -p gpurun_out/r2v
o=gpurun_out/r2v
for pipe in 3 2; do
python bench.py --steps 800 --pipeline $pipe --no-side-reports > $o/bench_p${pipe}.json 2> $o/bench_p${pipe}.err
python - <<PY
import json
d=json.loads([l for l in open("$o/bench_p${pipe}.json") if l.startswith("{")][-1])
r=d["roofline"]
print("pipe $pipe", "ms_per_step", round(d["ms_per_step"]*1e3,2), "value", round(d["value"]), "norm", round(r["ms_per_step"]*1e3,2), "frac", round(r["frac"],3), "overlap", round(r["overlapped"]["ms_per_launch"]*1e3,2), round(r["overlapped"]["frac"],3), "step frac", round(r["step"]["frac"],3), "serial", round(r["step"]["serial_ms"]*1e3,2))
PY
done
