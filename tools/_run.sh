mkdir -p gpurun_out/r3d
o=gpurun_out/r3d
timeout 600 python -m pytest tests/test_gpu_norm.py -x -q > $o/pytest.log 2>&1; tail -3 $o/pytest.log
for dt in f32 bf16; do for V in 65536 131072 262144; do
  timeout 120 python tools/microbench.py --mode dense --rows 576 --V $V --dtype $dt --iters 50 >> $o/mb.log 2>&1
  timeout 120 python tools/microbench.py --mode dense --rows 576 --V $V --dtype $dt --iters 50 --sample >> $o/mb.log 2>&1
done; done; cat $o/mb.log
