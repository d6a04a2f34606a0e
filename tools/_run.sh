mkdir -p gpurun_out/r2p
o=gpurun_out/r2p
python -m pytest tests -m gpu -x -q > $o/pytest.log 2>&1; tail -3 $o/pytest.log
for dt in f32 bf16; do for V in 32000 50272; do
  python tools/microbench.py --mode dense --rows 576 --V $V --dtype $dt --sample --iters 200 >> $o/mb.log 2>&1
done; done
python tools/microbench.py --mode dense --rows 2368 --sample --iters 100 >> $o/mb.log 2>&1
python tools/microbench.py --mode dense --rows 576 --sample --iters 200 >> $o/mb.log 2>&1
cat $o/mb.log
