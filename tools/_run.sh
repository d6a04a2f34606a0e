mkdir -p gpurun_out/r3e
o=gpurun_out/r3e
python -m pytest tests -m gpu -q > $o/pytest.log 2>&1; tail -2 $o/pytest.log
python tools/sweep.py > $o/sweep_n1.jsonl 2> $o/sweep.err
python tools/kernel_bench.py --mode all > $o/kb.jsonl 2> $o/kb.err
python tools/kernel_bench.py --mode verify_dense --V 50272 >> $o/kb.jsonl 2>> $o/kb.err
python tools/kernel_bench.py --mode verify_dense --B 256 >> $o/kb.jsonl 2>> $o/kb.err
python bench.py > $o/bench_n1.json 2> $o/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_bench.csv python bench.py --steps 32 --warmup 3 --no-side-reports --no-cpu-baseline > $o/ncu_bench.log 2>&1
tail -c 600 $o/bench_n1.json; wc -l $o/launches_bench.csv
