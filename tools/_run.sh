mkdir -p gpurun_out/r3g
o=gpurun_out/r3g
timeout 1200 python -m pytest tests -m gpu -x -q > $o/pytest.log 2>&1; tail -5 $o/pytest.log
python tools/sweep.py > $o/sweep_n1.jsonl 2> $o/sweep.err; grep -c . $o/sweep_n1.jsonl
