mkdir -p gpurun_out/r3c
o=gpurun_out/r3c
python -m pytest tests -m gpu -x -q > $o/pytest.log 2>&1; tail -3 $o/pytest.log
python tools/sweep.py > $o/sweep_n1.jsonl 2> $o/sweep.err; grep -c . $o/sweep_n1.jsonl
grep '"dense"' $o/sweep_n1.jsonl | cut -c1-170
