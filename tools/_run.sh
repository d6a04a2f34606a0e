mkdir -p gpurun_out/r2z
o=gpurun_out/r2z
python -m pytest tests -m gpu -x -q > $o/pytest.log 2>&1; tail -4 $o/pytest.log
