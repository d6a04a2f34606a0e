mkdir -p gpurun_out/r2r
o=gpurun_out/r2r
python -m pytest tests -m gpu -x -q > $o/pytest.log 2>&1; tail -4 $o/pytest.log
python tools/kernel_bench.py --mode verify_dense > $o/kb.jsonl 2> $o/kb.err
python tools/kernel_bench.py --mode verify_dense --V 50272 >> $o/kb.jsonl 2>> $o/kb.err
python tools/kernel_bench.py --mode verify_dense --B 256 >> $o/kb.jsonl 2>> $o/kb.err
python tools/kernel_bench.py --mode verify_multi >> $o/kb.jsonl 2>> $o/kb.err
python tools/kernel_bench.py --mode verify_bild >> $o/kb.jsonl 2>> $o/kb.err
python tools/kernel_bench.py --mode sample >> $o/kb.jsonl 2>> $o/kb.err
cut -c1-230 $o/kb.jsonl
