mkdir -p gpurun_out/r4b
o=gpurun_out/r4b
timeout 900 python -m pytest tests/test_gpu_verify.py tests/test_gpu_engine.py tests/test_gpu_dropins.py -x -q > $o/pytest.log 2>&1; tail -3 $o/pytest.log
python tools/verify_prof.py --B 128 > $o/vprof.log 2>&1; cat $o/vprof.log
python tools/kernel_bench.py --mode verify_dense > $o/kb.jsonl 2> $o/kb.err
python tools/kernel_bench.py --mode verify_dense --V 50272 >> $o/kb.jsonl 2>> $o/kb.err
python tools/kernel_bench.py --mode verify_dense --B 256 >> $o/kb.jsonl 2>> $o/kb.err
python tools/kernel_bench.py --mode sample >> $o/kb.jsonl 2>> $o/kb.err
python tools/kernel_bench.py --mode verify_multi >> $o/kb.jsonl 2>> $o/kb.err
python tools/kernel_bench.py --mode verify_bild >> $o/kb.jsonl 2>> $o/kb.err
cut -c1-160 $o/kb.jsonl
