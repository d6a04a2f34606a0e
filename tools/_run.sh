mkdir -p gpurun_out/r2sw
python tools/sweep.py > gpurun_out/r2sw/sweep_n1.jsonl 2> gpurun_out/r2sw/sweep_n1.err; wc -l gpurun_out/r2sw/sweep_n1.jsonl; tail -3 gpurun_out/r2sw/sweep_n1.err
