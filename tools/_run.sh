mkdir -p gpurun_out/r4a
o=gpurun_out/r4a
python -m pytest tests -m gpu -q > $o/pytest.log 2>&1; tail -2 $o/pytest.log
python __graft_entry__.py smoke > $o/smoke.log 2>&1; tail -1 $o/smoke.log
python bench.py > $o/bench_n1.json 2> $o/bench.err; tail -c 300 $o/bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > $o/bench_ref.json 2> $o/bench_ref.err; tail -c 400 $o/bench_ref.json
python tools/kernel_bench.py --mode all > $o/kb.jsonl 2> $o/kb.err
python tools/kernel_bench.py --mode verify_dense --V 50272 >> $o/kb.jsonl 2>> $o/kb.err
python tools/kernel_bench.py --mode verify_dense --B 256 >> $o/kb.jsonl 2>> $o/kb.err
python tools/sweep.py > $o/sweep_n1.jsonl 2> $o/sweep.err
T=/tmp/ncu; mkdir -p $T
for m in verify_dense verify_sparse verify_multi verify_bild sample max_fn kv_append build_step; do
  ncu --set full --clock-control none -k regex:'verify|max_fn|kv_|build_step|multi_commit' -c 1 -o $T/$m python tools/kernel_bench.py --mode $m --once > $T/$m.log 2>&1
done
ncu --set full --clock-control none -k regex:verify_row_kernel -c 1 -o $T/verify_dense_B256 python tools/kernel_bench.py --mode verify_dense --B 256 --once > $T/vd256.log 2>&1
ncu --set full --clock-control none -k regex:kv_select -c 1 -o $T/kv_select python tools/kernel_bench.py --mode kv_append --once > $T/kv_select.log 2>&1
ncu --set full --clock-control none -k regex:norm_ring -s 4 -c 1 -o $T/ring_topk_f32_V32000 python tools/microbench.py --mode topk --rows 576 --sample --iters 3 > $T/a.log 2>&1
ncu --set full --clock-control none -k regex:norm_ring -s 4 -c 1 -o $T/ring_dense_f32_V32000 python tools/microbench.py --mode dense --rows 576 --sample --iters 3 > $T/b.log 2>&1
ncu --set full --clock-control none -k regex:norm_ring -s 4 -c 1 -o $T/ring_dense_bf16_V50272 python tools/microbench.py --mode dense --rows 576 --V 50272 --dtype bf16 --sample --iters 3 > $T/c.log 2>&1
ncu --set full --clock-control none -k regex:norm_ring -s 4 -c 1 -o $T/ring_topk_bf16_V50272 python tools/microbench.py --mode topk --rows 576 --V 50272 --dtype bf16 --sample --iters 3 > $T/d.log 2>&1
ncu --set full --clock-control none -k regex:norm_ring -s 4 -c 1 -o $T/ring_dense_long_bf16_V131072 python tools/microbench.py --mode dense --rows 576 --V 131072 --dtype bf16 --sample --iters 3 > $T/e.log 2>&1
ncu --set full --clock-control none -k regex:norm_ring -s 4 -c 1 -o $T/ring_topk_long_f32_V131072 python tools/microbench.py --mode topk --rows 576 --V 131072 --sample --iters 3 > $T/f.log 2>&1
python tools/ncu_summary.py $T/*.ncu-rep > $o/ncu_kernels.md 2> $o/ncu_summary.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_bench.csv python bench.py --steps 32 --warmup 3 --no-side-reports --no-cpu-baseline > $o/ncu_bench.log 2>&1
ls -la $o; du -sh gpurun_out
