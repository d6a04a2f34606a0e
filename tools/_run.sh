mkdir -p gpurun_out/r2s
o=gpurun_out/r2s
python -m pytest tests/test_gpu_engine.py -x -q -k "evaluation or golden_runs or kvcache" > $o/pytest.log 2>&1; tail -30 $o/pytest.log
T=/tmp/ncu; mkdir -p $T
for m in verify_dense verify_sparse verify_multi verify_bild sample max_fn kv_append build_step; do
  ncu --set full --import-source on --clock-control none -k regex:'verify|max_fn|kv_|build_step|multi_commit' -c 1 -o $T/$m python tools/kernel_bench.py --mode $m --once > $T/$m.log 2>&1
done
ncu --set full --import-source on --clock-control none -k regex:kv_select -c 1 -o $T/kv_select python tools/kernel_bench.py --mode kv_append --once > $T/kv_select.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:norm_ring -s 4 -c 1 -o $T/ring_topk_f32_V32000 python tools/microbench.py --mode topk --rows 576 --sample --iters 3 > $T/a.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:norm_ring -s 4 -c 1 -o $T/ring_dense_f32_V32000 python tools/microbench.py --mode dense --rows 576 --sample --iters 3 > $T/b.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:norm_ring -s 4 -c 1 -o $T/ring_dense_bf16_V50272 python tools/microbench.py --mode dense --rows 576 --V 50272 --dtype bf16 --sample --iters 3 > $T/c.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:norm_ring -s 4 -c 1 -o $T/ring_topk_bf16_V50272 python tools/microbench.py --mode topk --rows 576 --V 50272 --dtype bf16 --sample --iters 3 > $T/d.log 2>&1
python tools/ncu_summary.py $T/*.ncu-rep > $o/ncu_kernels.md 2> $o/ncu_summary.err
cp $T/verify_dense.ncu-rep $T/ring_topk_f32_V32000.ncu-rep $o/ 2>/dev/null
ls -la $T $o | head -40; du -sh gpurun_out
