mkdir -p gpurun_out/r2w
o=gpurun_out/r2w
for rows in 32 64 128 148; do for m in topk dense; do for dt in f32 bf16; do
  python tools/microbench.py --mode $m --rows $rows --dtype $dt --sample --iters 300 >> $o/ring.log 2>&1
  python tools/microbench.py --mode $m --rows $rows --dtype $dt --sample --iters 300 --no-ring >> $o/noring.log 2>&1
done; done; done
paste -d'\n' $o/ring.log $o/noring.log | cut -c1-160
