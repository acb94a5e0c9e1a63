#!/bin/bash
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt; rm -f $OUT/conv_gn_micro.txt
run() { local name=$1; shift; timeout 600 python -m pytest "$@" -q -m gpu --timeout 300 -p no:cacheprovider -x > $OUT/$name.log 2>&1; echo "$name exit $? :: $(tail -1 $OUT/$name.log)" >> $OUT/summary.txt; }
run conv_gn tests/test_gpu_conv_gn.py
run unet tests/test_gpu_unet.py tests/test_gpu_ops.py
for pf in 1 0; do
  echo "== DMC_CONV_PREFETCH_COND=$pf" >> $OUT/conv_gn_micro.txt
  DMC_CONV_PREFETCH_COND=$pf timeout 300 python tools/bench_conv_gn.py >> $OUT/conv_gn_micro.txt 2>&1
done
for pf in 1 0 1 0; do
DMC_CONV_PREFETCH_COND=$pf timeout 900 python bench.py --batch 2048 --steps 3 --warmup 3 --ops-out $OUT/ops_pf$pf.json --no-cpu-baseline > $OUT/bench_pf$pf.log 2> $OUT/bench_pf$pf.err
echo "bench_pf$pf exit $? :: $(head -c 130 $OUT/bench_pf$pf.log)" >> $OUT/summary.txt
done
cat $OUT/summary.txt; cat $OUT/conv_gn_micro.txt
