#!/bin/bash
# Round 2, run 17: straight-line box loop in the conv epilogue (epi_fast) -- parity and same-box A/B (UNet and DiT op tables).
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt; rm -f $OUT/eps_errors.txt
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x > $OUT/suite.log 2>&1; echo "suite exit $? :: $(tail -1 $OUT/suite.log)" >> $OUT/summary.txt
for fast in 1 0; do
  DMC_CONV_EPI_FAST=$fast timeout 600 python tools/bench_ops.py --batch 1024 --out $OUT/ops_fast$fast.json > $OUT/ops_fast$fast.log 2>&1
  echo "unet ops fast=$fast exit $? :: $(tail -1 $OUT/ops_fast$fast.log | cut -c1-300)" >> $OUT/summary.txt
  DMC_CONV_EPI_FAST=$fast timeout 900 python bench.py --model dit --steps 3 --warmup 3 --ops-out $OUT/ops_dit_fast$fast.json --no-cpu-baseline > $OUT/bench_dit_fast$fast.log 2> $OUT/bench_dit_fast$fast.err
  echo "bench_dit fast=$fast exit $? :: $(head -c 200 $OUT/bench_dit_fast$fast.log)" >> $OUT/summary.txt
done
cat $OUT/summary.txt
