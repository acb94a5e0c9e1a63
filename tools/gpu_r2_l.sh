#!/bin/bash
# Round 2, run 15: shared-memory accesses of the tcgen05 kernels as LDS / STS (they were generic LD.E / ST.E), ping-pong attention.
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt; rm -f $OUT/eps_errors.txt
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x > $OUT/suite.log 2>&1; echo "suite exit $? :: $(tail -1 $OUT/suite.log)" >> $OUT/summary.txt
timeout 900 python bench.py --steps 3 --warmup 3 --ops-out $OUT/ops.json --no-cpu-baseline > $OUT/bench.log 2> $OUT/bench.err
echo "bench exit $? :: $(head -c 250 $OUT/bench.log)" >> $OUT/summary.txt
timeout 900 python bench.py --model dit --steps 3 --warmup 3 --ops-out $OUT/ops_dit.json --no-cpu-baseline > $OUT/bench_dit.log 2> $OUT/bench_dit.err
echo "bench_dit exit $? :: $(head -c 250 $OUT/bench_dit.log)" >> $OUT/summary.txt
timeout 900 python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_train.log 2> $OUT/bench_train.err
echo "bench_train exit $? :: $(head -c 250 $OUT/bench_train.log)" >> $OUT/summary.txt
cat $OUT/summary.txt
