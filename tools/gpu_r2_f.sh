#!/bin/bash
# round-2: full GPU suite + smoke + default bench (both arms) with the final fusion policy
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt; rm -f $OUT/eps_errors.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $OUT/gpu_suite.log 2>&1
echo "suite exit $? :: $(tail -1 $OUT/gpu_suite.log)" >> $OUT/summary.txt
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/summary.txt
timeout 900 python bench.py --steps 3 --warmup 3 --ops-out $OUT/ops.json > $OUT/bench.log 2> $OUT/bench.err
echo "bench exit $? :: $(head -c 300 $OUT/bench.log)" >> $OUT/summary.txt
DMC_FUSE_GN=0 timeout 900 python bench.py --steps 3 --warmup 3 --ops-out $OUT/ops_unfused.json --no-cpu-baseline > $OUT/bench_unfused.log 2> $OUT/bench_unfused.err
echo "bench_unfused exit $? :: $(head -c 300 $OUT/bench_unfused.log)" >> $OUT/summary.txt
DMC_FUSE_GN_SCOPE=all DMC_FUSE_GN_MIN_KB=30 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-roofline > $OUT/bench_all30.log 2> $OUT/bench_all30.err
echo "bench_all30 exit $? :: $(head -c 300 $OUT/bench_all30.log)" >> $OUT/summary.txt
timeout 600 python bench.py --workload ddpm1000 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/bench_ddpm1000.log 2> $OUT/bench_ddpm1000.err
echo "ddpm1000 exit $? :: $(head -c 300 $OUT/bench_ddpm1000.log)" >> $OUT/summary.txt
cat $OUT/summary.txt
