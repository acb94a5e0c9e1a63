#!/bin/bash
# what the driver runs at round end: GPU suite, smoke, the reference arm and the native arm with its flags
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt; rm -f $OUT/eps_errors.txt
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider > $OUT/gpu_suite.log 2>&1
echo "suite exit $? :: $(tail -1 $OUT/gpu_suite.log)" >> $OUT/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/summary.txt
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_ref.log 2> $OUT/bench_ref.err ) 2> $OUT/bench_ref.time
echo "bench_ref exit $? :: $(head -c 200 $OUT/bench_ref.log) :: $(grep real $OUT/bench_ref.time)" >> $OUT/summary.txt
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.log 2> $OUT/bench.err ) 2> $OUT/bench.time
echo "bench exit $? :: $(head -c 200 $OUT/bench.log) :: $(grep real $OUT/bench.time)" >> $OUT/summary.txt
cat $OUT/summary.txt
