#!/bin/bash
# last check of the round: GPU suite, smoke, a short native bench line (the 20-step lines of both arms are run 29's)
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt; rm -f $OUT/eps_errors.txt
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider > $OUT/gpu_suite.log 2>&1
echo "suite exit $? :: $(tail -1 $OUT/gpu_suite.log)" >> $OUT/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $? :: $(tail -1 $OUT/smoke.log)" >> $OUT/summary.txt
( time timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --ops-out $OUT/ops_final.json > $OUT/bench.log 2> $OUT/bench.err ) 2> $OUT/bench.time
echo "bench exit $? :: $(head -c 200 $OUT/bench.log) :: $(grep real $OUT/bench.time)" >> $OUT/summary.txt
cat $OUT/summary.txt
