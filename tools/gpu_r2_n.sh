#!/bin/bash
# Round 2, run 20: everything of this session together (attention with P in tensor memory, straight-line epilogue, LDS/STS):
# full suite, the three bench workloads with op tables, then the ncu evidence (tools/gpu_profile.sh).
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt; rm -f $OUT/eps_errors.txt
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x > $OUT/suite.log 2>&1; echo "suite exit $? :: $(tail -1 $OUT/suite.log)" >> $OUT/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 --ops-out $OUT/ops.json --no-cpu-baseline > $OUT/bench.log 2> $OUT/bench.err
echo "bench exit $? :: $(head -c 250 $OUT/bench.log)" >> $OUT/summary.txt
timeout 900 python bench.py --model dit --batch 1024 --steps 5 --warmup 3 --ops-out $OUT/ops_dit.json --no-cpu-baseline > $OUT/bench_dit.log 2> $OUT/bench_dit.err
echo "bench_dit exit $? :: $(head -c 250 $OUT/bench_dit.log)" >> $OUT/summary.txt
timeout 900 python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_train.log 2> $OUT/bench_train.err
echo "bench_train exit $? :: $(head -c 250 $OUT/bench_train.log)" >> $OUT/summary.txt
timeout 900 python bench.py --workload ddpm1000 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_ddpm.log 2> $OUT/bench_ddpm.err
echo "bench_ddpm exit $? :: $(head -c 250 $OUT/bench_ddpm.log)" >> $OUT/summary.txt
PROFILE_OPS=up_blocks.6.0.conv1,up_blocks.9.0.conv1,down_blocks.0.0.conv1,up_blocks.6.1.qkv,up_blocks.6.1.proj,up_blocks.9.0.conv1.0,attention,output.2.taps bash tools/gpu_profile.sh
cat $OUT/summary.txt
