#!/bin/bash
# round-2: head as 1x1 GEMM + tap gather (tests + A/B), then the ncu evidence for profiles/ (tools/gpu_profile.sh)
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt
run() { local name=$1; shift; timeout 600 python -m pytest "$@" -q -m gpu --timeout 300 -p no:cacheprovider -x > $OUT/$name.log 2>&1; echo "$name exit $? :: $(tail -1 $OUT/$name.log)" >> $OUT/summary.txt; }
run head tests/test_gpu_ops.py -k "head"
run unet tests/test_gpu_unet.py tests/test_gpu_config1.py
timeout 900 python bench.py --steps 3 --warmup 3 --ops-out $OUT/ops.json --no-cpu-baseline > $OUT/bench.log 2> $OUT/bench.err
echo "bench exit $? :: $(head -c 250 $OUT/bench.log)" >> $OUT/summary.txt
DMC_HEAD_TAPS=0 timeout 900 python bench.py --steps 3 --warmup 3 --ops-out $OUT/ops_nohead.json --no-cpu-baseline > $OUT/bench_nohead.log 2> $OUT/bench_nohead.err
echo "bench_nohead exit $? :: $(head -c 250 $OUT/bench_nohead.log)" >> $OUT/summary.txt
PROFILE_OPS=up_blocks.6.0.conv1,up_blocks.9.0.conv1,down_blocks.0.0.conv1,up_blocks.6.1.qkv,up_blocks.9.0.conv1.0,attention,output.2.taps,output.2.gather NCU_SKIP=1500 NCU_COUNT=400 BENCH_BATCH=1024 bash tools/gpu_profile.sh
cat $OUT/summary.txt
