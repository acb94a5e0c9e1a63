#!/bin/bash
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt; rm -f $OUT/eps_errors.txt
run() { local name=$1; shift; timeout 600 python -m pytest "$@" -q -m gpu --timeout 300 -p no:cacheprovider -x > $OUT/$name.log 2>&1; echo "$name exit $? :: $(tail -1 $OUT/$name.log)" >> $OUT/summary.txt; }
run aff tests/test_gpu_conv_gn.py -k "qkv_operand"
run conv_gn tests/test_gpu_conv_gn.py
run unet tests/test_gpu_unet.py tests/test_gpu_config1.py tests/test_gpu_sched.py tests/test_gpu_ops.py
timeout 900 python bench.py --steps 3 --warmup 3 --ops-out $OUT/ops.json --no-cpu-baseline > $OUT/bench.log 2> $OUT/bench.err
echo "bench exit $? :: $(head -c 250 $OUT/bench.log)" >> $OUT/summary.txt
DMC_FUSE_NORM_QKV=0 timeout 900 python bench.py --steps 3 --warmup 3 --ops-out $OUT/ops_noaff.json --no-cpu-baseline > $OUT/bench_noaff.log 2> $OUT/bench_noaff.err
echo "bench_noaff exit $? :: $(head -c 250 $OUT/bench_noaff.log)" >> $OUT/summary.txt
cat $OUT/summary.txt
