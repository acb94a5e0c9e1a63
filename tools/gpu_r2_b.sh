#!/bin/bash
# round-2 run 2: the fused-GroupNorm conv epilogue -- op-level tests first (each group in its own process, bounded), then the
# whole-model tests and the bench with the per-op table
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt
run() { local name=$1; shift; timeout 600 python -m pytest "$@" -q -m gpu --timeout 300 -p no:cacheprovider -x > $OUT/$name.log 2>&1; echo "$name exit $? :: $(tail -1 $OUT/$name.log)" >> $OUT/summary.txt; }
run conv_gn tests/test_gpu_conv_gn.py
run ops tests/test_gpu_ops.py
run unet tests/test_gpu_unet.py
run config1 tests/test_gpu_config1.py tests/test_gpu_optim.py
run rest tests/test_gpu_sched.py tests/test_gpu_dit.py tests/test_dim.py tests/test_gpu_eval_shape.py tests/test_gpu_dropin_sample.py
run train tests/test_gpu_train.py tests/test_gpu_train_step.py tests/test_gpu_zz_train_fixture.py
timeout 900 python bench.py --steps 2 --warmup 3 --ops-out $OUT/ops.json --no-cpu-baseline > $OUT/bench.log 2> $OUT/bench.err
echo "bench exit $? :: $(head -c 600 $OUT/bench.log)" >> $OUT/summary.txt
DMC_FUSE_GN=0 timeout 900 python bench.py --steps 2 --warmup 3 --ops-out $OUT/ops_unfused.json --no-cpu-baseline > $OUT/bench_unfused.log 2> $OUT/bench_unfused.err
echo "bench_unfused exit $? :: $(head -c 300 $OUT/bench_unfused.log)" >> $OUT/summary.txt
cat $OUT/summary.txt
