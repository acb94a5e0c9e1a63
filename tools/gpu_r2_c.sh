#!/bin/bash
# round-2 run 3: reworked fused-GroupNorm epilogue (stores after the hand-shake) -- tests, then A/B bench with per-op tables
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt
run() { local name=$1; shift; timeout 600 python -m pytest "$@" -q -m gpu --timeout 300 -p no:cacheprovider -x > $OUT/$name.log 2>&1; echo "$name exit $? :: $(tail -1 $OUT/$name.log)" >> $OUT/summary.txt; }
run conv_gn tests/test_gpu_conv_gn.py
run unet tests/test_gpu_unet.py
run config1 tests/test_gpu_config1.py tests/test_gpu_optim.py
for kb in 16 4; do
DMC_FUSE_GN_MIN_KB=$kb timeout 900 python bench.py --batch 2048 --steps 2 --warmup 3 --ops-out $OUT/ops_kb$kb.json --no-cpu-baseline > $OUT/bench_kb$kb.log 2> $OUT/bench_kb$kb.err
echo "bench_kb$kb exit $? :: $(head -c 200 $OUT/bench_kb$kb.log)" >> $OUT/summary.txt
done
DMC_FUSE_GN=0 timeout 900 python bench.py --batch 2048 --steps 2 --warmup 3 --ops-out $OUT/ops_unfused.json --no-cpu-baseline > $OUT/bench_unfused.log 2> $OUT/bench_unfused.err
echo "bench_unfused exit $? :: $(head -c 200 $OUT/bench_unfused.log)" >> $OUT/summary.txt
cat $OUT/summary.txt
