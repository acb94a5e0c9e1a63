#!/bin/bash
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt; rm -f $OUT/conv_gn_micro.txt
run() { local name=$1; shift; timeout 600 python -m pytest "$@" -q -m gpu --timeout 300 -p no:cacheprovider -x > $OUT/$name.log 2>&1; echo "$name exit $? :: $(tail -1 $OUT/$name.log)" >> $OUT/summary.txt; }
run conv_gn tests/test_gpu_conv_gn.py
run unet tests/test_gpu_unet.py
for dbg in 0 1 2; do
  echo "== DMC_GN_DEBUG=$dbg" >> $OUT/conv_gn_micro.txt
  DMC_GN_DEBUG=$dbg timeout 300 python tools/bench_conv_gn.py >> $OUT/conv_gn_micro.txt 2>&1
done
echo "== DMC_GN_SC_SMEM=0" >> $OUT/conv_gn_micro.txt
DMC_GN_SC_SMEM=0 timeout 300 python tools/bench_conv_gn.py >> $OUT/conv_gn_micro.txt 2>&1
for kb in 16 30 40; do
DMC_FUSE_GN_MIN_KB=$kb timeout 900 python bench.py --batch 2048 --steps 2 --warmup 3 --ops-out $OUT/ops_kb$kb.json --no-cpu-baseline > $OUT/bench_kb$kb.log 2> $OUT/bench_kb$kb.err
echo "bench_kb$kb exit $? :: $(head -c 200 $OUT/bench_kb$kb.log)" >> $OUT/summary.txt
done
DMC_FUSE_GN=0 timeout 900 python bench.py --batch 2048 --steps 2 --warmup 3 --ops-out $OUT/ops_unfused.json --no-cpu-baseline > $OUT/bench_unfused.log 2> $OUT/bench_unfused.err
echo "bench_unfused exit $? :: $(head -c 200 $OUT/bench_unfused.log)" >> $OUT/summary.txt
cat $OUT/summary.txt; cat $OUT/conv_gn_micro.txt
