#!/bin/bash
mkdir -p gpurun_out; OUT=gpurun_out
for dbg in 0 1 2 3 7 15; do
  echo "== DMC_GN_DEBUG=$dbg" >> $OUT/conv_gn_micro.txt
  DMC_GN_DEBUG=$dbg timeout 300 python tools/bench_conv_gn.py >> $OUT/conv_gn_micro.txt 2>&1
done
cat $OUT/conv_gn_micro.txt
