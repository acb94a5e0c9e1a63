#!/bin/bash
# Round 2, run 28: row slabs for the 2x2 Upsample phase convolutions -- parity and same-box A/B
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py tests/test_gpu_conv_gn.py -q -m gpu --timeout 300 -p no:cacheprovider -x > $OUT/suite_o.log 2>&1; echo "tests exit $? :: $(tail -1 $OUT/suite_o.log)" >> $OUT/summary.txt
for v in 1 0 1 0; do
  DMC_CONV_SLAB_PHASE=$v timeout 300 python tools/bench_ops.py --batch 1024 --out $OUT/ops_phase_slab$v.json > $OUT/ops_phase_slab$v.log 2>&1
  echo "phase slab=$v :: $(tail -1 $OUT/ops_phase_slab$v.log | cut -c1-260)" >> $OUT/summary.txt
done
cat $OUT/summary.txt
