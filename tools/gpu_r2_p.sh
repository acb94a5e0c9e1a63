#!/bin/bash
# Round 2, run 30: DiT qkv weight rows padded to a multiple of 256 (1152 -> 1280: the 256-column tile) -- parity and same-box A/B
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt
timeout 600 python -m pytest tests/test_gpu_dit.py tests/test_dim.py tests/test_gpu_zz_train_dit.py -q -m gpu --timeout 300 -p no:cacheprovider -x > $OUT/suite_p.log 2>&1; echo "tests exit $? :: $(tail -1 $OUT/suite_p.log)" >> $OUT/summary.txt
for v in 1 0 1 0; do
  DMC_DIT_PAD_QKV=$v timeout 300 python bench.py --model dit --batch 1024 --steps 3 --warmup 3 --ops-out $OUT/ops_dit_pad$v.json --no-cpu-baseline > $OUT/bench_dit_pad$v.log 2>&1
  echo "pad=$v :: $(head -c 160 $OUT/bench_dit_pad$v.log)" >> $OUT/summary.txt
done
cat $OUT/summary.txt
