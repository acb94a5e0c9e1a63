#!/bin/bash
# First-contact / regression run on the GPU box: every GPU test group in its own process (a trapped kernel poisons
# only its own group), logs under gpurun_out/.
mkdir -p gpurun_out
OUT=gpurun_out
: > $OUT/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/smi.txt 2>&1
ls /root/reference > $OUT/ref_ls.txt 2>&1
run() {  # name, pytest args...
  local name=$1; shift
  timeout 900 python -m pytest "$@" -q -m gpu --timeout 300 -p no:cacheprovider > $OUT/$name.log 2>&1
  echo "$name exit $? :: $(tail -1 $OUT/$name.log)" >> $OUT/summary.txt
}
run sched tests/test_gpu_sched.py
for k in conv3x3 conv_kernel_variants conv_split fused_output_head conv1x1 shortcut head_conv phase stem groupnorm attention conditioning "upsample_nearest or bad_arguments"; do
  run "ops_${k%% *}" tests/test_gpu_ops.py -k "$k"
done
run unet tests/test_gpu_unet.py
run dit tests/test_gpu_dit.py
run train_kernels tests/test_gpu_train.py
run train_step tests/test_gpu_train_step.py
if [ "$1" != "nobench" ]; then
  timeout 900 python bench.py --batch ${BENCH_BATCH:-1024} --steps 1 --warmup 3 --ops-out $OUT/ops.json > $OUT/bench.log 2>&1
  echo "bench exit $? :: $(tail -c 300 $OUT/bench.log)" >> $OUT/summary.txt
fi
cat $OUT/summary.txt
