#!/bin/bash
# 8-GPU box (gpurun --gpus 8): BASELINE configs[2] / [3] / [4] at N = 8 (and the N = 2 / 4 training points), the 2-GPU tests
# the 1-GPU boxes skip.  Every line lands in gpurun_out/ and is copied to profiles/ by hand.
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/multi_summary.txt
tr() { local n=$1 port=$2; shift 2; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"; }
timeout 600 python -m pytest tests/test_gpu_train_step.py -q -m gpu -k "two_gpus or DistributedDataParallel or ddp" -p no:cacheprovider > $OUT/multi_tests.log 2>&1
echo "2-GPU tests exit $? :: $(tail -1 $OUT/multi_tests.log)" >> $OUT/multi_summary.txt
timeout 600 bash -c "$(declare -f tr); tr 8 29541 --steps 3 --warmup 3 --no-cpu-baseline" > $OUT/bench_b4096_n8.log 2> $OUT/bench_b4096_n8.err
echo "ddim50_cfg n8 exit $? :: $(head -c 260 $OUT/bench_b4096_n8.log)" >> $OUT/multi_summary.txt
timeout 600 bash -c "$(declare -f tr); tr 8 29542 --workload dit_ddim50 --batch 1024 --steps 3 --warmup 3 --no-cpu-baseline" > $OUT/bench_dit_b1024_n8.log 2> $OUT/bench_dit_b1024_n8.err
echo "dit n8 exit $? :: $(head -c 260 $OUT/bench_dit_b1024_n8.log)" >> $OUT/multi_summary.txt
for n in 8 4 2; do
  timeout 600 bash -c "$(declare -f tr); tr $n 2955$n --workload train --steps 20 --warmup 5 --no-cpu-baseline" > $OUT/train_n${n}_ddp.log 2> $OUT/train_n${n}_ddp.err
  echo "train ddp n$n exit $? :: $(head -c 200 $OUT/train_n${n}_ddp.log)" >> $OUT/multi_summary.txt
  DMC_NATIVE_ALLREDUCE=1 DMC_FUSED_OPT=1 timeout 600 bash -c "$(declare -f tr); tr $n 2956$n --workload train --steps 20 --warmup 5 --no-cpu-baseline" > $OUT/train_n${n}_native.log 2> $OUT/train_n${n}_native.err
  echo "train native n$n exit $? :: $(head -c 200 $OUT/train_n${n}_native.log)" >> $OUT/multi_summary.txt
done
timeout 600 python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline > $OUT/train_n1.log 2> $OUT/train_n1.err
echo "train n1 exit $? :: $(head -c 200 $OUT/train_n1.log)" >> $OUT/multi_summary.txt
DMC_FUSED_OPT=1 timeout 600 python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline > $OUT/train_n1_fusedopt.log 2> $OUT/train_n1_fusedopt.err
echo "train n1 fused opt exit $? :: $(head -c 200 $OUT/train_n1_fusedopt.log)" >> $OUT/multi_summary.txt
cat $OUT/multi_summary.txt
