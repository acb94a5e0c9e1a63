#!/bin/bash
# 8-GPU box (gpurun --gpus 8), round 2 run 22: the reference trainer's own DDP(model) line around the native UNet (the wrapper is
# handed all parameters but one to ignore, the engine's per-entry all-reduce averages them) at N = 8 against N = 1 on the same box,
# stock DDP for comparison, the 2-GPU equivalence check, and DDIM-50 + CFG at N = 8 with this session's kernels.
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/multi_summary.txt
tr() { local n=$1 port=$2; shift 2; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"; }
timeout 400 python -m pytest tests/test_gpu_train_step.py -q -m gpu -k "two_gpus" -p no:cacheprovider > $OUT/multi_tests.log 2>&1
echo "2-GPU test exit $? :: $(tail -1 $OUT/multi_tests.log)" >> $OUT/multi_summary.txt
timeout 400 bash -c "$(declare -f tr); tr 8 29551 --workload train --steps 20 --warmup 5 --no-cpu-baseline" > $OUT/train_n8_ddp_coop.log 2> $OUT/train_n8_ddp_coop.err
echo "train DDP(model) n8 exit $? :: $(head -c 200 $OUT/train_n8_ddp_coop.log)" >> $OUT/multi_summary.txt
DMC_DDP_NATIVE=0 timeout 400 bash -c "$(declare -f tr); tr 8 29552 --workload train --steps 20 --warmup 5 --no-cpu-baseline" > $OUT/train_n8_ddp_stock.log 2> $OUT/train_n8_ddp_stock.err
echo "train stock DDP n8 exit $? :: $(head -c 200 $OUT/train_n8_ddp_stock.log)" >> $OUT/multi_summary.txt
timeout 400 python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline > $OUT/train_n1.log 2> $OUT/train_n1.err
echo "train n1 exit $? :: $(head -c 200 $OUT/train_n1.log)" >> $OUT/multi_summary.txt
timeout 500 bash -c "$(declare -f tr); tr 8 29541 --steps 3 --warmup 3 --no-cpu-baseline" > $OUT/bench_b4096_n8.log 2> $OUT/bench_b4096_n8.err
echo "ddim50_cfg n8 exit $? :: $(head -c 260 $OUT/bench_b4096_n8.log)" >> $OUT/multi_summary.txt
cat $OUT/multi_summary.txt
