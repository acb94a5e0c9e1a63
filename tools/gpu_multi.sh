#!/bin/bash
# N-GPU smoke of the sharded bench (NCCL all-gather of the final images); run with gpurun --gpus N
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py \
    --gpus $N --steps 1 --warmup 3 --batch ${BENCH_BATCH:-4096} > gpurun_out/bench_n$N.log 2>&1
echo "bench n=$N exit $?"; tail -c 1500 gpurun_out/bench_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py \
    --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1
echo "reference arm n=$N exit $?"; tail -c 600 gpurun_out/bench_ref_n$N.log
