#!/bin/bash
# 8-GPU box, run 33: BASELINE configs[3] (DiT patch-2 DDIM-50, batch 1024 sharded over 8) with the second session's kernels
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/multi3_summary.txt
tr() { local n=$1 port=$2; shift 2; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"; }
timeout 400 bash -c "$(declare -f tr); tr 8 29571 --workload dit_ddim50 --batch 1024 --steps 5 --warmup 3 --no-cpu-baseline" > $OUT/bench_dit_b1024_n8.log 2> $OUT/bench_dit_b1024_n8.err
echo "dit n8 exit $? :: $(grep '^{' $OUT/bench_dit_b1024_n8.log | head -c 260)" >> $OUT/multi3_summary.txt
timeout 300 python bench.py --workload dit_ddim50 --batch 1024 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_dit_b1024_n1.log 2> $OUT/bench_dit_b1024_n1.err
echo "dit n1 exit $? :: $(head -c 260 $OUT/bench_dit_b1024_n1.log)" >> $OUT/multi3_summary.txt
cat $OUT/multi3_summary.txt
