#!/bin/bash
# round-2 first contact: whole GPU suite, smoke, bench (both arms), side workloads, first compute-sanitizer pass
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/smi.txt 2>&1
nproc >> $OUT/smi.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $OUT/gpu_suite.log 2>&1
echo "suite exit $? :: $(tail -1 $OUT/gpu_suite.log)" >> $OUT/summary.txt
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/summary.txt
timeout 900 python bench.py --steps 2 --warmup 3 --ops-out $OUT/ops.json > $OUT/bench.log 2> $OUT/bench.err
echo "bench exit $? :: $(tail -c 400 $OUT/bench.log)" >> $OUT/summary.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.log 2> $OUT/bench_ref.err
echo "bench_ref exit $? :: $(tail -c 300 $OUT/bench_ref.log)" >> $OUT/summary.txt
timeout 600 python bench.py --workload eval_ddpm1000_cfg --steps 1 --warmup 3 --no-cpu-baseline > $OUT/bench_eval.log 2> $OUT/bench_eval.err
echo "bench_eval exit $? :: $(head -c 300 $OUT/bench_eval.log)" >> $OUT/summary.txt
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "conv_kernel_variants" -p no:cacheprovider > $OUT/sanitizer_memcheck_conv.log 2>&1
echo "memcheck conv exit $? :: $(tail -2 $OUT/sanitizer_memcheck_conv.log | tr '\n' ' ')" >> $OUT/summary.txt
cat $OUT/summary.txt
