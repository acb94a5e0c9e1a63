#!/bin/bash
# Round 2, run 40: 4096 images per launch (UNet.max_images_per_launch): suite, bench line + op table, ncu per-launch metrics of one
# forward at that size (roofline.traffic)
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt; rm -f $OUT/eps_errors.txt
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider > $OUT/gpu_suite.log 2>&1
echo "suite exit $? :: $(tail -1 $OUT/gpu_suite.log)" >> $OUT/summary.txt
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --ops-out $OUT/ops_4096.json --no-cpu-baseline > $OUT/bench.log 2> $OUT/bench.err
echo "bench exit $? :: $(head -c 200 $OUT/bench.log)" >> $OUT/summary.txt
PF="python tools/profile_forward.py --batch 2048"
$PF --mode forward --table-out $OUT/op_table.json > $OUT/pf_forward_plain.log 2>&1 &&
timeout 600 ncu --profile-from-start off --clock-control none --csv --log-file $OUT/forward_metrics.csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed \
    $PF --mode forward > $OUT/ncu_forward.log 2>&1
echo "forward metrics exit $?" >> $OUT/summary.txt
cat $OUT/summary.txt
