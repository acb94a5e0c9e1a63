#!/bin/bash
# ncu evidence for profiles/: (1) the launch list of the bench command, (2) per-launch DRAM bytes / tensor-pipe activity of
# one whole CFG forward, (3) `--set full` captures of representative launches.  Every ncu command runs only after the
# same command line has exited 0 without ncu.
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --batch ${BENCH_BATCH:-1024} --steps 1 --warmup 3 --no-cpu-baseline --no-roofline"
PF="python tools/profile_forward.py --batch ${BENCH_BATCH:-1024}"
OPS=${PROFILE_OPS:-up_blocks.6.0.conv1,up_blocks.9.0.conv1,down_blocks.0.0.conv1,up_blocks.6.1.qkv,up_blocks.6.1.proj,up_blocks.9.0.conv1.0,attention,output.2}

$BENCH > $OUT/bench_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-2000} -c ${NCU_COUNT:-400} --csv \
    --log-file $OUT/launches.csv $BENCH > $OUT/ncu_launches.log 2>&1
echo "launch list exit $?" >> $OUT/summary.txt

$PF --mode forward --table-out $OUT/op_table.json > $OUT/pf_forward_plain.log 2>&1 &&
timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file $OUT/forward_metrics.csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed \
    $PF --mode forward > $OUT/ncu_forward.log 2>&1
echo "forward metrics exit $?" >> $OUT/summary.txt

$PF --mode ops --ops $OPS > $OUT/pf_ops_plain.log 2>&1 &&
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -f -o $OUT/prof_ops \
    $PF --mode ops --ops $OPS > $OUT/ncu_ops.log 2>&1
echo "full capture exit $?" >> $OUT/summary.txt
ls -la $OUT >> $OUT/summary.txt
tail -5 $OUT/summary.txt
