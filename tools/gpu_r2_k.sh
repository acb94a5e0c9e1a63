#!/bin/bash
# Round 2, run 14: ping-pong attention kernel (L = 256) -- parity, A/B against the one-warpgroup-per-tile kernel (UNet + DiT op
# tables), and an ncu --set full capture (with source) of the 1x1 GEMMs around it.
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/summary.txt; rm -f $OUT/eps_errors.txt
run() { local name=$1; shift; timeout 600 python -m pytest "$@" -q -m gpu --timeout 300 -p no:cacheprovider -x > $OUT/$name.log 2>&1; echo "$name exit $? :: $(tail -1 $OUT/$name.log)" >> $OUT/summary.txt; }
run attn tests/test_gpu_ops.py -k "attention"
run models tests/test_gpu_unet.py tests/test_gpu_dit.py tests/test_gpu_config1.py
timeout 900 python bench.py --steps 3 --warmup 3 --ops-out $OUT/ops.json --no-cpu-baseline > $OUT/bench.log 2> $OUT/bench.err
echo "bench exit $? :: $(head -c 250 $OUT/bench.log)" >> $OUT/summary.txt
DMC_ATTN_PP=0 timeout 900 python bench.py --steps 3 --warmup 3 --ops-out $OUT/ops_oldattn.json --no-cpu-baseline > $OUT/bench_oldattn.log 2> $OUT/bench_oldattn.err
echo "bench_oldattn exit $? :: $(head -c 250 $OUT/bench_oldattn.log)" >> $OUT/summary.txt
timeout 900 python bench.py --model dit --steps 3 --warmup 3 --ops-out $OUT/ops_dit.json --no-cpu-baseline > $OUT/bench_dit.log 2> $OUT/bench_dit.err
echo "bench_dit exit $? :: $(head -c 250 $OUT/bench_dit.log)" >> $OUT/summary.txt
DMC_ATTN_PP=0 timeout 900 python bench.py --model dit --steps 3 --warmup 3 --ops-out $OUT/ops_dit_oldattn.json --no-cpu-baseline > $OUT/bench_dit_oldattn.log 2> $OUT/bench_dit_oldattn.err
echo "bench_dit_oldattn exit $? :: $(head -c 250 $OUT/bench_dit_oldattn.log)" >> $OUT/summary.txt
# ncu --set full with source: the epilogue-bound 1x1 GEMMs and the new attention kernel (each command first without ncu)
PF="python tools/profile_forward.py --batch 1024"
OPS=up_blocks.6.1.qkv,attention,up_blocks.6.1.proj
$PF --mode ops --ops $OPS > $OUT/pf_ops_plain.log 2>&1 &&
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -f -o $OUT/prof_unet_ops \
    $PF --mode ops --ops $OPS > $OUT/ncu_unet_ops.log 2>&1
echo "ncu unet ops exit $?" >> $OUT/summary.txt
OPSD=blocks.5.qkv,blocks.5.fc1,blocks.5.fc2,blocks.5.out_proj
$PF --model dit --mode ops --ops $OPSD > $OUT/pf_dit_ops_plain.log 2>&1 &&
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -f -o $OUT/prof_dit_ops \
    $PF --model dit --mode ops --ops $OPSD > $OUT/ncu_dit_ops.log 2>&1
echo "ncu dit ops exit $?" >> $OUT/summary.txt
ls -la $OUT >> $OUT/summary.txt
cat $OUT/summary.txt
