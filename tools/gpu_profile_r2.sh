#!/bin/bash
# Runs on the GPU box (under gpurun): bench lines of both arms, ncu launch list of the bench command, ncu --set full
# captures of the hot kernels of the complete-space path (config 3) and of the selected-space path (config 5 style).
# Usage: tools/gpu_profile_r2.sh <tag>
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.csv
lscpu | grep -E "Model name|^CPU\(s\)" >> $OUT/${TAG}_gpu.csv
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
echo "ref rc=$?"; tail -c 600 $OUT/${TAG}_bench_ref.json
# launch list of the same command (shares of the step, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --gate-rows 50 > $OUT/${TAG}_launches_run.log 2>&1
echo "ncu list rc=$?"
# full captures: complete-space fill + long-row SpMV at config 3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fill_complete_kernel|spmv_rows' -c 2 \
    -o $OUT/${TAG}_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras --gate-rows 50 > $OUT/${TAG}_full_run.log 2>&1
echo "ncu full cfg3 rc=$?"
# full captures: selected-space kernels on a config-5-style space of 2 M determinants
export PYCI_B200_CFG5=32,10,2000000
C5="python bench.py --workload cfg5 --steps 1 --warmup 0 --no-cpu-baseline --no-extras --gate-rows 20"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:join_rows_kernel -s 4 -c 1 -o $OUT/${TAG}_join -f $C5 > $OUT/${TAG}_join_run.log 2>&1
echo "ncu join rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^.*fill_kernel' -c 1 -o $OUT/${TAG}_fillsel -f $C5 > $OUT/${TAG}_fillsel_run.log 2>&1
echo "ncu fill_kernel rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_short_rows -s 3 -c 1 -o $OUT/${TAG}_spmvshort -f $C5 > $OUT/${TAG}_spmvshort_run.log 2>&1
echo "ncu spmv_short rc=$?"
ls -la $OUT | tail -20
