#!/bin/bash
# Runs on the GPU box (under gpurun): bench line, ncu launch list of the same command, one full ncu
# capture of the two hot kernels.  Usage: tools/gpu_profile.sh <tag> [workload-for-full-capture]
TAG=${1:-r1}
FULLW=${2:-syn12}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.csv
lscpu | grep -E "Model name|^CPU\(s\)" >> $OUT/${TAG}_gpu.csv
python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"; tail -c 3000 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
echo "ref rc=$?"; tail -c 1500 $OUT/${TAG}_bench_ref.json
# launch list of the same command (shares of the step, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_launches_run.log 2>&1
echo "ncu list rc=$?"
# full capture of the hot kernels on a smaller problem (ncu replays every launch ~40 times)
ncu --set full --clock-control none --import-source on -k regex:'fill_complete_kernel|fill_sorted_kernel|fill_kernel|spmv_rows' -c 3 \
    -o $OUT/${TAG}_full -f python bench.py --workload $FULLW --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_full_run.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT
