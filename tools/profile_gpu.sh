#!/bin/bash
# Run on the GPU box (gpurun): launch list + one ncu --set full capture of a steady-state cycle.
# usage: tools/profile_gpu.sh <config> <tag>
cfg=${1:-3}; tag=${2:-r01}
mkdir -p gpurun_out
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${tag}_launches_cfg${cfg}.csv python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_launches_cfg${cfg}.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:^k_ -s 40 -c 10 \
    -o gpurun_out/${tag}_full_cfg${cfg} -f python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_full_cfg${cfg}.log 2>&1
ls -la gpurun_out/ | tail -8
