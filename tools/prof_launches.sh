#!/bin/bash
# ncu launch list (time + DRAM bytes per launch) of a short bench run; summarise here with tools/ncu_launch_summary.py
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_$TAG.log 2>&1
tail -2 gpurun_out/launches_$TAG.log | cut -c1-200
