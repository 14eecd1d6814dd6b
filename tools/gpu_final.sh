#!/bin/bash
# tools/gpu_final.sh — the round's evidence run on one B200: tests, bench lines, ncu launch list, feature costs
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/final_gputest.log 2>&1; tail -3 gpurun_out/final_gputest.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; cut -c1-400 gpurun_out/final_bench.json
python bench.py --impl reference > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err
python bench.py --impl plugin > gpurun_out/final_plugin.json 2> gpurun_out/final_plugin.err
bash tools/prof_launches.sh final
python tools/zlib_cost.py > gpurun_out/zlib_cost.json 2> gpurun_out/zlib_cost.err; cut -c1-600 gpurun_out/zlib_cost.json
python tools/feature_cost.py > gpurun_out/feature_cost.json 2> gpurun_out/feature_cost.err
for w in cfg3 cfg5 cfg5x8; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/final_bench_$w.json 2> gpurun_out/final_bench_$w.err; cut -c1-300 gpurun_out/final_bench_$w.json; echo; done
