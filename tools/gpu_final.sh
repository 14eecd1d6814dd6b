#!/bin/bash
# tools/gpu_final.sh — the round's evidence run on one B200: tests, the bench line, ncu launch list, full-size ncu pages
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/final_gputest.log 2>&1; tail -3 gpurun_out/final_gputest.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; cut -c1-300 gpurun_out/final_bench.json; echo
bash tools/prof_launches.sh final
bash tools/prof_full.sh final 'block_decode_fused|merge_tiles|encode_emit|encode_tables' 20 4 > gpurun_out/prof_full.log 2>&1; tail -3 gpurun_out/prof_full.log
python tools/feature_cost.py > gpurun_out/feature_cost.json 2> gpurun_out/feature_cost.err
