#!/bin/bash
# tools/gpu_restamp.sh — after a change of csrc: the range / sub-job tests, then the ncu launch list the traffic stamp is taken from
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_subjobs.py tests/test_gpu_zlib.py -x -q) > gpurun_out/gputest.log 2>&1; tail -2 gpurun_out/gputest.log
bash tools/prof_launches.sh final
