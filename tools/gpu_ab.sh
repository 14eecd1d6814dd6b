#!/bin/bash
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_zlib.py -x -q) > gpurun_out/gputest.log 2>&1; tail -4 gpurun_out/gputest.log
python tools/zlib_cost.py > gpurun_out/zlib_cost4.json 2> gpurun_out/zlib_cost4.err; cut -c1-700 gpurun_out/zlib_cost4.json
