#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/fuzz_gpu_vs_reference.py 240 4711 > gpurun_out/fuzz_gpu.log 2>&1; tail -5 gpurun_out/fuzz_gpu.log | cut -c1-600
python bench.py --impl plugin > gpurun_out/plugin2.json 2> gpurun_out/plugin2.err; cut -c1-900 gpurun_out/plugin2.json
