#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_2gpu.txt 2>&1
(time timeout 600 python -m pytest tests/test_gpu_subjobs.py -x -q) > gpurun_out/gputest_2gpu.log 2>&1; tail -4 gpurun_out/gputest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
python - <<P
import json
try:
    d = json.load(open("gpurun_out/bench_2gpu.json"))
    print(d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"], d["clocks"], d["config"].get("numa"))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/bench_2gpu.err").read()[-1500:])
P
