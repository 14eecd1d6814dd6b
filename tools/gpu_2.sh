#!/bin/bash
# tools/gpu_2.sh — what ran on the two-GPU box (gpurun --gpus 2): BASELINE configs[4] as written, one rank per GPU under torchrun
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload cfg5x8 --steps 5 --warmup 3 > gpurun_out/bench_cfg5x8_2gpu.json 2> gpurun_out/bench_cfg5x8_2gpu.err
python - <<P
import json
try:
    d = json.load(open("gpurun_out/bench_cfg5x8_2gpu.json"))
    print(d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"], d["clocks"], d["config"]["parallelism"])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/bench_cfg5x8_2gpu.err").read()[-1500:])
P
