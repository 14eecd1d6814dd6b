#!/bin/bash
# tools/gpu_ab.sh — one GPU call: tests, then the bench job under several library variants / knobs (device-resident part only)
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e"
run() { tag=$1; shift; env "$@" $B > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err; python - <<P
import json
try:
    d = json.load(open("gpurun_out/ab_$tag.json"))
    print("$tag", d["ms_per_step"], {k["name"]: k["us"] for k in d["kernels"][:9]}, d["output_digest"]["matches_oracle"])
except Exception as e:
    print("$tag", "FAILED", e)
P
}
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/gputest.log 2>&1; tail -4 gpurun_out/gputest.log
run new X=1
run base B200C_LIB=$PWD/tools/scratch/libb200c_base.so
run chunk2 B200C_PART_CHUNK=2
run chunk3 B200C_PART_CHUNK=3
run trace B200C_LIB=$PWD/tools/scratch/libb200c_trace.so
grep -h "stitch trace" gpurun_out/ab_trace.err | tail -2
