#!/bin/bash
# tools/gpu_ab.sh — one GPU call: tests, then the bench job under several library variants / knobs (device-resident part only)
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e"
run() { tag=$1; shift; env "$@" $B > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err; python - <<P
import json
try:
    d = json.load(open("gpurun_out/ab_$tag.json"))
    print("$tag", d["ms_per_step"], d["stage_us"], {k["name"]: k["us"] for k in d["kernels"][:9]}, d["output_digest"]["matches_oracle"])
except Exception as e:
    print("$tag", "FAILED", e)
P
}
(time timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_grandparents.py tests/test_gpu_parity.py tests/test_gpu_option_matrix.py tests/test_gpu_subcompactions.py tests/test_gpu_bloom.py -x -q) > gpurun_out/gputest.log 2>&1; tail -4 gpurun_out/gputest.log
run r8 X=1
run r0 B200C_EMIT_RESERVE=0
run r16 B200C_EMIT_RESERVE=16
run r32 B200C_EMIT_RESERVE=32
