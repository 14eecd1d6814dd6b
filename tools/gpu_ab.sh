#!/bin/bash
mkdir -p gpurun_out
for n in 16 4 32; do
python bench.py --no-cpu-baseline --steps 3 --pipeline-ranges $n > gpurun_out/bench_p$n.json 2> gpurun_out/bench_p$n.err; python - <<P
import json
d = json.load(open("gpurun_out/bench_p$n.json"))
print($n, d["ms_per_step"], d["e2e"]["single_job_ms"], d["e2e"]["single_job_pipelined"], d["roofline"]["traffic"])
P
done
