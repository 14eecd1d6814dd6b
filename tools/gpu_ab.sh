#!/bin/bash
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/final_gputest.log 2>&1; tail -3 gpurun_out/final_gputest.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; python - <<P
import json
d = json.load(open("gpurun_out/final_bench.json"))
k = {x["name"]: x["us"] for x in d["kernels"]}
print(d["ms_per_step"], d["value"], d["stage_us"], "stitch", k.get("~encode.stitch"), "tables", k.get("encode.tables"), d["e2e"], d["roofline"]["frac"], d["clocks"], d["output_digest"]["matches_oracle"])
P
bash tools/prof_launches.sh final
