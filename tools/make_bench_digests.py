#!/usr/bin/env python
"""Writes tests/golden/bench_digests.json: sha256 digests of the CPU ORACLE's output files for the jobs bench.py times (rank 0, scale 1).
Needs a GPU only because the synthetic inputs are generated and staged there (toplingdb_b200.synth); the digests themselves come from
oracle/liboracle.so.  bench.py compares its own outputs with these digests (it never imports the oracle);
tests/test_gpu_fullsize.py re-derives them on every GPU test run and fails when the committed file is stale.

    gpurun -- 'python tools/make_bench_digests.py gpurun_out/bench_digests.json'   then copy the file to tests/golden/"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import helpers as H
    from toplingdb_b200 import synth
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "bench_digests.json")
    res = {}
    for workload in (sys.argv[2:] or ["cfg2", "cfg5", "cfg3"]):
        w = synth.WORKLOADS[workload]
        images, _ = synth.stage_bench_inputs(workload, rank=0, scale=1.0)
        host = [t.cpu().numpy().tobytes() for t in images]
        del images
        torch.cuda.empty_cache()
        p = H.Params(bottommost_level=bool(w["bottommost"]), creation_time=0, **synth.BENCH_JOB)
        files, _, st = H.oracle_compact(p, host)
        res[f"{workload}:rank0:scale1.0"] = synth.files_digest(files)
        print(workload, len(files), "files", st.num_output_records, "records", res[f"{workload}:rank0:scale1.0"], flush=True)
        os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
        json.dump(res, open(out, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
