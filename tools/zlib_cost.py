"""What compressed inputs cost on the device: the unmodified reference (oracle/_ref/ref_compact, built with -DZLIB) writes a job's input
files with kZlibCompression (word-soup values, compressible ~3x), the product compacts them (device-resident images, per-kernel events),
and the reference's own CompactionJob on the same files is timed beside it (CompactionJobStats.elapsed_micros, one thread).
Prints one JSON object; usage on the B200 box:  python tools/zlib_cost.py [entries_per_run] [runs] > gpurun_out/zlib_cost.json"""
import json
import os
import random
import statistics
import struct
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import helpers as H
    import sstfmt
    from gpu_harness import job_from_params
    per_run = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
    nruns = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    rnd = random.Random(12)
    words = [b"compaction", b"level", b"block", b"table", b"restart", b"varint", b"checksum", b"footer", b"index", b"merge", b"snapshot"]
    soup = b" ".join(rnd.choice(words) for _ in range(20000))
    ops = H.Ops()
    for r in range(nruns):
        for i in range(per_run):
            k = (i * nruns + r) * 0x9E3779B97F4A7C15 & ((1 << 64) - 1)
            o = rnd.randrange(0, len(soup) - 200)
            ops.put(struct.pack(">QQ", k, 0), soup[o:o + 100 + (k & 63)])
        ops.flush()
    # keys of a run must be written in any order (the memtable sorts); values: 100..163 bytes of text
    out = {}
    for comp in ("zlib", "none"):
        with tempfile.TemporaryDirectory(prefix="b200c_zlibcost_") as w:
            ref = H.run_reference(ops, workdir=w, target_file_size=64 << 20, output_level=1, input_compression=comp)
            p = H.params_from_reference(ref)
            in_bytes = sum(len(d) for d in ref["inputs"])
            st_ref = ref["manifest"]["stats"]
            kv = st_ref["total_input_raw_key_bytes"] + st_ref["total_input_raw_value_bytes"]
            job = job_from_params(p, output_mem="device", profile=1)
            keep = []
            for i, d in enumerate(ref["inputs"]):
                t = torch.frombuffer(bytearray(d), dtype=torch.uint8).cuda()
                keep.append(t)
                job.add_input(t, level=0, file_number=i)
            times, kt = [], {}
            for it in range(6):
                job.run()
                if it >= 2:
                    times.append(job.stats().total_us)
                    for name, us in job.kernel_times():
                        kt.setdefault(name, []).append(us)
            same = job.outputs() == ref["outputs"]
            job.close()
            us = statistics.mean(times)
            out[comp] = {"input_sst_bytes": in_bytes, "input_kv_bytes": kv, "entries": st_ref["num_input_records"],
                         "device_us": round(us, 1), "device_kv_MB_per_s": round(kv / us, 1),
                         "kernels_us": {k: round(statistics.mean(v), 1) for k, v in sorted(kt.items(), key=lambda x: -statistics.mean(x[1]))[:8]},
                         "outputs_equal_reference": same,
                         "reference_cpu_us_one_thread": st_ref.get("elapsed_micros"),
                         "reference_cpu_kv_MB_per_s": round(kv / st_ref["elapsed_micros"], 1) if st_ref.get("elapsed_micros") else None}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
