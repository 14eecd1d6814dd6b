"""The bench workloads (BASELINE.json configs): plain data, importable without torch (bench.py --impl reference, tests)."""
# shared by bench.py and tests/test_gpu_fullsize.py so that the byte-exact
# oracle comparison runs on exactly the job the bench times
WORKLOADS = {
    # name: (k runs, raw KV bytes per run, value bytes, overlap, deletion fraction, bottommost)
    "cfg2": dict(k=8, run_bytes=256 << 20, vlen=32, overlap=0.0, del_frac=0.0, bottommost=False,
                 desc="8-way merge, 8x256MiB synthetic sorted runs, 16B keys / 32B values"),
    "cfg3": dict(k=16, run_bytes=256 << 20, vlen=256, overlap=0.3, del_frac=0.1, bottommost=True,
                 desc="16-way merge, 30% key overlap + 10% tombstones, 16B keys / 256B values"),
    "cfg5": dict(k=4, run_bytes=64 << 20, vlen=128, overlap=0.0, del_frac=0.0, bottommost=False,
                 desc="4-way x 64MiB sub-compaction, 16B keys / 128B values"),
    # BASELINE.json configs[4] as written: 64 independent sub-compactions, 8 per GPU (each a cfg5 job), all of a GPU's jobs in flight at once
    "cfg5x8": dict(base="cfg5", jobs=8, k=4, run_bytes=64 << 20, vlen=128, overlap=0.0, del_frac=0.0, bottommost=False,
                   desc="8 concurrent 4-way x 64MiB sub-compactions per GPU, 16B keys / 128B values"),
}
BENCH_JOB = dict(output_level=1, max_output_file_size=64 << 20, file_creation_times=[1700000000], first_file_number=1, db_id="bench",
                 db_session_id="BENCH", db_host_id="b200")
