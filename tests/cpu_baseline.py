"""CPU baseline for bench.py: the reference's own compaction (oracle/_ref/ref_compact: the unmodified reference
compiled in this container; CompactionJobStats.elapsed_micros is the clock) on a bounded sample of a bench workload.
Falls back to the CPU oracle port when oracle/_ref is not present.  Bench/test infrastructure only."""
import json
import os
import shutil
import subprocess
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_compact")
KEY_MULT = 0x1000193


def _ops_file(path, w, n_total, seed=2):
    """write script with the bench generator's key layout: run r = every k-th key; newer runs are written later"""
    k, vlen = w["k"], w["vlen"]
    rng = np.random.default_rng(seed)
    with open(path, "wb") as f:
        f.write(b"B2OPS\0\0\1")
        for r in reversed(range(k)):  # run k-1 is the oldest
            idx = np.arange(r, n_total, k, dtype=np.uint64)
            n = idx.size
            kidx = idx.copy()
            if w["overlap"] > 0 and k > 1:
                sel = rng.random(n) < w["overlap"]
                kidx = np.where(sel, idx - r + ((r + 1) % k), kidx)
                kidx = np.minimum(kidx, n_total - 1)
                keep = np.ones(n, dtype=bool)
                keep[1:] = kidx[1:] != kidx[:-1]
                kidx = kidx[keep]
                n = kidx.size
            dele = rng.random(n) < w["del_frac"] if w["del_frac"] > 0 else np.zeros(n, dtype=bool)
            hi = (kidx * np.uint64(KEY_MULT)).astype(">u8")
            lo = ((kidx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0x7F4A7C15)) ^ (kidx << np.uint64(17))).astype(">u8")
            rec = np.zeros(n, dtype=[("op", "u1"), ("kl", "<u4"), ("vl", "<u4"), ("hi", ">u8"), ("lo", ">u8"), ("v", f"V{vlen}")])
            rec["op"], rec["kl"], rec["vl"], rec["hi"], rec["lo"] = 1, 16, vlen, hi, lo
            rec["v"] = rng.integers(0, 256, size=(n, vlen), dtype=np.uint8).view(f"V{vlen}").reshape(n)
            if dele.any():  # deletions have a shorter record: write them separately, keeping key order irrelevant for a memtable
                puts = rec[~dele]
                puts.tofile(f)
                d = np.zeros(int(dele.sum()), dtype=[("op", "u1"), ("kl", "<u4"), ("hi", ">u8"), ("lo", ">u8")])
                d["op"], d["kl"], d["hi"], d["lo"] = 2, 16, hi[dele], lo[dele]
                d.tofile(f)
            else:
                rec.tofile(f)
            f.write(b"\x03")
        f.write(b"\x00")


def usable_cores():
    """host threads this process may really use: CPU affinity, capped by the cgroup CPU quota when there is one"""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if q > 0:
                    n = min(n, max(1, int(q / per + 0.5)))
            break
        except Exception:
            continue
    return n


def run_sample(w, sample_bytes, threads):
    """threads > 1: that many concurrent reference processes, each compacting its own disjoint key range of the sample
    (the reference's own sub-compaction / dcompact parallelism is key-range partitioning); aggregate MB/s =
    total input KV bytes / slowest process' CompactionJobStats.elapsed_micros."""
    entry = 24 + w["vlen"]
    n_total = max(w["k"] * 1000, sample_bytes // entry)
    per = n_total // max(1, threads)
    desc = f"{threads} x ({w['k']} runs x {per // w['k']} entries) = {n_total * entry / 2**20:.0f} MiB raw KV of: {w['desc']}"
    if os.path.exists(REF_BIN):
        d = tempfile.mkdtemp(prefix="b200c_cpu_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            procs = []
            for t in range(threads):
                os.makedirs(os.path.join(d, str(t)))
                _ops_file(os.path.join(d, str(t), "ops.bin"), w, per, seed=2 + t)
            for t in range(threads):
                procs.append(subprocess.Popen([REF_BIN, os.path.join(d, str(t), "ops.bin"), os.path.join(d, str(t), "w"),
                                               "output_level=1", "max_subcompactions=1", "target_file_size=67108864", "copy=0",
                                               f"barrier_dir={d}", f"barrier_n={threads}"],
                                              stdout=subprocess.DEVNULL))
            for pr in procs:
                if pr.wait() != 0:
                    raise RuntimeError("ref_compact failed")
            kv, secs = 0, 0.0
            for t in range(threads):
                stt = json.load(open(os.path.join(d, str(t), "w", "manifest.json")))["stats"]
                kv += stt["total_input_raw_key_bytes"] + stt["total_input_raw_value_bytes"]
                secs = max(secs, stt["elapsed_micros"] / 1e6)
            return {"mbps": kv / secs / 1e6, "seconds": secs, "kind": "reference", "threads": threads,
                    "sample": desc + "; unmodified reference CompactionJob::Run (ProcessKeyValueCompaction), "
                                     "CompactionJobStats.elapsed_micros, one process per key range"}
        finally:
            shutil.rmtree(d, ignore_errors=True)
    # port: the CPU oracle on inputs it builds itself (single thread)
    import struct
    import helpers as H
    rng = np.random.default_rng(2)
    runs = []
    k = w["k"]
    for r in range(k):
        idx = np.arange(r, n_total, k, dtype=np.uint64)
        ents = []
        for i in idx.tolist():
            key = struct.pack(">QQ", (i * KEY_MULT) & (2**64 - 1), (i * 0x9E3779B97F4A7C15) & (2**64 - 1))
            ents.append((key + struct.pack("<Q", ((1 + i) << 8) | 1), rng.bytes(w["vlen"])))
        runs.append(H.oracle_build_sst(H.Params(), H.kvstream(ents)))
    p = H.Params(bottommost_level=w["bottommost"])
    t0 = time.perf_counter()
    H.oracle_compact(p, runs)
    secs = time.perf_counter() - t0
    return {"mbps": n_total * entry / secs / 1e6, "seconds": secs, "kind": "port", "threads": 1, "sample": desc + "; CPU oracle port"}


def run_subcompactions(w, sample_bytes, threads):
    """ONE reference job with max_subcompactions = threads (CompactionJob::GenSubcompactionBoundaries splits the key
    space at the input files' anchors, one thread per range).  Returns MB/s over CompactionJobStats.elapsed_micros and how many
    sub-compactions the reference really formed."""
    if not os.path.exists(REF_BIN):
        return None
    entry = 24 + w["vlen"]
    n_total = max(w["k"] * 1000, sample_bytes // entry)
    d = tempfile.mkdtemp(prefix="b200c_cpu_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        _ops_file(os.path.join(d, "ops.bin"), w, n_total, seed=2)
        subprocess.check_call([REF_BIN, os.path.join(d, "ops.bin"), os.path.join(d, "w"), "output_level=1",
                               f"max_subcompactions={threads}", "target_file_size=67108864", "copy=0"], stdout=subprocess.DEVNULL)
        man = json.load(open(os.path.join(d, "w", "manifest.json")))
        stt = man["stats"]
        kv = stt["total_input_raw_key_bytes"] + stt["total_input_raw_value_bytes"]
        secs = stt["elapsed_micros"] / 1e6
        return {"mbps": kv / secs / 1e6, "seconds": secs, "max_subcompactions": threads,
                "sub_compactions_formed": max(1, len(man.get("subcompactions", []))),
                "sample": f"one job, {w['k']} runs x {n_total // w['k']} entries = {n_total * entry / 2**20:.0f} MiB raw KV"}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def run_plugin_sample(w, sample_bytes):
    """The executor-plugin path end to end: ONE job of the reference DB (oracle/_ref/ref_compact_b200, executor=b200) -- the DB picks the
    inputs, RunRemote hands them to B200CompactionExecutor::Execute, which reads the files through the DB's FileSystem into pinned
    buffers, runs the job on the GPU, writes + syncs the output files; the DB renames and installs them.  The clock is
    CompactionJobStats.elapsed_micros (set by Execute: file reads, H2D, kernels, D2H, file writes, fsync).  The same job of the stock
    binary (local CPU compaction) is timed beside it.  Returns None when the binaries are missing."""
    b200_bin = os.path.join(ROOT, "oracle", "_ref", "ref_compact_b200")
    if not (os.path.exists(REF_BIN) and os.path.exists(b200_bin)):
        return None
    entry = 24 + w["vlen"]
    n_total = max(w["k"] * 1000, sample_bytes // entry)
    d = tempfile.mkdtemp(prefix="b200c_plugin_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    out = {"sample": f"one job of the reference DB, {w['k']} L0 files x {n_total // w['k']} entries = {n_total * entry / 2**20:.0f} MiB raw KV "
                     f"({w['desc']}), files on {'tmpfs' if d.startswith('/dev/shm') else 'disk'}; the process has run the same job once "
                     f"before the timed one (warm=1)"}
    try:
        _ops_file(os.path.join(d, "ops.bin"), w, n_total, seed=2)
        # warm=1: script + job once through a throw-away DB of the same process first -- a DB's second and later compactions
        # plugin_4_ranges: the same job with max_subcompactions=4 -- the plugin splits it into key ranges that share the uploaded inputs
        for name, binary, extra, subs in (("plugin", b200_bin, ["executor=b200", "warm=1"], 1), ("plugin_4_ranges", b200_bin, ["executor=b200", "warm=1"], 4),
                                          ("local", REF_BIN, ["warm=1"], 1)):
            best = None
            for rep in range(1):
                wd = os.path.join(d, f"{name}{rep}")
                subprocess.check_call([binary, os.path.join(d, "ops.bin"), wd, "output_level=1", f"max_subcompactions={subs}",
                                       "target_file_size=67108864", "copy=0"] + extra, stdout=subprocess.DEVNULL)
                man = json.load(open(os.path.join(wd, "manifest.json")))
                stt = man["stats"]
                kv = stt["total_input_raw_key_bytes"] + stt["total_input_raw_value_bytes"]
                secs = stt["elapsed_micros"] / 1e6
                if best is None or secs < best["seconds"]:
                    best = {"mbps": kv / secs / 1e6, "seconds": secs, "executor": man.get("executor", "local"),
                            "remote_compact_read_bytes": man.get("remote_compact_read_bytes", 0), "kv_bytes": kv}
                shutil.rmtree(wd, ignore_errors=True)
            out[name] = best
        return out
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":  # scaling probe: python tests/cpu_baseline.py 1 8 32 128
    import sys
    W = dict(k=8, run_bytes=256 << 20, vlen=32, overlap=0.0, del_frac=0.0, bottommost=False, desc="cfg2")
    print("usable cores:", usable_cores(), "cpu_count:", os.cpu_count())
    for t in [int(a) for a in sys.argv[1:]] or [1, usable_cores()]:
        r = run_sample(W, sample_bytes=t * (16 << 20), threads=t)
        print(f"threads={t:4d}  {r['mbps']:9.1f} MB/s  slowest job {r['seconds']*1e3:8.1f} ms  ({r['mbps']/t:7.1f} MB/s per thread)")
