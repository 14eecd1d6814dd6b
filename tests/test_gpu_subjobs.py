"""One job over several key ranges ON the device (GPU box): b200c_job_plan_ranges + b200c_job_create_sub of the C ABI and the
executor plugin's use of them (CompactionJob::Prepare / GenSubcompactionBoundaries, db/compaction/compaction_job.cc:264-281,465-640;
CompactionResults::output_files[sub], compaction_executor.h:120-158).
 (a) library: the planned boundaries cut the inputs into ranges of comparable size; every sub-job (shared device copies of the inputs,
     run concurrently from threads) writes byte for byte what the CPU oracle writes for its range, and the ranges partition the job;
 (b) plugin: the UNMODIFIED reference DB with `max_subcompactions = 4` hands the job to the B200 executor, which splits it; the DB
     installs all result groups and ends up with the same contents (full-scan digest, entry and file statistics) as after the same job on
     the reference's own CPU path, and the level is sorted and non-overlapping."""
import os
import random
import struct
import subprocess
import tempfile
import threading
import json

import pytest

try:
    import torch  # noqa: F401
except Exception:  # pragma: no cover
    torch = None

import helpers as H
import scenarios as S
import sstfmt

pytestmark = pytest.mark.gpu


def _seeded(seed, nruns=6, n=30000, vlen=40):
    rnd = random.Random(seed)
    runs, seq = [], 1
    universe = n * nruns // 2
    for _ in range(nruns):
        run = []
        for kk in sorted(rnd.sample(range(universe), n)):
            t = 0 if rnd.random() < 0.1 else 1
            run.append((struct.pack(">QQ", 9, kk * 5) + struct.pack("<Q", (seq << 8) | t), b"" if t == 0 else rnd.randbytes(vlen)))
            seq += 1
        runs.append(run)
    return [H.oracle_build_sst(H.Params(), H.kvstream(r)) for r in reversed(runs)]


@pytest.mark.parametrize("mode", ["host", "device", "pipelined"])
def test_sub_jobs_share_inputs_and_match_the_oracle(mode):
    """pipelined: the host inputs are uploaded in key order (b200c_job_upload_by_ranges) and every sub-job starts when the blocks of its
    own range have arrived -- the device may still be receiving the later ranges while the first ones are compacted"""
    from gpu_harness import job_from_params
    inputs = _seeded(5)
    p = H.Params(output_level=1, bottommost_level=True, max_output_file_size=256 << 10, file_creation_times=[3])
    parent = job_from_params(p)
    keep = []
    for i, data in enumerate(inputs):
        if mode == "device":
            t = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
            keep.append(t)
            parent.add_input(t, level=0, file_number=i)
        else:
            parent.add_input(data, level=0, file_number=i, deferred=mode == "pipelined")
    bounds = parent.plan_ranges(4, min_range_bytes=64 << 10)
    assert len(bounds) == 3 and bounds == sorted(bounds) and len(set(bounds)) == 3
    assert parent.plan_ranges(1) == [] and len(parent.plan_ranges(64, min_range_bytes=1 << 40)) == 0  # never smaller than min_range_bytes
    if mode == "pipelined":
        parent.upload_by_ranges(bounds)
    ranges = list(zip([None] + bounds, bounds + [None]))
    subs = [parent.sub_job(range_start=a, range_end=b, first_file_number=1000 * (i + 1)) for i, (a, b) in enumerate(ranges)]
    errs = []

    def run(j):
        try:
            j.run()
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    ths = [threading.Thread(target=run, args=(j,)) for j in subs]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    assert not errs, errs
    total_in = total_out = 0
    sizes = []
    for (a, b), j, i in zip(ranges, subs, range(len(subs))):
        q = H.Params(output_level=1, bottommost_level=True, max_output_file_size=256 << 10, file_creation_times=[3], range_start=a, range_end=b,
                     first_file_number=1000 * (i + 1))
        want, _, wst = H.oracle_compact(q, inputs)
        got = j.outputs()
        assert [len(f) for f in got] == [len(f) for f in want], (a, b)
        assert got == want, (a, b)
        st = j.stats()
        assert (st.num_input_records, st.num_output_records) == (wst.num_input_records, wst.num_output_records)
        total_in += st.num_input_records
        total_out += st.num_output_records
        sizes.append(sum(len(f) for f in got))
    whole, _, wst = H.oracle_compact(H.Params(output_level=1, bottommost_level=True, max_output_file_size=256 << 10, file_creation_times=[3]), inputs)
    assert (total_in, total_out) == (wst.num_input_records, wst.num_output_records)  # the ranges partition the job
    assert max(sizes) < 3 * max(1, min(sizes)), sizes  # about equal input bytes per range
    for j in subs:
        j.close()
    parent.close()


@pytest.mark.parametrize("devices", ["0", "0,1"])
def test_reference_db_splits_a_job_over_ranges_through_the_b200_executor(devices):
    """devices "0,1": the ranges of ONE job are dealt to two GPUs (B200CompactOptions::devices); needs a box with two of them"""
    if not (os.path.exists(H.REF_BIN) and os.path.exists(H.REF_B200_BIN)):
        pytest.fail("oracle/_ref/ref_compact(_b200) missing: run __graft_entry__.build() where /root/reference exists")
    import toplingdb_b200 as T
    ndev = len(devices.split(","))
    if T.device_count() < ndev:
        pytest.skip(f"needs {ndev} CUDA devices")
    ops, opts = S.ALL["cfg2_mini"]()
    opts = dict(opts, max_subcompactions=4)
    want = H.run_reference(ops, **opts)
    with tempfile.TemporaryDirectory(prefix="b200c_subjobs_") as w:
        with open(os.path.join(w, "ops.bin"), "wb") as f:
            f.write(ops.bytes())
        env = dict(os.environ, B200C_PLUGIN_TRACE="1")
        args = [H.REF_B200_BIN, os.path.join(w, "ops.bin"), os.path.join(w, "w"), "executor=b200", f"b200_devices={devices}"] + \
               [f"{k}={v}" for k, v in opts.items()]
        r = subprocess.run(args, capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr[-3000:]
        gm = json.load(open(os.path.join(w, "w", "manifest.json")))
        got = [open(os.path.join(w, "w", "outputs" + m["name"]), "rb").read() for m in gm["outputs"]]
    wm = want["manifest"]
    assert "split into" in r.stderr and f"key ranges over {ndev} device(s)" in r.stderr, r.stderr[-2000:]
    assert gm["executor"] == "B200Compact" and gm["remote_compact_read_bytes"] > 0
    assert (gm["scan_count"], gm["scan_digest"]) == (wm["scan_count"], wm["scan_digest"])
    # same entries, in order, whatever the file cuts of the ranges are
    ge = [e for f in got for e in sstfmt.parse_sst(f)["entries"]]
    we = [e for f in want["outputs"] for e in sstfmt.parse_sst(f)["entries"]]
    assert ge == we
    prev = None
    for m in gm["outputs"]:  # a sorted, non-overlapping level
        assert prev is None or bytes.fromhex(prev) < bytes.fromhex(m["smallestkey"])  # (user keys in the manifest)
        prev = m["largestkey"]
    for k in ("num_input_records", "num_output_records", "num_records_replaced", "num_expired_deletion_records",
              "num_input_deletion_records", "total_input_raw_key_bytes", "total_input_raw_value_bytes"):
        assert gm["stats"][k] == wm["stats"][k], k
