"""toplingdb_b200/native.py (the ctypes host mirror used by the GPU tests and bench.py) against the test double of the library
(tests/native/mock_b200c.c built as a shared object): the keyword arguments of CompactionJob must arrive in b200c_params as given --
scalars, snapshot and file-creation-time arrays, grandparent structs with their key bytes, key-range bounds, strings -- and the inputs in
call order.  Runs in a child process with the module's library path pointed at the double, so the real library of this process is left
alone.  CPU only."""
import json
import os
import subprocess
import sys
import textwrap

import helpers as H

ROOT = H.ROOT

CHILD = textwrap.dedent("""
    import json, os, sys
    sys.path.insert(0, {root!r})
    import toplingdb_b200.native as N
    N.LIB_PATH, N._lib = {so!r}, None
    import toplingdb_b200 as T
    job = T.CompactionJob(output_level=3, bottommost_level=True, max_output_file_size=123456, block_size=2048, block_restart_interval=8,
                          format_version=5, checksum="crc32c", snapshots=[5, 9, 1000], column_family_id=7, column_family_name="cf7",
                          db_id="the-db", db_session_id="SESSION", db_host_id="host", creation_time=11, oldest_key_time=12,
                          file_creation_times=[21, 22], first_file_number=900, compaction_filter="ttl", ttl=60, ttl_now=1700000000,
                          grandparents=[(b"a\\x00b", b"c", 10), (b"c", b"\\xff" * 16, 20)], level_compaction_dynamic_file_size=0,
                          max_compaction_bytes=777, target_output_file_size=61728, range_start=b"k1", range_end=b"", paranoid_file_checks=1,
                          bloom_millibits_per_key=9500)
    job.add_input(b"x" * 100, level=0, file_number=31)
    job.add_input(b"y" * 50, level=2, file_number=12)
    try:
        job.run()
        print("RAN")
    except T.B200cError as e:
        print("ERR", e.code, str(e))
    job.close()
""")


def test_keyword_arguments_arrive_in_the_c_struct(tmp_path):
    so = str(tmp_path / "libmock.so")
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-std=c11", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "native", "mock_b200c.c"), "-o", so])
    dump = str(tmp_path / "dump.jsonl")
    r = subprocess.run([sys.executable, "-c", CHILD.format(root=ROOT, so=so)], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, B200C_MOCK_DUMP=dump))
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.startswith("ERR 5 "), r.stdout  # the double refuses to run: B200C_ERR_NOT_SUPPORTED surfaces as B200cError
    j = json.loads(open(dump).read().splitlines()[0])
    assert (j["output_level"], j["bottommost_level"], j["max_output_file_size"], j["block_size"], j["block_restart_interval"]) == (3, 1, 123456, 2048, 8)
    assert (j["format_version"], j["checksum"], j["snapshots"]) == (5, 1, [5, 9, 1000])
    assert (j["column_family_id"], j["column_family_name"], j["db_id"], j["db_session_id"], j["db_host_id"]) == (7, "cf7", "the-db", "SESSION", "host")
    assert (j["creation_time"], j["oldest_key_time"], j["file_creation_times"], j["first_file_number"]) == (11, 12, [21, 22], 900)
    assert (j["compaction_filter"], j["ttl"], j["ttl_now"]) == (2, 60, 1700000000)
    assert j["grandparents"] == [{"smallestkey": b"a\x00b".hex(), "largestkey": b"c".hex(), "size": 10},
                                 {"smallestkey": b"c".hex(), "largestkey": (b"\xff" * 16).hex(), "size": 20}]
    assert (j["level_compaction_dynamic_file_size"], j["max_compaction_bytes"], j["target_output_file_size"]) == (0, 777, 61728)
    assert (j["has_range_start"], j["has_range_end"], j["paranoid_file_checks"], j["bloom_millibits_per_key"]) == (1, 1, 1, 9500)
    assert j["inputs"] == [{"level": 0, "file_number": 31, "len": 100, "mem_kind": 0}, {"level": 2, "file_number": 12, "len": 50, "mem_kind": 0}]


CHILD_RANGES = textwrap.dedent("""
    import json, os, sys
    sys.path.insert(0, {root!r})
    import toplingdb_b200.native as N
    N.LIB_PATH, N._lib = {so!r}, None
    import toplingdb_b200 as T
    parent = T.CompactionJob(output_level=2, block_size=1024, db_session_id="S", first_file_number=5)
    parent.add_input(b"x" * 64, level=0, file_number=1, deferred=True)
    bounds = parent.plan_ranges(4, min_range_bytes=10)
    print("BOUNDS", [b.hex() for b in bounds])
    parent.upload_by_ranges(bounds)
    subs = [parent.sub_job(range_start=a, range_end=b, first_file_number=100 * (i + 1))
            for i, (a, b) in enumerate(zip([None] + bounds, bounds + [None]))]
    for j in subs:
        j.close()
    parent.close()
""")


def test_sub_job_and_range_calls_marshal_their_keys(tmp_path):
    """plan_ranges / upload_by_ranges / sub_job: boundary keys with embedded zero bytes and different lengths travel as 16-byte slots +
    lengths; a sub-job inherits its parent's keyword arguments, takes its own range and file numbers; a deferred host input is added
    with B200C_MEM_HOST_DEFERRED"""
    so = str(tmp_path / "libmock.so")
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-std=c11", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "native", "mock_b200c.c"), "-o", so])
    dump = str(tmp_path / "dump.jsonl")
    k0, k1 = b"a\x00b", b"a\x00b\x00\x00c" + b"\xff" * 10
    r = subprocess.run([sys.executable, "-c", CHILD_RANGES.format(root=ROOT, so=so)], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, B200C_MOCK_DUMP=dump, B200C_MOCK_BOUNDARIES=f"{k0.hex()},{k1.hex()}"))
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.split("BOUNDS ")[1].strip() == str([k0.hex(), k1.hex()])
    jobs = [json.loads(line) for line in open(dump).read().splitlines()]
    subs = [j for j in jobs if j["has_range_start"] or j["has_range_end"]]
    parent = [j for j in jobs if j["inputs"]][0]
    assert parent["inputs"] == [{"level": 0, "file_number": 1, "len": 64, "mem_kind": 2}]
    assert [(s["range_start"], s["range_end"], s["first_file_number"]) for s in subs] == \
           [("", k0.hex(), 100), (k0.hex(), k1.hex(), 200), (k1.hex(), "", 300)]
    assert all((s["output_level"], s["block_size"], s["db_session_id"]) == (2, 1024, "S") for s in subs)
