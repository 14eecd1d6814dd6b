"""One job over several key ranges in the executor plugin (CPU, against the test double of the library).

The reference splits a job whose `max_subcompactions > 1` into key ranges that run on threads (CompactionJob::Prepare /
GenSubcompactionBoundaries, db/compaction/compaction_job.cc:264-281,465-640) and RunRemote accepts any number of result groups from an
executor (`CompactionResults::output_files[sub]`, compaction_job.cc:986-1000).  The plugin asks the library for boundaries
(b200c_job_plan_ranges), creates one sub-job per range over the SAME inputs (b200c_job_create_sub), runs them concurrently and answers
with one result group per range.  Checked here without a GPU:
  * what reaches the C ABI: ranges [None, k0) [k0, k1) [k1, None), job-unique file numbers per range, inputs added once (to the parent);
  * the result path: the test double "produces", per range, the files the unmodified reference wrote for that key range in a local run;
    RunRemote must install them so that the DB is exactly what the local run left behind."""
import json
import os
import struct
import subprocess
import tempfile

import pytest

import helpers as H
import scenarios as S
import sstfmt

MOCK_BIN = os.path.join(H.ROOT, "oracle", "_ref", "ref_compact_mock")
pytestmark = pytest.mark.skipif(not os.path.exists(MOCK_BIN), reason="oracle/_ref/ref_compact_mock not built (needs /root/reference)")


def _run_mock(ops, opts, env_extra, executor):
    with tempfile.TemporaryDirectory(prefix="b200c_mocksub_") as w:
        with open(os.path.join(w, "ops.bin"), "wb") as f:
            f.write(ops.bytes())
        dump = os.path.join(w, "dump.jsonl")
        env = dict(os.environ, B200C_MOCK_DUMP=dump, B200C_PLUGIN_TRACE="1", **env_extra)
        args = [MOCK_BIN, os.path.join(w, "ops.bin"), os.path.join(w, "w"), f"executor={executor}"] + [f"{k}={v}" for k, v in opts.items()]
        r = subprocess.run(args, capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        jobs = [json.loads(line) for line in open(dump)] if os.path.exists(dump) else []
        man = json.load(open(os.path.join(w, "w", "manifest.json")))
        outs = [open(os.path.join(w, "w", "outputs" + m["name"]), "rb").read() for m in man["outputs"]]
    return jobs, man, outs, r.stderr


def test_ranges_reach_the_c_abi():
    ops, opts = S.ALL["cfg3_mini"]()
    opts = dict(opts, max_subcompactions=4)
    k0, k1 = b"\x00" * 7 + b"\x40", b"\x00" * 7 + b"\x80" + b"\x01"
    jobs, man, _, err = _run_mock(ops, opts, dict(B200C_MOCK_BOUNDARIES=f"{k0.hex()},{k1.hex()}"), "b200+fallback")
    want_inputs = sorted(m["file_number"] for m in man["inputs"])
    mine = [j for j in jobs if j["output_level"] == man["output_level"]]
    parents = [j for j in mine if sorted(i["file_number"] for i in j["inputs"]) == want_inputs]
    subs = [j for j in mine if not j["inputs"] and (j["has_range_start"] or j["has_range_end"])]
    assert len(parents) >= 1 and len(subs) == 3, (len(parents), len(subs))
    assert "split into 3 key ranges over 1 device(s)" in err
    assert [(s["has_range_start"], s["range_start"], s["has_range_end"], s["range_end"]) for s in subs] == \
           [(0, "", 1, k0.hex()), (1, k0.hex(), 1, k1.hex()), (1, k1.hex(), 0, "")]
    nums = [s["first_file_number"] for s in subs]
    assert len(set(nums)) == 3 and all(b - a >= 1 << 14 for a, b in zip(nums, nums[1:]))  # room for 16384 files per range
    for s in subs:  # a range is the same job otherwise
        for k in ("bottommost_level", "max_output_file_size", "block_size", "format_version", "checksum", "snapshots", "db_session_id"):
            assert s[k] == parents[-1][k], k


def test_ranges_are_dealt_to_the_listed_devices():
    """B200CompactOptions::devices = {0, 1}: every device in use gets its own parent (= its own copy of the inputs), the ranges go to
    the devices round-robin"""
    ops, opts = S.ALL["cfg3_mini"]()
    opts = dict(opts, max_subcompactions=4, b200_devices="0,1")
    k0, k1 = b"\x00" * 7 + b"\x40", b"\x00" * 7 + b"\x80"
    jobs, man, _, err = _run_mock(ops, opts, dict(B200C_MOCK_BOUNDARIES=f"{k0.hex()},{k1.hex()}"), "b200+fallback")
    want_inputs = sorted(m["file_number"] for m in man["inputs"])
    mine = [j for j in jobs if j["output_level"] == man["output_level"]]
    parents = [j for j in mine if sorted(i["file_number"] for i in j["inputs"]) == want_inputs and not (j["has_range_start"] or j["has_range_end"])]
    subs = [j for j in mine if not j["inputs"] and (j["has_range_start"] or j["has_range_end"])]
    assert "split into 3 key ranges over 2 device(s)" in err
    assert sorted(p["device"] for p in parents[-2:]) == [0, 1]
    assert [s["device"] for s in subs] == [0, 1, 0]


def test_a_job_that_must_not_be_split_stays_whole():
    ops, opts = S.ALL["cfg3_mini"]()
    jobs, man, _, err = _run_mock(ops, dict(opts, max_subcompactions=1), dict(B200C_MOCK_BOUNDARIES="0000000000000040"), "b200+fallback")
    assert "split into" not in err and not [j for j in jobs if j["has_range_start"] or j["has_range_end"]]
    jobs, man, _, err = _run_mock(ops, dict(opts, max_subcompactions=4, b200_subs=1), dict(B200C_MOCK_BOUNDARIES="0000000000000040"),
                                  "b200+fallback")
    assert "split into" not in err  # B200CompactOptions::max_subcompactions = 1 switches the split off


def _canned_group(files, stats_line, d):
    os.makedirs(d)
    lines = []
    for i, data in enumerate(files):
        name = f"{i:06d}.sst"
        with open(os.path.join(d, name), "wb") as f:
            f.write(data)
        t = sstfmt.parse_sst(data)
        pr = t["properties"]
        seqs = [struct.unpack("<Q", ik[-8:])[0] >> 8 for ik, _ in t["entries"]]
        u = lambda k: sstfmt.prop_u64(pr, k)
        lines.append(" ".join([name, t["entries"][0][0].hex(), t["entries"][-1][0].hex(), str(min(seqs)), str(max(seqs)),
                               str(u("rocksdb.num.entries")), str(u("rocksdb.deleted.keys")), str(u("rocksdb.raw.key.size")),
                               str(u("rocksdb.raw.value.size")), str(u("rocksdb.num.data.blocks")), str(u("rocksdb.data.size")),
                               str(t["footer"]["index"][1])]))
    lines.append(stats_line)
    with open(os.path.join(d, "meta.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


@pytest.mark.parametrize("name", ["cfg2_mini", "cfg3_mini"])
def test_result_groups_of_the_ranges_are_installed_in_key_order(name):
    ops, opts = S.ALL[name]()
    opts = dict(opts, max_subcompactions=4)
    want = H.run_reference(ops, **dict(opts, max_subcompactions=1))  # one range: the level's files in key order
    files = want["outputs"]
    if len(files) < 2:
        pytest.skip("scenario writes a single output file")
    cut = len(files) // 2
    boundary = sstfmt.parse_sst(files[cut])["entries"][0][0][:-8]  # first user key of the second group
    assert len(boundary) <= 16
    st = want["manifest"]["stats"]
    keys = ("num_input_records", "num_output_records", "num_input_deletion_records", "num_records_replaced", "num_expired_deletion_records",
            "total_input_raw_key_bytes", "total_input_raw_value_bytes")
    nin = len(want["inputs"])
    with tempfile.TemporaryDirectory(prefix="b200c_cannedsub_") as d:
        # the counters of the ranges add up: everything in the first group, nothing in the second; both read the same input files
        _canned_group(files[:cut], "STATS " + " ".join(str(st[k]) for k in keys) + f" {st['total_input_bytes']} {st['total_output_bytes']} {nin}",
                      os.path.join(d, "sub0"))
        _canned_group(files[cut:], "STATS " + " ".join("0" for _ in keys) + f" {st['total_input_bytes']} 0 {nin}", os.path.join(d, "sub1"))
        jobs, gm, got, err = _run_mock(ops, opts, dict(B200C_MOCK_OUTPUTS=d, B200C_MOCK_BOUNDARIES=boundary.hex()), "b200")
    assert "split into 2 key ranges" in err
    wm = want["manifest"]
    assert gm["executor"] == "B200Compact" and gm["remote_compact_read_bytes"] > 0
    assert got == files  # both groups installed, in key order
    for k in ("size", "smallest_seqno", "largest_seqno", "num_entries", "num_deletions", "smallestkey", "largestkey"):
        assert [m[k] for m in gm["outputs"]] == [m[k] for m in wm["outputs"]], k
    assert (gm["scan_count"], gm["scan_digest"]) == (wm["scan_count"], wm["scan_digest"])
    for k in keys:
        assert gm["stats"][k] == wm["stats"][k], k
