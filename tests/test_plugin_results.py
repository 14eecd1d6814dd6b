"""The second half of the executor plugin on the CPU: what it does with a finished job.  oracle/_ref/ref_compact_mock links the real plugin
against a test double of the library (tests/native/mock_b200c.c) which, with B200C_MOCK_OUTPUTS set, "produces" the very files the
unmodified reference wrote for the same data in a separate local run (no compaction logic in the double).  The plugin then writes them
into its output directory, fills CompactionResults (FileMinMeta, CompactionJobStats) and the reference's RunRemote renames, re-opens and
installs them (db/compaction/compaction_job.cc:1019-1100).  The DB must end up exactly as after the local run: same files at the output
level, same per-file metadata, same statistics, same full-scan digest, and the job must have gone through the RunRemote branch.
Also with `max_subcompactions > 1`: the DB plans several sub-compactions, the executor answers with one result group, and RunRemote has to
cope with the different count (compaction_job.cc:986-1000)."""
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


def _canned(ref, d):
    """the local run's outputs + what a finished b200c job reports about them (b200c_file_meta, b200c_stats)"""
    lines = []
    for i, data in enumerate(ref["outputs"]):
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
    st = ref["manifest"]["stats"]
    lines.append("STATS " + " ".join(str(st[k]) for k in ("num_input_records", "num_output_records", "num_input_deletion_records",
                                                         "num_records_replaced", "num_expired_deletion_records", "total_input_raw_key_bytes",
                                                         "total_input_raw_value_bytes", "total_input_bytes", "total_output_bytes")) +
                 f" {len(ref['inputs'])}")
    with open(os.path.join(d, "meta.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


@pytest.mark.parametrize("name,extra", [("basic_bottommost", {}), ("snapshots", {}), ("varlen_keys", {}), ("cfg3_mini", {}),
                                        ("cfg2_mini", {}), ("crc32c_small_blocks", {}), ("cfg3_mini", dict(bloom_bits=10)), ("tiny", {}),
                                        ("cfg3_mini", dict(max_subcompactions=4)), ("cfg2_mini", dict(max_subcompactions=8))])
def test_plugin_hands_finished_outputs_to_run_remote(name, extra):
    ops, opts = S.ALL[name]()
    opts = dict(opts, **extra)
    want = H.run_reference(ops, **opts)  # the reference's own CPU path
    assert want["outputs"]
    with tempfile.TemporaryDirectory(prefix="b200c_canned_") as d:
        _canned(want, d)
        with tempfile.TemporaryDirectory(prefix="b200c_mockrun_") as w:
            with open(os.path.join(w, "ops.bin"), "wb") as f:
                f.write(ops.bytes())
            env = dict(os.environ, B200C_MOCK_OUTPUTS=d)
            args = [MOCK_BIN, os.path.join(w, "ops.bin"), os.path.join(w, "w"), "executor=b200"] + [f"{k}={v}" for k, v in opts.items()]
            r = subprocess.run(args, capture_output=True, text=True, env=env)
            assert r.returncode == 0, r.stderr[-2000:]
            gm = json.load(open(os.path.join(w, "w", "manifest.json")))
            got = [open(os.path.join(w, "w", "outputs" + m["name"]), "rb").read() for m in gm["outputs"]]
    wm = want["manifest"]
    assert gm["executor"] == "B200Compact" and gm["remote_compact_read_bytes"] > 0  # RunRemote branch, not a silent local run
    assert got == want["outputs"]  # the installed files are the ones handed over, in level order
    for k in ("size", "smallest_seqno", "largest_seqno", "num_entries", "num_deletions", "smallestkey", "largestkey"):
        assert [m[k] for m in gm["outputs"]] == [m[k] for m in wm["outputs"]], k
    assert (gm["scan_count"], gm["scan_digest"]) == (wm["scan_count"], wm["scan_digest"])
    for k in ("num_input_records", "num_output_records", "num_records_replaced", "num_expired_deletion_records",
              "num_input_deletion_records", "total_input_raw_key_bytes", "total_input_raw_value_bytes"):
        assert gm["stats"][k] == wm["stats"][k], k
