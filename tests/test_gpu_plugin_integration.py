"""End-to-end drop-in test (GPU box): the UNMODIFIED reference DB (oracle/_ref/libtoplingdb_ref.so) with
`compaction_executor_factory` = the product's B200 CompactionExecutor plugin (toplingdb_b200/plugin/, the mirror of
db/compaction/compaction_executor.h:153-185).  The reference's own CompactionJob::Run takes its RunRemote() branch
(db/compaction/compaction_job.cc:642-653,903-1100), the plugin feeds the job's input SSTs through the C ABI, and the
reference installs, re-opens and reads back the files the GPU wrote.  Compared against the same script run with the
reference's local CPU path: data blocks, index blocks and table properties byte for byte (only the fields that
identify the DB instance / wall clock differ between two reference runs), CompactionJobStats, and a digest of a full
DB scan done by the reference's own reader with checksum verification on."""
import os

import pytest

try:
    import torch  # noqa: F401  (page the CUDA libraries in at collection time)
except Exception:  # pragma: no cover
    torch = None

import helpers as H
import scenarios as S
import sstfmt

pytestmark = pytest.mark.gpu

# properties that legitimately differ between two runs of the reference itself (new DB instance, wall clock)
VOLATILE = {"rocksdb.creating.db.identity", "rocksdb.creating.session.identity", "rocksdb.creating.host.identity",
            "rocksdb.creation.time", "rocksdb.file.creation.time", "rocksdb.oldest.key.time",
            # on the RunRemote branch the executor numbers its files itself and the DB renames them to numbers it
            # allocates afterwards (compaction_job.cc:1022-1034), so this property is the worker-local number by design
            "rocksdb.original.file.number"}
CASES = [c for c in S.ALL if c != "long_keys"]


def _need_bins():
    if not (os.path.exists(H.REF_BIN) and os.path.exists(H.REF_B200_BIN)):
        pytest.fail("oracle/_ref/ref_compact(_b200) missing: run __graft_entry__.build() where /root/reference exists")


def _blocks(data):
    t = sstfmt.parse_sst(data)
    out = []
    for _, h in t["index"]:
        out.append(data[h[0]:h[0] + h[1] + 5])  # payload + type + checksum
    io, isz = t["footer"]["index"]
    return out, data[io:io + isz + 5], {k: v for k, v in t["properties"].items() if k not in VOLATILE}


@pytest.mark.parametrize("case", CASES)
def test_reference_db_compacts_through_b200_executor(case):
    _need_bins()
    ops, opts = S.ALL[case]()
    want = H.run_reference(ops, **opts)
    got = H.run_reference(ops, binary=H.REF_B200_BIN, executor="b200", **opts)
    gm, wm = got["manifest"], want["manifest"]
    assert gm["executor"] == "B200Compact" and wm["executor"] == "local"
    # the job really went through RunRemote -> plugin -> device (the ticker is only bumped on that branch)
    assert gm["remote_compact_read_bytes"] > 0, "compaction did not take the RunRemote/B200 branch"
    assert (gm["scan_count"], gm["scan_digest"]) == (wm["scan_count"], wm["scan_digest"])
    assert len(got["outputs"]) == len(want["outputs"])
    for k in ("smallest_seqno", "largest_seqno", "num_entries", "num_deletions", "smallestkey", "largestkey"):
        assert [m[k] for m in gm["outputs"]] == [m[k] for m in wm["outputs"]], k

    def size_without_number(m, data):  # the executor's job-unique rocksdb.original.file.number may be a wider varint (see VOLATILE)
        return m["size"] - len(sstfmt.parse_sst(data)["properties"]["rocksdb.original.file.number"])

    assert [size_without_number(m, d) for m, d in zip(gm["outputs"], got["outputs"])] == \
           [size_without_number(m, d) for m, d in zip(wm["outputs"], want["outputs"])]
    for g, w in zip(got["outputs"], want["outputs"]):
        gb, gi, gp = _blocks(g)
        wb, wi, wp = _blocks(w)
        assert gb == wb
        assert gi == wi
        assert gp == wp
    for k in H.STAT_KEYS:
        assert gm["stats"][k] == wm["stats"][k], k


def test_job_the_device_rejects_falls_back_to_the_reference_cpu_path():
    """user keys longer than the 16-byte device key columns: Execute() returns NotSupported, and with
    AllowFallbackToLocal() the reference runs the job itself (compaction_job.cc:648-651)."""
    _need_bins()
    ops, opts = S.ALL["long_keys"]()
    want = H.run_reference(ops, **opts)
    got = H.run_reference(ops, binary=H.REF_B200_BIN, executor="b200+fallback", **opts)
    assert (got["manifest"]["scan_count"], got["manifest"]["scan_digest"]) == (want["manifest"]["scan_count"], want["manifest"]["scan_digest"])
    assert [len(o) for o in got["outputs"]] == [len(o) for o in want["outputs"]]  # (the fallback runs the job locally: the DB's own numbers)
