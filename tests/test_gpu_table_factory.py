"""The TableFactory side of the boundary (toplingdb_b200/plugin/b200_table_factory.{h,cc}): the UNMODIFIED reference DB, configured with
B200TableFactory as its table factory, writes every table -- memtable flushes (FlushJob -> BuildTable) and the outputs of its own
CompactionJob -- through B200TableBuilder, i.e. through b200c_job_encode_kv on the GPU.  The files must be the ones the stock
BlockBasedTableFactory writes for the same operations: data blocks, index block and properties (minus the per-run identity fields),
the same FileMetaData, the same statistics; the file cuts of the compaction prove that FileSize() advances exactly as the stock
builder's does (compaction_outputs.cc:277)."""
import os

import pytest

try:
    import torch  # noqa: F401
except Exception:  # pragma: no cover
    torch = None

import helpers as H
import scenarios as S
import sstfmt

pytestmark = pytest.mark.gpu

VOLATILE = {"rocksdb.creating.db.identity", "rocksdb.creating.session.identity", "rocksdb.creating.host.identity",
            "rocksdb.creation.time", "rocksdb.file.creation.time", "rocksdb.oldest.key.time"}
CASES = [c for c in S.ALL if c != "long_keys"]


def _need_bins():
    if not (os.path.exists(H.REF_BIN) and os.path.exists(H.REF_B200_BIN)):
        pytest.fail("oracle/_ref/ref_compact(_b200) missing: run __graft_entry__.build() where /root/reference exists")


def _parts(data):
    t = sstfmt.parse_sst(data)
    ft = sstfmt.parse_footer(data)
    io, isz = ft["index"]
    return data[:io], data[io:io + isz + 5], {k: v for k, v in t["properties"].items() if k not in VOLATILE}


def _same_tables(got, want):
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert len(g) == len(w)
        gd, gi, gp = _parts(g)
        wd, wi, wp = _parts(w)
        assert gd == wd, "data (+ filter) blocks differ"
        assert gi == wi, "index block differs"
        assert gp == wp, "table properties differ"


@pytest.mark.parametrize("case", CASES)
def test_reference_db_writes_its_tables_through_the_b200_table_factory(case):
    _need_bins()
    ops, opts = S.ALL[case]()
    want = H.run_reference(ops, **opts)
    got = H.run_reference(ops, binary=H.REF_B200_BIN, table_factory="b200+nofallback", **opts)
    gm, wm = got["manifest"], want["manifest"]
    assert gm["table_factory"] == "B200BlockBasedTable" and gm["b200_device_tables"] >= len(got["inputs"]) + len(got["outputs"]) > 0
    assert (gm["scan_count"], gm["scan_digest"]) == (wm["scan_count"], wm["scan_digest"])
    _same_tables(got["inputs"], want["inputs"])    # flushed by FlushJob through the B200 builder
    _same_tables(got["outputs"], want["outputs"])  # written by the reference's own CompactionJob through the B200 builder
    for k in ("size", "smallest_seqno", "largest_seqno", "num_entries", "num_deletions", "smallestkey", "largestkey"):
        assert [m[k] for m in gm["outputs"]] == [m[k] for m in wm["outputs"]], k
    for k in H.STAT_KEYS:
        assert gm["stats"][k] == wm["stats"][k], k


def test_tables_outside_the_device_rule_set_are_written_by_the_stock_builder():
    """user keys longer than 16 bytes: B200TableBuilder replays its records into the reference's BlockBasedTableBuilder"""
    _need_bins()
    ops, opts = S.ALL["long_keys"]()
    want = H.run_reference(ops, **opts)
    got = H.run_reference(ops, binary=H.REF_B200_BIN, table_factory="b200", **opts)
    assert got["manifest"]["b200_fallback_tables"] > 0
    _same_tables(got["inputs"], want["inputs"])
    _same_tables(got["outputs"], want["outputs"])


def test_both_halves_of_the_boundary_together():
    """B200TableFactory as the table factory AND B200Compact as the executor: flushes through the builder, the compaction through
    RunRemote; the executor asks the configured factory for its BlockBasedTableOptions."""
    _need_bins()
    ops, opts = S.ALL["basic_bottommost"]()
    want = H.run_reference(ops, **opts)
    got = H.run_reference(ops, binary=H.REF_B200_BIN, table_factory="b200+nofallback", executor="b200", **opts)
    gm = got["manifest"]
    assert gm["executor"] == "B200Compact" and gm["remote_compact_read_bytes"] > 0 and gm["b200_device_tables"] >= len(got["inputs"])
    assert (gm["scan_count"], gm["scan_digest"]) == (want["manifest"]["scan_count"], want["manifest"]["scan_digest"])
    _same_tables(got["inputs"], want["inputs"])
    assert [m["num_entries"] for m in gm["outputs"]] == [m["num_entries"] for m in want["manifest"]["outputs"]]
