"""Sub-compaction key ranges (SubcompactionState::start/end; ProcessKeyValueCompaction wraps the merged input in a ClippingIterator,
db/compaction/compaction_job.cc:1433-1519, db/compaction/clipping_iterator.h): the oracle's range clipping pinned against the
unmodified reference running the same job with max_subcompactions > 1.  The reference reports per sub-compaction only statistics;
helpers.subcompaction_ranges derives the exact ranges from them.  Concurrent sub-compactions draw file numbers from one counter, so
the property `rocksdb.original.file.number` (and with it the properties block bytes) is the only thing not compared."""
import pytest

import helpers as H
import scenarios as S
import sstfmt

needs_ref = pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
SUB_STATS = ("num_input_deletion_records", "num_expired_deletion_records", "num_records_replaced", "total_input_raw_key_bytes",
             "total_input_raw_value_bytes")


def file_parts(data):
    """(data blocks with trailers, index block with trailer, properties minus the file number)"""
    t = sstfmt.parse_sst(data)
    blocks = [data[h[0]:h[0] + h[1] + 5] for _, h in t["index"]]
    io, isz = t["footer"]["index"]
    return blocks, data[io:io + isz + 5], {k: v for k, v in t["properties"].items() if k != "rocksdb.original.file.number"}


CASES = [("cfg2_mini", {}, 4), ("cfg3_mini", {}, 4), ("snapshots", dict(n=6000), 3), ("nonbottom_tombstones", dict(n=8000), 4),
         ("snapshots_nonbottom", dict(n=6000), 4), ("varlen_keys", dict(n=8000), 3), ("filter_empty_value", dict(n=8000), 2),
         ("ttl_filter", dict(n=8000), 3)]


@needs_ref
@pytest.mark.parametrize("case,kw,subs", CASES)
def test_clipped_oracle_jobs_reproduce_the_reference_subcompactions(case, kw, subs):
    ops, opts = S.ALL[case](**kw)
    ref = H.run_reference(ops, max_subcompactions=subs, **opts)
    man = ref["manifest"]
    ranges = H.subcompaction_ranges(ref)
    assert len(ranges) >= 2, "the reference did not split this job"
    props = [sstfmt.parse_sst(o)["properties"] for o in ref["outputs"]]
    k = 0
    total_in = 0
    for start, end, rstats in ranges:
        p = H.params_from_reference(ref)
        p.range_start, p.range_end = start, end
        # file numbers / creation times of this range's files as the reference assigned them
        files, metas, st = H.oracle_compact(p, ref["inputs"])
        want = ref["outputs"][k:k + len(files)]
        assert len(want) == len(files)
        p.file_creation_times = [sstfmt.prop_u64(q, "rocksdb.file.creation.time") for q in props[k:k + len(files)]] or [0]
        files, metas, st = H.oracle_compact(p, ref["inputs"])
        for got, exp in zip(files, want):
            assert len(got) == len(exp)
            assert file_parts(got) == file_parts(exp)
        for key in SUB_STATS:
            assert getattr(st, key) == rstats[key], key
        total_in += st.num_input_records
        k += len(files)
    assert k == len(ref["outputs"])
    assert total_in == man["stats"]["num_input_records"]


def test_empty_and_unbounded_ranges():
    g = H.load_golden("basic_bottommost")
    p = H.params_from_reference(g)
    whole, _, st = H.oracle_compact(p, g["inputs"])
    keys = [ik[:-8] for f in whole for ik, _ in sstfmt.parse_sst(f)["entries"]]
    mid = keys[len(keys) // 2]
    p.range_start, p.range_end = mid, mid  # empty range: no output file at all
    files, _, st0 = H.oracle_compact(p, g["inputs"])
    assert files == [] and st0.num_input_records == 0
    p.range_start, p.range_end = None, mid
    a, _, sa = H.oracle_compact(p, g["inputs"])
    p.range_start, p.range_end = mid, None
    b, _, sb = H.oracle_compact(p, g["inputs"])
    got = [ik for f in a + b for ik, _ in sstfmt.parse_sst(f)["entries"]]
    assert got == [ik for f in whole for ik, _ in sstfmt.parse_sst(f)["entries"]]
    assert sa.num_input_records + sb.num_input_records == st.num_input_records
    assert sa.num_output_records + sb.num_output_records == st.num_output_records


@needs_ref
@pytest.mark.parametrize("seed,n", [(5, 40000), (6, 30000)])
def test_subcompactions_of_a_picker_built_job_with_grandparents_and_filters(seed, n):
    """everything at once, as a live DB does it: DB::CompactRange picks L0 -> L1 with the L2 files as grandparents, the job is split
    into 4 sub-compactions, the table has a Bloom filter policy.  Every sub-compaction starts its own grandparent walk at its first
    key (CompactionOutputs is per sub-compaction), which is what a clipped oracle job does."""
    ops, opts = S.grandparent_cuts(n=n, seed=seed)
    ref = H.run_reference(ops, max_subcompactions=4, bloom_bits=10, **opts)
    assert len(ref["manifest"]["grandparents"]) >= 2
    ranges = H.subcompaction_ranges(ref)
    assert len(ranges) >= 2
    props = [sstfmt.parse_sst(o)["properties"] for o in ref["outputs"]]
    k = 0
    for start, end, rstats in ranges:
        p = H.params_from_reference(ref)
        p.range_start, p.range_end = start, end
        nfiles = len(H.oracle_compact(p, ref["inputs"])[0])
        p.file_creation_times = [sstfmt.prop_u64(q, "rocksdb.file.creation.time") for q in props[k:k + nfiles]] or [0]
        files, _, st = H.oracle_compact(p, ref["inputs"])
        want = ref["outputs"][k:k + nfiles]
        assert [len(f) for f in files] == [len(f) for f in want]
        for got, exp in zip(files, want):
            assert file_parts(got) == file_parts(exp)
            t = sstfmt.parse_sst(got)
            fo, fs = t["metaindex"]["fullfilter.rocksdb.BuiltinBloomFilter"]
            assert got[fo:fo + fs + 5] == exp[fo:fo + fs + 5]
        for key in SUB_STATS:
            assert getattr(st, key) == rstats[key], key
        k += nfiles
    assert k == len(ref["outputs"])
