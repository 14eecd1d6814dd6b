"""Sub-compaction key ranges on the device (GPU box): `range_start` / `range_end` of the C ABI clip the merged input like the
ClippingIterator that ProcessKeyValueCompaction puts in front of a sub-compaction's CompactionIterator
(db/compaction/compaction_job.cc:1433-1519, db/compaction/clipping_iterator.h).  Checked against
 (a) the unmodified reference running the same job with max_subcompactions > 1 (ranges derived from its per-sub-compaction
     statistics, tests/helpers.py subcompaction_ranges; the oracle half of this is tests/test_oracle_subcompactions.py),
 (b) the CPU oracle on seeded jobs with ranges that hit the edge cases (empty, one-sided, bounds that are / are not keys of the
     input, bounds shorter than the keys, several versions of the boundary key, together with grandparents / a filter),
 (c) the partition property: the outputs of the ranges of one job, concatenated, hold exactly the entries of the unclipped job."""
import random
import struct

import pytest

try:
    import torch  # noqa: F401
except Exception:  # pragma: no cover
    torch = None

import gp_cases
import helpers as H
import scenarios as S
import sstfmt
from test_oracle_subcompactions import CASES, SUB_STATS, file_parts

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case,kw,subs", CASES)
def test_device_ranges_reproduce_the_reference_subcompactions(case, kw, subs):
    from gpu_harness import run_product
    if not H.have_ref():
        pytest.fail("oracle/_ref missing: run __graft_entry__.build() where /root/reference exists")
    ops, opts = S.ALL[case](**kw)
    ref = H.run_reference(ops, max_subcompactions=subs, **opts)
    ranges = H.subcompaction_ranges(ref)
    assert len(ranges) >= 2
    props = [sstfmt.parse_sst(o)["properties"] for o in ref["outputs"]]
    k = 0
    for start, end, rstats in ranges:
        p = H.params_from_reference(ref)
        p.range_start, p.range_end = start, end
        nfiles = len(H.oracle_compact(p, ref["inputs"])[0])
        want = ref["outputs"][k:k + nfiles]
        p.file_creation_times = [sstfmt.prop_u64(q, "rocksdb.file.creation.time") for q in props[k:k + nfiles]] or [0]
        files, metas, st = run_product(p, ref["inputs"])
        assert [len(f) for f in files] == [len(f) for f in want]
        for got, exp in zip(files, want):
            assert file_parts(got) == file_parts(exp)
        for key in SUB_STATS:
            assert getattr(st, key) == rstats[key], key
        k += nfiles
    assert k == len(ref["outputs"])


def _seeded(seed, nruns=5, n=6000, vlen=24, snapshots=False):
    rnd = random.Random(seed)
    runs, seq = [], 1
    universe = n * nruns // 2
    for _ in range(nruns):
        run = []
        for kk in sorted(rnd.sample(range(universe), n)):
            t = 0 if rnd.random() < 0.1 else 1
            run.append((struct.pack(">QQ", 7, kk * 3) + struct.pack("<Q", (seq << 8) | t), b"" if t == 0 else rnd.randbytes(vlen)))
            seq += 1
        runs.append(run)
    inputs = [H.oracle_build_sst(H.Params(), H.kvstream(r)) for r in reversed(runs)]
    snaps = sorted(rnd.sample(range(1, seq), 5)) if snapshots else []
    return inputs, snaps, universe


def _k(v):
    return struct.pack(">QQ", 7, v)


@pytest.mark.parametrize("bottom", [False, True])
@pytest.mark.parametrize("snapshots", [False, True])
def test_seeded_ranges_match_the_oracle(bottom, snapshots):
    from gpu_harness import run_product
    inputs, snaps, universe = _seeded(41 + int(bottom) + 2 * int(snapshots), snapshots=snapshots)
    top = 3 * universe
    ranges = [(None, _k(top // 3)), (_k(top // 3), _k(2 * top // 3)), (_k(2 * top // 3), None),  # a partition (bounds are keys)
              (_k(top // 2 + 1), _k(top // 2 + 2)),        # bounds between keys, nothing inside
              (_k(top // 2), _k(top // 2)),                # empty
              (_k(top // 2), _k(top // 4)),                # end in front of start: empty
              (_k(top + 10), None), (None, _k(0)),         # behind / in front of every key
              (None, None),
              (_k(0x3000)[:15], _k(0x9000)[:15]),             # bounds shorter than the keys (proper prefixes sort first)
              (b"", _k(3000)), (_k(3000), b"\xff" * 16)]
    for start, end in ranges:
        p = H.Params(bottommost_level=bottom, max_output_file_size=96 << 10, snapshots=snaps, file_creation_times=[7, 8],
                     range_start=start, range_end=end)
        want, _, wst = H.oracle_compact(p, inputs)
        files, _, st = run_product(p, inputs)
        assert [len(f) for f in files] == [len(f) for f in want], (start, end)
        assert files == want, (start, end)
        for key in ("num_input_records", "num_output_records", "num_records_replaced", "num_input_deletion_records",
                    "num_expired_deletion_records", "total_input_raw_key_bytes", "total_input_raw_value_bytes"):
            assert getattr(st, key) == getattr(wst, key), (key, start, end)


def test_ranges_with_grandparents_and_a_filter():
    """a sub-compaction starts its grandparent walk at its own first key (CompactionOutputs is per sub-compaction)"""
    from gpu_harness import run_product
    p, inputs = gp_cases.build(**gp_cases.CASES["many_small_grandparents"])
    keys = sorted({ik[:-8] for d in inputs for ik, _ in sstfmt.parse_sst(d)["entries"]})
    cuts = [keys[len(keys) // 3], keys[2 * len(keys) // 3]]
    for start, end in zip([None] + cuts, cuts + [None]):
        p.range_start, p.range_end = start, end
        p.compaction_filter = "remove_empty_value"
        want, _, wst = H.oracle_compact(p, inputs)
        files, _, st = run_product(p, inputs)
        assert files == want
        assert st.num_record_drop_user == wst.num_record_drop_user and st.num_input_records == wst.num_input_records


def test_ranges_partition_a_full_size_job():
    """cfg2-shaped job (8 runs) split into 4 ranges at quartiles of the key space: every range on its own, then the union"""
    from gpu_harness import run_product
    import toplingdb_b200 as T
    from toplingdb_b200 import sharding
    inputs, _, universe = _seeded(77, nruns=8, n=40000, vlen=32)
    top = 3 * universe
    p = H.Params(output_level=1, bottommost_level=False, max_output_file_size=2 << 20, file_creation_times=[5])
    whole, _, wst = run_product(p, inputs)
    got_entries, n_in, n_out = [], 0, 0
    for start, end in sharding.subcompaction_ranges([_k(top // 4), _k(top // 2), _k(3 * top // 4)]):
        p.range_start, p.range_end = start, end
        files, metas, st = run_product(p, inputs, device_inputs=True)
        for f in files:
            got_entries += sstfmt.parse_sst(f)["entries"]
        n_in += st.num_input_records
        n_out += st.num_output_records
    want_entries = [e for f in whole for e in sstfmt.parse_sst(f)["entries"]]
    assert got_entries == want_entries
    assert (n_in, n_out) == (wst.num_input_records, wst.num_output_records)


def test_range_bound_longer_than_the_device_columns_is_rejected():
    from gpu_harness import job_from_params
    import toplingdb_b200 as T
    p = H.Params(range_start=b"k" * 17)
    with pytest.raises(T.B200cError) as ei:
        job_from_params(p)
    assert ei.value.code == T.native.ERR_NOT_SUPPORTED
