"""Full Bloom filter block on the device (GPU box): BlockBasedTableOptions::filter_policy = NewBloomFilterPolicy(bits), format_version 5
(FastLocalBloom over XXPH3 of the whole user keys; table/block_based/filter_policy.cc:60-127,304-506, util/bloom_impl.h:156-214,
BlockBasedTableBuilder::WriteFilterBlock block_based_table_builder.cc:1488-1538).  The product computes the hashes from the key columns,
counts the distinct consecutive hashes per output file, sizes the block, ORs the probe bits into the file image and shifts the index
block behind it (toplingdb_b200/csrc/bloom_rules.h, encode.cu bloom_*).  Checked against
 (a) output files of the unmodified reference, byte for byte,
 (b) the CPU oracle (pinned to the reference for this in tests/test_oracle_bloom.py) on seeded jobs: many files, every key length
     0..16, several versions per key, odd bits-per-key values, together with grandparent cuts / key ranges / paranoid re-read,
 (c) the reference DB writing its files through the B200 executor plugin and then answering Get() through those filters."""
import os
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

pytestmark = pytest.mark.gpu


def _first_diff(a, b):
    return next((j for j in range(min(len(a), len(b))) if a[j] != b[j]), min(len(a), len(b)))


@pytest.mark.parametrize("case,kw,bits", [("basic_bottommost", {}, 10), ("snapshots_nonbottom", {}, 10), ("varlen_keys", {}, 10),
                                          ("cfg3_mini", {}, 10), ("cfg2_mini", {}, 6.5), ("tiny", {}, 10), ("all_deleted", {}, 10),
                                          ("same_user_key_across_blocks", {}, 16), ("nonbottom_tombstones", dict(n=6000), 3),
                                          ("output_level0", {}, 24), ("crc32c_small_blocks", {}, 10), ("ttl_filter", {}, 10)])
def test_device_writes_the_reference_filter_block(case, kw, bits):
    from gpu_harness import run_product
    if not H.have_ref():
        pytest.fail("oracle/_ref missing: run __graft_entry__.build() where /root/reference exists")
    ops, opts = S.ALL[case](**kw)
    ref = H.run_reference(ops, bloom_bits=bits, **opts)
    p = H.params_from_reference(ref)
    assert p.bloom_millibits_per_key == int(bits * 1000)
    files, metas, st = run_product(p, ref["inputs"])
    assert [len(f) for f in files] == [len(f) for f in ref["outputs"]]
    for i, (a, b) in enumerate(zip(files, ref["outputs"])):
        assert a == b, f"{case}: output {i} differs at byte {_first_diff(a, b)} of {len(a)}"
    for k in H.STAT_KEYS:
        assert getattr(st, k) == ref["manifest"]["stats"][k], k


def _seeded_inputs(seed, nruns, n, vlen, varlen):
    rnd = random.Random(seed)
    runs, seq = [], 1
    universe = n * nruns // 2
    for _ in range(nruns):
        run = {}
        for kk in rnd.sample(range(universe), n):
            kb = struct.pack(">QQ", kk * 0x9E3779B97F4A7C15 & ((1 << 64) - 1), kk)
            if varlen:
                kb = kb[:kk % 17]
            t = 0 if rnd.random() < 0.1 else 1
            run[kb] = (kb + struct.pack("<Q", (seq << 8) | t), b"" if t == 0 else rnd.randbytes(vlen))
            seq += 1
        runs.append([run[k] for k in sorted(run)])
    return [H.oracle_build_sst(H.Params(), H.kvstream(r)) for r in reversed(runs)], seq


@pytest.mark.parametrize("seed,nruns,n,vlen,varlen,millibits,target,snaps", [
    (1, 6, 8000, 24, False, 10000, 128 << 10, False), (2, 4, 6000, 100, True, 10000, 256 << 10, True),
    (3, 8, 4000, 8, False, 1000, 64 << 10, False), (4, 3, 20000, 40, False, 13500, 4 << 20, True),
    (5, 5, 5000, 300, True, 50001, 512 << 10, False), (6, 2, 30000, 0, False, 7777, 96 << 10, False)])
def test_seeded_jobs_with_filters_match_the_oracle(seed, nruns, n, vlen, varlen, millibits, target, snaps):
    from gpu_harness import run_product
    inputs, seq = _seeded_inputs(seed, nruns, n, vlen, varlen)
    rnd = random.Random(seed)
    p = H.Params(bottommost_level=bool(seed % 2), max_output_file_size=target, file_creation_times=[3, 4, 5],
                 snapshots=sorted(rnd.sample(range(1, seq), 4)) if snaps else [], bloom_millibits_per_key=millibits)
    want, _, wst = H.oracle_compact(p, inputs)
    files, _, st = run_product(p, inputs, device_inputs=bool(seed % 2))
    assert [len(f) for f in files] == [len(f) for f in want]
    for i, (a, b) in enumerate(zip(files, want)):
        assert a == b, f"output {i} differs at byte {_first_diff(a, b)} of {len(a)}"
    assert st.num_output_records == wst.num_output_records


def test_filters_with_grandparent_cuts_ranges_and_paranoid_reread():
    from gpu_harness import run_product
    p, inputs = gp_cases.build(**gp_cases.CASES["many_large_grandparents"])
    p.bloom_millibits_per_key = 10000
    want, _, _ = H.oracle_compact(p, inputs)
    assert run_product(p, inputs, paranoid_file_checks=1)[0] == want
    p.range_start = p.grandparents[len(p.grandparents) // 2][0]
    want, _, _ = H.oracle_compact(p, inputs)
    assert run_product(p, inputs, paranoid_file_checks=1)[0] == want


def test_filter_needs_format_version_5():
    from gpu_harness import job_from_params
    import toplingdb_b200 as T
    with pytest.raises(T.B200cError) as ei:
        job_from_params(H.Params(format_version=4, bloom_millibits_per_key=10000))
    assert ei.value.code == T.native.ERR_NOT_SUPPORTED


@pytest.mark.parametrize("case,bits", [("cfg3_mini", 10), ("snapshots_nonbottom", 7)])
def test_reference_db_reads_through_filters_the_executor_wrote(case, bits):
    if not (os.path.exists(H.REF_BIN) and os.path.exists(H.REF_B200_BIN)):
        pytest.fail("oracle/_ref/ref_compact(_b200) missing")
    ops, opts = S.ALL[case]()
    want = H.run_reference(ops, bloom_bits=bits, **opts)
    got = H.run_reference(ops, binary=H.REF_B200_BIN, executor="b200", bloom_bits=bits, **opts)
    gm, wm = got["manifest"], want["manifest"]
    assert gm["executor"] == "B200Compact" and gm["remote_compact_read_bytes"] > 0
    assert (gm["scan_count"], gm["scan_digest"]) == (wm["scan_count"], wm["scan_digest"])
    assert gm["get_missing"] == 0 and gm["get_found"] == wm["get_found"] == wm["scan_count"]
    assert len(got["outputs"]) == len(want["outputs"])
    for g, w in zip(got["outputs"], want["outputs"]):
        tg, tw = sstfmt.parse_sst(g), sstfmt.parse_sst(w)
        (go, gs), (wo, ws) = tg["metaindex"]["fullfilter.rocksdb.BuiltinBloomFilter"], tw["metaindex"]["fullfilter.rocksdb.BuiltinBloomFilter"]
        assert (go, gs) == (wo, ws) and g[go:go + gs + 5] == w[wo:wo + ws + 5]  # filter block + trailer
        assert [g[h[0]:h[0] + h[1] + 5] for _, h in tg["index"]] == [w[h[0]:h[0] + h[1] + 5] for _, h in tw["index"]]
        for k in ("rocksdb.filter.size", "rocksdb.num.filter_entries", "rocksdb.filter.policy", "rocksdb.index.size", "rocksdb.data.size"):
            assert tg["properties"][k] == tw["properties"][k], k
