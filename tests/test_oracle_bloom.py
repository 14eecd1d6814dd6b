"""Full Bloom filter block (BlockBasedTableOptions::filter_policy = NewBloomFilterPolicy(bits), format_version 5: FastLocalBloom over
XXPH3 hashes of the whole user keys; table/block_based/filter_policy.cc:60-127,304-506, util/bloom_impl.h:156-214,
table/block_based/full_filter_block.cc, BlockBasedTableBuilder::WriteFilterBlock block_based_table_builder.cc:1488-1538): the oracle's
restatement pinned against output files of the compiled reference, byte for byte (filter block, metaindex, properties, footer)."""
import pytest

import helpers as H
import scenarios as S
import sstfmt

needs_ref = pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("case,kw,bits", [("basic_bottommost", {}, 10), ("snapshots_nonbottom", {}, 10), ("varlen_keys", {}, 10),
                                          ("long_keys", {}, 10), ("cfg3_mini", {}, 10), ("cfg2_mini", {}, 6.5), ("tiny", {}, 10),
                                          ("all_deleted", {}, 10), ("same_user_key_across_blocks", {}, 16),
                                          ("nonbottom_tombstones", dict(n=6000), 3), ("output_level0", {}, 24)])
def test_oracle_writes_the_reference_filter_block(case, kw, bits):
    ops, opts = S.ALL[case](**kw)
    ref = H.run_reference(ops, bloom_bits=bits, **opts)
    p = H.params_from_reference(ref)
    assert p.bloom_millibits_per_key == int(bits * 1000)
    files, metas, st = H.oracle_compact(p, ref["inputs"])
    assert [len(f) for f in files] == [len(f) for f in ref["outputs"]]
    assert files == ref["outputs"]
    for f in files:
        t = sstfmt.parse_sst(f)
        off, size = t["metaindex"]["fullfilter.rocksdb.BuiltinBloomFilter"]
        assert size == sstfmt.prop_u64(t["properties"], "rocksdb.filter.size") and size % 64 == 5
        assert t["properties"]["rocksdb.filter.policy"] == b"bloomfilter"
        ukeys = {ik[:-8] for ik, _ in t["entries"]}
        assert sstfmt.prop_u64(t["properties"], "rocksdb.num.filter_entries") == len(ukeys)
    for k in H.STAT_KEYS:
        assert getattr(st, k) == ref["manifest"]["stats"][k], k


def test_xxph3_known_answers_of_the_reference():
    """util/hash_test.cc TEST(HashTest, Hash64SmallValueSchema) (seed 0 = GetSliceHash64), extracted by tests/golden/make_hash64_kat.py"""
    import ctypes as C
    import json
    import os
    L = H.oracle()
    L.orc_xxph3_64.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint64)]
    kat = json.load(open(os.path.join(H.GOLDEN_DIR, "hash64_kat.json")))["vectors"]
    assert len(kat) >= 40
    out = C.c_uint64()
    for v in kat:
        d = bytes.fromhex(v["hex"])
        assert L.orc_xxph3_64(d, len(d), C.byref(out)) == 0 and out.value == v["hash64"], v
    assert L.orc_xxph3_64(bytes(129), 129, C.byref(out)) != 0  # the striped long-input loop is not restated
