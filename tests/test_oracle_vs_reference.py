"""Live cross-check of the CPU oracle against the compiled, unmodified reference (oracle/_ref) on freshly
generated, larger and re-seeded scenarios.  Skipped where oracle/_ref is not built."""
import pytest

import helpers as H
import scenarios as S

pytestmark = pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built (needs /root/reference)")

CASES = [(n, seed) for n in S.ALL for seed in (101, 202)]


@pytest.mark.ref
@pytest.mark.parametrize("name,seed", CASES)
def test_oracle_matches_live_reference(name, seed):
    ops, opts = S.ALL[name](seed=seed)
    ref = H.run_reference(ops, **opts)
    p = H.params_from_reference(ref)
    files, metas, st = H.oracle_compact(p, ref["inputs"])
    assert [len(f) for f in files] == [len(o) for o in ref["outputs"]]
    assert files == ref["outputs"]
    for k in H.STAT_KEYS:
        assert getattr(st, k) == ref["manifest"]["stats"][k], k


@pytest.mark.ref
def test_reference_multi_file_cut_and_meta():
    ops, opts = S.basic_bottommost(n=3000, nruns=4, seed=77)
    opts["target_file_size"] = 100 << 10
    ref = H.run_reference(ops, **opts)
    assert len(ref["outputs"]) > 5
    p = H.params_from_reference(ref)
    files, metas, _ = H.oracle_compact(p, ref["inputs"])
    assert files == ref["outputs"]
    for m, want in zip(metas, ref["manifest"]["outputs"]):
        assert m.file_number == want["file_number"] and m.file_size == want["size"]


def _matrix():
    import random
    rnd = random.Random(20260923)
    names = ["basic_bottommost", "nonbottom_tombstones", "snapshots", "snapshots_nonbottom", "varlen_keys", "cfg2_mini", "cfg3_mini",
             "filter_empty_value", "ttl_filter_nonbottom", "same_user_key_across_blocks"]
    out = []
    for i in range(20):
        fv = rnd.choice([3, 4, 5, 5])
        out.append((rnd.choice(names), 300 + i, dict(block_size=rnd.choice([512, 1024, 4096, 16384]),
                                                       restart_interval=rnd.choice([1, 4, 16, 32]), format_version=fv,
                                                       checksum=rnd.choice(["xxh3", "crc32c"]),
                                                       bloom_bits=rnd.choice([0, 0, 10, 14.5]) if fv == 5 else 0,
                                                       max_subcompactions=1)))
    return out


@pytest.mark.ref
@pytest.mark.parametrize("name,seed,table", _matrix())
def test_oracle_matches_reference_over_a_table_option_matrix(name, seed, table):
    """block size / restart interval / format version / checksum / filter policy drawn at random (fixed seed) for re-seeded scenarios:
    the byte layout of every block kind moves with them (restart arrays, delta-encoded index values from format_version 4, filter block
    and metaindex entry), the oracle has to follow the reference through all of it"""
    ops, opts = S.ALL[name](seed=seed)
    opts = dict(opts, **table)
    ref = H.run_reference(ops, **opts)
    p = H.params_from_reference(ref)
    assert (p.block_size, p.block_restart_interval, p.format_version) == (table["block_size"], table["restart_interval"], table["format_version"])
    files, metas, st = H.oracle_compact(p, ref["inputs"])
    assert [len(f) for f in files] == [len(o) for o in ref["outputs"]]
    assert files == ref["outputs"]
    for k in H.STAT_KEYS:
        assert getattr(st, k) == ref["manifest"]["stats"][k], k
