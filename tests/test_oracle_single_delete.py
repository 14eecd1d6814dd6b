"""SingleDelete in the oracle's CompactionIterator restatement (db/compaction/compaction_iterator.cc:635-661,662-887), pinned against
the compiled reference on seeded scenarios: Put / SingleDelete pairs across runs and snapshot stripes, repeated pairs, dangling
SingleDeletes, bottommost and not.  The reference's own known answers for these rules are in tests/test_compaction_job_kat.py.
Groundwork only: the device path still rejects kTypeSingleDeletion inputs with B200C_ERR_NOT_SUPPORTED."""
import pytest

import helpers as H
import scenarios as S
import sstfmt

pytestmark = pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("name", ["single_deletes", "single_deletes_nonbottom"])
@pytest.mark.parametrize("seed", [19, 20, 21, 22])
def test_oracle_matches_live_reference_with_single_deletes(name, seed):
    ops, opts = S.ORACLE_ONLY[name](seed=seed)
    ref = H.run_reference(ops, **opts)
    n_sd = sum(1 for d in ref["inputs"] for ik, _ in sstfmt.parse_sst(d)["entries"] if ik[-8] == 7)
    assert n_sd > 100
    p = H.params_from_reference(ref)
    files, _, st = H.oracle_compact(p, ref["inputs"])
    assert files == ref["outputs"]
    for k in H.STAT_KEYS:
        assert getattr(st, k) == ref["manifest"]["stats"][k], k


def test_single_delete_and_delete_on_one_key_fails_the_job():
    """enforce_single_del_contracts (default true): Status::Corruption (compaction_iterator.cc:779-800)"""
    newer = H.oracle_build_sst(H.Params(), H.kvstream([(H.ikey(b"k", 9, 7), b"")]))
    older = H.oracle_build_sst(H.Params(), H.kvstream([(H.ikey(b"k", 5, 0), b"")]))
    with pytest.raises(Exception):
        H.oracle_compact(H.Params(bottommost_level=False), [newer, older])
