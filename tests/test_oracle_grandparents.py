"""Grandparent-aware output cutting (CompactionOutputs::ShouldStopBefore, db/compaction/compaction_outputs.cc:231-354:
max_compaction_bytes overlap, skippable-file and pre-cut rules) restated in the CPU oracle and pinned against the compiled
reference.  The job is built by the DB's own picker (DB::CompactRange) because DB::CompactFiles never attaches grandparents.
GPU counterpart: tests/test_gpu_grandparents.py; host-compiled rank rules: tests/test_gp_rules_host.py."""
import pytest

import helpers as H
import scenarios as S

needs_ref = pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("seed,n", [(18, 8000), (5, 40000), (6, 30000)])
def test_oracle_cuts_files_where_the_reference_does(seed, n):
    ops, opts = S.grandparent_cuts(n=n, seed=seed)
    ref = H.run_reference(ops, **opts)
    man = ref["manifest"]
    assert man["mode"] == "range" and len(man["grandparents"]) >= 2 and not man["bottommost_level"]
    assert man["max_output_file_size"] == 2 * man["target_output_file_size"]  # compaction.cc:289-295
    p = H.params_from_reference(ref)
    files, metas, st = H.oracle_compact(p, ref["inputs"])
    assert [len(f) for f in files] == [len(f) for f in ref["outputs"]]
    assert files == ref["outputs"]
    for k in H.STAT_KEYS:
        assert getattr(st, k) == man["stats"][k], k
    # the grandparent rules really decided the cuts: without them the size rule alone gives fewer files
    p.grandparents = []
    plain, _, _ = H.oracle_compact(p, ref["inputs"])
    if n >= 30000:
        assert len(plain) < len(files)
