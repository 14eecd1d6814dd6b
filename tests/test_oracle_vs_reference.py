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
