"""CPU sanity for tests/gp_cases.py: the synthetic grandparent jobs really exercise the grandparent cut rules of the oracle
(its restatement of compaction_outputs.cc:231-354 is pinned to the compiled reference in test_oracle_grandparents.py), so the
GPU parity test over the same cases (test_gpu_grandparents.py) compares something meaningful."""
import pytest

import gp_cases
import helpers as H


@pytest.mark.parametrize("name", sorted(gp_cases.CASES))
def test_case_exercises_the_grandparent_rules(name):
    p, inputs = gp_cases.build(**gp_cases.CASES[name])
    assert p.grandparents == sorted(p.grandparents) and all(a <= b for a, b, _ in p.grandparents)
    assert all(p.grandparents[i][1] <= p.grandparents[i + 1][0] for i in range(len(p.grandparents) - 1))
    with_gp, _, st = H.oracle_compact(p, inputs)
    p.grandparents = []
    plain, _, st0 = H.oracle_compact(p, inputs)
    assert st.num_output_records == st0.num_output_records
    if name == "grandparents_behind_the_stream":
        assert with_gp == plain
    else:
        assert len(with_gp) > len(plain), (len(with_gp), len(plain))
