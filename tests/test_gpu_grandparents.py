"""Grandparent-aware output cutting on the device (GPU box).  CompactionOutputs::ShouldStopBefore
(db/compaction/compaction_outputs.cc:231-354) cuts an output file in front of a key that crosses grandparent-file
boundaries when the overlap, skippable-file or pre-cut rule says so; the product evaluates the same rules on entry ranks
inside the encoder's serial stitch walk (toplingdb_b200/csrc/encode.cu, chase_tile / gp_*).  Checked here against
 (a) the unmodified reference itself, on jobs its own picker builds (DB::CompactRange attaches the grandparents),
 (b) the CPU oracle (pinned to the reference for these rules in test_oracle_grandparents.py) on synthetic grandparent shapes
     that take every branch (tests/gp_cases.py),
 (c) the reference DB running that picker-built job through the B200 executor plugin."""
import os

import pytest

try:
    import torch  # noqa: F401
except Exception:  # pragma: no cover
    torch = None

import gp_cases
import helpers as H
import scenarios as S

pytestmark = pytest.mark.gpu


def _first_diff(a, b):
    return next((j for j in range(min(len(a), len(b))) if a[j] != b[j]), min(len(a), len(b)))


@pytest.mark.parametrize("seed,n", [(18, 8000), (5, 40000), (6, 30000), (7, 30000)])
def test_device_cuts_files_where_the_reference_does(seed, n):
    from gpu_harness import run_product
    if not H.have_ref():
        pytest.fail("oracle/_ref missing: run __graft_entry__.build() where /root/reference exists")
    ops, opts = S.grandparent_cuts(n=n, seed=seed)
    ref = H.run_reference(ops, **opts)
    man = ref["manifest"]
    assert man["mode"] == "range" and len(man["grandparents"]) >= 2
    p = H.params_from_reference(ref)
    files, metas, st = run_product(p, ref["inputs"])
    assert [len(f) for f in files] == [len(f) for f in ref["outputs"]]
    for i, (a, b) in enumerate(zip(files, ref["outputs"])):
        assert a == b, f"output {i} differs at byte {_first_diff(a, b)}"
    for k in H.STAT_KEYS:
        assert getattr(st, k) == man["stats"][k], k
    for m, want in zip(metas, man["outputs"]):
        assert (m.file_size, m.num_entries, m.num_deletions) == (want["size"], want["num_entries"], want["num_deletions"])
        assert bytes(m.smallest_ikey[: m.smallest_ikey_len - 8]).hex() == want["smallestkey"]
        assert bytes(m.largest_ikey[: m.largest_ikey_len - 8]).hex() == want["largestkey"]


@pytest.mark.parametrize("name", sorted(gp_cases.CASES))
@pytest.mark.parametrize("device_inputs", [False, True])
def test_synthetic_grandparent_shapes_match_the_oracle(name, device_inputs):
    from gpu_harness import run_product
    if device_inputs and name not in ("dynamic_mixed", "many_large_grandparents"):
        pytest.skip("device-resident inputs: two shapes are enough")
    p, inputs = gp_cases.build(**gp_cases.CASES[name])
    want, wmetas, wst = H.oracle_compact(p, inputs)
    files, metas, st = run_product(p, inputs, device_inputs=device_inputs)
    assert [len(f) for f in files] == [len(f) for f in want]
    for i, (a, b) in enumerate(zip(files, want)):
        assert a == b, f"{name}: output {i} differs at byte {_first_diff(a, b)}"
    for k in ("num_output_records", "num_records_replaced", "num_input_deletion_records", "total_input_raw_key_bytes"):
        assert getattr(st, k) == getattr(wst, k), k


def test_job_is_reusable_after_a_grandparent_run():
    """the cut list and the boundary state live in per-job device buffers: a second job on the same device starts clean"""
    from gpu_harness import run_product
    p, inputs = gp_cases.build(**gp_cases.CASES["many_large_grandparents"])
    a, _, _ = run_product(p, inputs)
    p.grandparents = []
    b, _, _ = run_product(p, inputs)
    want, _, _ = H.oracle_compact(p, inputs)
    assert b == want and len(a) > len(b)


def test_grandparent_keys_longer_than_the_device_columns_are_rejected():
    from gpu_harness import job_from_params
    import toplingdb_b200 as T
    p, inputs = gp_cases.build(**gp_cases.CASES["dynamic_mixed"])
    a, b, sz = p.grandparents[-1]
    p.grandparents[-1] = (a, b + b"x" * 8, sz)
    with pytest.raises(T.B200cError) as ei:
        job_from_params(p)
    assert ei.value.code == T.native.ERR_NOT_SUPPORTED


@pytest.mark.parametrize("seed,n", [(5, 40000), (18, 8000)])
def test_reference_db_runs_the_picker_built_job_through_the_b200_executor(seed, n):
    if not (os.path.exists(H.REF_BIN) and os.path.exists(H.REF_B200_BIN)):
        pytest.fail("oracle/_ref/ref_compact(_b200) missing")
    ops, opts = S.grandparent_cuts(n=n, seed=seed)
    want = H.run_reference(ops, **opts)
    got = H.run_reference(ops, binary=H.REF_B200_BIN, executor="b200", **opts)
    gm, wm = got["manifest"], want["manifest"]
    assert gm["executor"] == "B200Compact" and gm["remote_compact_read_bytes"] > 0
    assert len(gm["grandparents"]) == len(wm["grandparents"]) >= 2
    assert (gm["scan_count"], gm["scan_digest"]) == (wm["scan_count"], wm["scan_digest"])
    for k in ("num_entries", "num_deletions", "smallestkey", "largestkey"):
        assert [m[k] for m in gm["outputs"]] == [m[k] for m in wm["outputs"]], k
    # file sizes: equal up to the width of the one property that differs by design -- rocksdb.original.file.number is the executor's own
    # job-unique number on the RunRemote branch (the DB renames the file to a number it allocates afterwards, compaction_job.cc:1022-1034)
    import sstfmt

    def size_without_number(m, data):
        return m["size"] - len(sstfmt.parse_sst(data)["properties"]["rocksdb.original.file.number"])

    assert [size_without_number(m, d) for m, d in zip(gm["outputs"], got["outputs"])] == \
           [size_without_number(m, d) for m, d in zip(wm["outputs"], want["outputs"])]
