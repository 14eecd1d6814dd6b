"""What the executor plugin hands to the C ABI.  oracle/_ref/ref_compact_mock is the reference driver + the real plugin linked against a
TEST DOUBLE of libb200c (tests/native/mock_b200c.c): it claims a device, records the b200c_params of b200c_job_create and the inputs
added, then refuses to run, and with `executor=b200+fallback` the reference compacts the job itself.  The recorded translation
(CompactionParams + BlockBasedTableOptions + Compaction -> b200c_params, toplingdb_b200/plugin/b200_compaction_executor.cc Execute) is
compared with what the reference reports about the very same job in the driver's manifest.  CPU only."""
import json
import os
import subprocess
import tempfile

import pytest

import helpers as H
import scenarios as S

MOCK_BIN = os.path.join(H.ROOT, "oracle", "_ref", "ref_compact_mock")
pytestmark = pytest.mark.skipif(not os.path.exists(MOCK_BIN), reason="oracle/_ref/ref_compact_mock not built (needs /root/reference)")
CKSUM = {"xxh3": 4, "crc32c": 1}


def run(fn, **extra):
    ops, opts = fn()
    opts = dict(opts, **extra)
    with tempfile.TemporaryDirectory(prefix="b200c_mock_") as d:
        with open(os.path.join(d, "ops.bin"), "wb") as f:
            f.write(ops.bytes())
        env = dict(os.environ, B200C_MOCK_DUMP=os.path.join(d, "dump.jsonl"))
        args = [MOCK_BIN, os.path.join(d, "ops.bin"), os.path.join(d, "w"), "executor=b200+fallback"] + [f"{k}={v}" for k, v in opts.items()]
        r = subprocess.run(args, capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        jobs = [json.loads(line) for line in open(os.path.join(d, "dump.jsonl"))]
        man = json.load(open(os.path.join(d, "w", "manifest.json")))
    # the measured job: same output level and the same input files as the manifest lists
    want_inputs = sorted(m["file_number"] for m in man["inputs"])
    mine = [j for j in jobs if j["output_level"] == man["output_level"] and sorted(i["file_number"] for i in j["inputs"]) == want_inputs]
    assert mine, (len(jobs), want_inputs)
    return mine[-1], man, opts


def check_common(j, man, opts):
    assert j["abi_version"] == 8 and j["device"] == 0 and j["output_mem"] == 0
    assert j["bottommost_level"] == int(man["bottommost_level"])
    assert j["max_output_file_size"] == man.get("max_output_file_size", man["target_file_size"])
    assert (j["block_size"], j["block_restart_interval"], j["format_version"]) == (man["block_size"], man["restart_interval"], man["format_version"])
    assert j["checksum"] == CKSUM[man["checksum"]] and j["index_block_restart_interval"] == 1 and j["block_size_deviation"] == 10
    assert j["snapshots"] == man["snapshots"]
    assert j["db_id"] == man["db_id"] and j["db_session_id"] == man["db_session_id"] and j["column_family_name"] == "default"
    assert j["bloom_millibits_per_key"] == man["bloom_millibits_per_key"]
    assert j["has_range_start"] == j["has_range_end"] == 0
    # L0 inputs newest first (MakeInputIterator order), whole files in host memory
    nums = [i["file_number"] for i in j["inputs"] if i["level"] == 0]
    assert nums == sorted(nums, reverse=True) and all(i["mem_kind"] == 0 for i in j["inputs"])
    assert sorted(i["len"] for i in j["inputs"]) == sorted(m["size"] for m in man["inputs"])


@pytest.mark.parametrize("name,extra", [("basic_bottommost", {}), ("snapshots_nonbottom", {}), ("crc32c_small_blocks", {}), ("output_level0", {}),
                                        ("cfg3_mini", dict(bloom_bits=10, paranoid=1)), ("varlen_keys", dict(format_version=4)),
                                        ("snapshots", dict(bloom_bits=6.5))])
def test_table_and_job_options_reach_the_c_abi(name, extra):
    j, man, opts = run(S.ALL[name], **extra)
    check_common(j, man, opts)
    assert j["compaction_filter"] == 0 and j["grandparents"] == []
    assert j["paranoid_file_checks"] == int(bool(extra.get("paranoid")))


@pytest.mark.parametrize("name,enum", [("filter_empty_value", 1), ("filter_empty_value_nonbottom", 1), ("ttl_filter", 2), ("ttl_filter_nonbottom", 2)])
def test_builtin_compaction_filters_are_recognised(name, enum):
    j, man, opts = run(S.ALL[name])
    check_common(j, man, opts)
    assert j["compaction_filter"] == enum
    if enum == 2:
        assert (j["ttl"], j["ttl_now"]) == (man["ttl"], man["now"])


@pytest.mark.parametrize("seed,n", [(18, 8000), (5, 40000)])
def test_grandparents_of_a_picker_built_job_reach_the_c_abi(seed, n):
    j, man, opts = run(lambda: S.grandparent_cuts(n=n, seed=seed))
    check_common(j, man, opts)
    assert [(g["smallestkey"], g["largestkey"], g["size"]) for g in j["grandparents"]] == \
           [(g["smallestkey"], g["largestkey"], g["size"]) for g in man["grandparents"]]
    assert len(j["grandparents"]) >= 2
    assert (j["max_compaction_bytes"], j["target_output_file_size"]) == (man["max_compaction_bytes"], man["target_output_file_size"])
    assert j["level_compaction_dynamic_file_size"] == int(man["level_compaction_dynamic_file_size"])


def test_a_refused_job_without_fallback_fails_with_the_library_message():
    """AllowFallbackToLocal() == false: the executor's NotSupported status (with the library's own message) is the job's status"""
    ops, opts = S.ALL["basic_bottommost"]()
    with tempfile.TemporaryDirectory(prefix="b200c_mock_") as d:
        with open(os.path.join(d, "ops.bin"), "wb") as f:
            f.write(ops.bytes())
        args = [MOCK_BIN, os.path.join(d, "ops.bin"), os.path.join(d, "w"), "executor=b200"] + [f"{k}={v}" for k, v in opts.items()]
        r = subprocess.run(args, capture_output=True, text=True, env=dict(os.environ, B200C_MOCK_DUMP=os.path.join(d, "dump.jsonl")))
    assert r.returncode != 0
    assert "mock library: records the job" in r.stderr and "Not implemented" in r.stderr  # Status::NotSupported prints as "Not implemented"


@pytest.mark.parametrize("code,status_text", [(4, "Corruption"), (1, "Invalid argument"), (6, "Memory limit"), (3, "Operation aborted"),
                                              (2, "Operation aborted")])
def test_library_status_codes_become_the_reference_statuses(code, status_text):
    """FromB200(): enum b200c_status -> Status; the job fails with that status and the library's message"""
    ops, opts = S.ALL["tiny"]()
    with tempfile.TemporaryDirectory(prefix="b200c_mock_") as d:
        with open(os.path.join(d, "ops.bin"), "wb") as f:
            f.write(ops.bytes())
        args = [MOCK_BIN, os.path.join(d, "ops.bin"), os.path.join(d, "w"), "executor=b200"] + [f"{k}={v}" for k, v in opts.items()]
        r = subprocess.run(args, capture_output=True, text=True, env=dict(os.environ, B200C_MOCK_FAIL=str(code)))
    assert r.returncode != 0
    assert status_text in r.stderr and "injected failure" in r.stderr, r.stderr[-500:]
