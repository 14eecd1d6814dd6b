"""output tables must be readable by the reference's stock tools.  oracle/_ref/ref_sst_check drives the reference's own
SstFileDumper (what `sst_dump --command=verify|scan` uses, tools/sst_dump_tool.cc) and the public stand-alone SstFileReader over a
file: every block checksum verified, the whole file scanned, table properties read.  Run here on the committed reference fixtures --
the GPU parity tests prove the device writes these very bytes (tests/test_gpu_parity.py::test_full_job_matches_reference_fixture) --
and on oracle-built tables with a Bloom filter block, which the device reproduces byte for byte as well (tests/test_gpu_bloom.py)."""
import json
import os
import subprocess

import pytest

import helpers as H
import sstfmt

TOOL = os.path.join(H.ROOT, "oracle", "_ref", "ref_sst_check")
pytestmark = pytest.mark.skipif(not os.path.exists(TOOL), reason="oracle/_ref not built (needs /root/reference)")


def _fnv(pairs):
    h = 1469598103934665603
    for x in pairs:
        for b in x:
            h = ((h ^ b) * 1099511628211) & ((1 << 64) - 1)
        h = ((h ^ len(x)) * 1099511628211) & ((1 << 64) - 1)
    return h


def _check(path, data):
    out = json.loads(subprocess.check_output([TOOL, path], text=True))
    t = sstfmt.parse_sst(data)
    assert out["dumper_entries"] == len(t["entries"]) == out["num_entries"]
    # what an application sees through SstFileReader: the newest version of every user key, unless it is a tombstone
    visible, last = [], None
    for ik, v in t["entries"]:
        uk = ik[:-8]
        if uk != last and ik[-8] == 1:
            visible += [uk, v]
        last = uk
    assert out["reader_entries"] == len(visible) // 2
    assert int(out["reader_digest"], 16) == _fnv(visible)
    assert out["num_data_blocks"] == len(t["index"])
    return out


@pytest.mark.parametrize("case", [c for c in H.golden_cases()])
def test_reference_tools_read_the_fixture_outputs(case, tmp_path):
    g = H.load_golden(case)
    for i, data in enumerate(g["outputs"]):
        p = tmp_path / f"{i:06d}.sst"
        p.write_bytes(data)
        _check(str(p), data)


@pytest.mark.parametrize("millibits", [10000, 6500])
def test_reference_tools_read_tables_with_a_bloom_filter_block(millibits, tmp_path):
    g = H.load_golden("cfg3_mini")
    p = H.params_from_reference(g)
    p.bloom_millibits_per_key = millibits
    files, _, _ = H.oracle_compact(p, g["inputs"])
    for i, data in enumerate(files):
        f = tmp_path / f"{i:06d}.sst"
        f.write_bytes(data)
        out = _check(str(f), data)
        assert out["filter_policy_name"] == "bloomfilter" and out["filter_size"] > 0
        assert out["num_filter_entries"] == len({ik[:-8] for ik, _ in sstfmt.parse_sst(data)["entries"]})
