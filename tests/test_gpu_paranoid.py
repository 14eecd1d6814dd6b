"""paranoid_file_checks on the device (GPU box).  The reference re-reads every compaction output and compares an OutputValidator hash
of its keys and values with the one taken while writing (CompactionJob::Run, db/compaction/compaction_job.cc:829-853;
db/output_validator.cc:31-69); on the RunRemote branch it cannot do that itself (`precalculated_hash = 0`, :1065-1068), so
CompactionParams::paranoid_file_checks tells the executor to.  The product decodes its finished output images again (block checksums
verified) and compares every key, trailer and value with what the encoder consumed."""
import os
import subprocess
import sys

import pytest

try:
    import torch  # noqa: F401
except Exception:  # pragma: no cover
    torch = None

import gp_cases
import helpers as H
import scenarios as S

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("case", ["basic_bottommost", "snapshots_nonbottom", "varlen_keys", "crc32c_small_blocks",
                                  "same_user_key_across_blocks", "tiny", "all_deleted", "cfg3_mini", "ttl_filter", "output_level0"])
def test_paranoid_run_writes_the_same_files(case):
    from gpu_harness import run_product
    g = H.load_golden(case)
    p = H.params_from_reference(g)
    files, _, st = run_product(p, g["inputs"], paranoid_file_checks=1)
    assert files == g["outputs"]
    for k in H.STAT_KEYS:
        assert getattr(st, k) == g["manifest"]["stats"][k], k


def test_paranoid_run_with_grandparent_cuts_ranges_and_large_values():
    from gpu_harness import run_product
    p, inputs = gp_cases.build(**gp_cases.CASES["large_values"])
    want, _, _ = H.oracle_compact(p, inputs)
    assert run_product(p, inputs, paranoid_file_checks=1)[0] == want
    p, inputs = gp_cases.build(**gp_cases.CASES["many_large_grandparents"])
    p.range_start = p.grandparents[len(p.grandparents) // 3][0]
    want, _, _ = H.oracle_compact(p, inputs)
    assert run_product(p, inputs, paranoid_file_checks=1, device_inputs=True)[0] == want


_DAMAGE = """
import sys
sys.path.insert(0, {tests!r})
import helpers as H
from gpu_harness import run_product
import toplingdb_b200 as T
g = H.load_golden("basic_bottommost")
p = H.params_from_reference(g)
try:
    run_product(p, g["inputs"], paranoid_file_checks={paranoid})
    print("OK")
except T.B200cError as e:
    print("ERR", e.code == T.native.ERR_CORRUPTION, str(e))
"""


@pytest.mark.parametrize("offset,what", [(100, "a key or value byte of the first data block"), (4000, "a later data block"),
                                         (2, "an entry header")])
def test_damaged_output_is_caught(offset, what):
    """B200C_TEST_FLIP_OUTPUT_BYTE flips one byte of output 0 after the encoder finished and before the read-back (the hook only
    exists on the paranoid path); run in a child process so the environment variable does not leak into other tests"""
    env = dict(os.environ, B200C_TEST_FLIP_OUTPUT_BYTE=str(offset))
    code = _DAMAGE.format(tests=os.path.join(ROOT, "tests"), paranoid=1)
    out = subprocess.run([sys.executable, "-c", code], env=env, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.stdout.startswith("ERR True"), (what, out.stdout, out.stderr[-2000:])
    assert "Paranoid" in out.stdout or "checksum" in out.stdout
    # without paranoid_file_checks nothing is read back, so the hook is inert and the job succeeds
    code = _DAMAGE.format(tests=os.path.join(ROOT, "tests"), paranoid=0)
    out = subprocess.run([sys.executable, "-c", code], env=env, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.stdout.startswith("OK"), (out.stdout, out.stderr[-2000:])


def test_reference_db_with_paranoid_file_checks_runs_through_the_executor():
    if not (os.path.exists(H.REF_BIN) and os.path.exists(H.REF_B200_BIN)):
        pytest.fail("oracle/_ref/ref_compact(_b200) missing")
    ops, opts = S.ALL["cfg3_mini"]()
    want = H.run_reference(ops, paranoid=1, **opts)
    got = H.run_reference(ops, binary=H.REF_B200_BIN, executor="b200", paranoid=1, **opts)
    assert got["manifest"]["remote_compact_read_bytes"] > 0
    assert (got["manifest"]["scan_count"], got["manifest"]["scan_digest"]) == (want["manifest"]["scan_count"], want["manifest"]["scan_digest"])
    assert H.sizes_without_file_number(got["outputs"]) == H.sizes_without_file_number(want["outputs"])
