"""CPU-side checks of the drop-in boundary: the library builds, loads, exports every symbol include/b200c.h declares,
and refuses to run without a device (there is no CPU data path)."""
import ctypes as C
import os
import re

import pytest

import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import toplingdb_b200 as T
    L = T.load_library()
    hdr = open(os.path.join(ROOT, "include", "b200c.h")).read()
    declared = set(re.findall(r"B200C_API [^;(]*?\b(b200c_[a-z_]+)\(", hdr))
    assert len(declared) >= 16
    for name in declared:
        assert hasattr(L, name), name
    assert set(T.native.EXPORTS) == declared
    assert L.b200c_abi_version() == 6


def test_params_defaults_match_reference_defaults():
    import toplingdb_b200 as T
    p = T.native.Params()
    T.lib().b200c_params_init(C.byref(p))
    # include/rocksdb/table.h:237-564, advanced_options.h:599
    assert (p.block_size, p.block_size_deviation, p.block_restart_interval, p.index_block_restart_interval) == (4096, 10, 16, 1)
    assert (p.format_version, p.checksum, p.max_output_file_size) == (5, 4, 64 << 20)


def test_no_cpu_fallback_without_a_device():
    import torch
    import toplingdb_b200 as T
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(T.B200cError) as ei:
        T.CompactionJob()
    assert ei.value.code == T.native.ERR_NO_DEVICE
    with pytest.raises(T.B200cError):
        T.block_checksums("xxh3", [b"abc"])


def test_product_does_not_reference_the_oracle():
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "toplingdb_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                s = open(os.path.join(dp, f), errors="replace").read()
                if "liboracle" in s or "compaction_oracle" in s or "oracle/" in s:
                    bad.append(f)
    assert not bad, bad


@pytest.mark.skipif(not (os.path.exists(H.REF_BIN) and os.path.exists(H.REF_B200_BIN)), reason="oracle/_ref not built")
def test_plugin_without_a_device_lets_the_reference_run_the_job_itself():
    """No CUDA device: B200CompactionExecutorFactory::ShouldRunLocal() must answer true (compaction_executor.h:162), so the
    reference's CompactionJob::Run takes RunLocal() (compaction_job.cc:645-647) -- the plugin never computes anything on
    the CPU itself.  Same files as a run without the plugin, and no remote-compaction bytes accounted."""
    import scenarios as S
    if os.path.exists("/dev/nvidia0"):
        pytest.skip("a CUDA device is present: covered by tests/test_gpu_plugin_integration.py")
    ops, opts = S.ALL["cfg2_mini"](per_run=300)
    want = H.run_reference(ops, **opts)
    got = H.run_reference(ops, binary=H.REF_B200_BIN, executor="b200", **opts)
    assert got["manifest"]["executor"] == "B200Compact"
    assert got["manifest"]["remote_compact_read_bytes"] == 0
    assert (got["manifest"]["scan_count"], got["manifest"]["scan_digest"]) == (want["manifest"]["scan_count"], want["manifest"]["scan_digest"])
    assert [len(o) for o in got["outputs"]] == [len(o) for o in want["outputs"]]
