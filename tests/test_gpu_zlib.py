"""Compressed inputs on the device (GPU box): kZlibCompression data blocks are inflated by inflate_blocks_kernel (one thread per block,
toplingdb_b200/csrc/inflate_rules.h) into an arena in front of the block decoder, their checksums verified over the stored bytes; a
compressed index block is inflated on the host with the same decoder.  Against the reference-written fixtures, the CPU oracle on
seeded jobs (host and device-resident inputs, key ranges), a corrupted compressed block, and the unmodified reference DB running the
job through the executor plugin."""
import os

import pytest

try:
    import torch  # noqa: F401
except Exception:  # pragma: no cover
    torch = None

import helpers as H
import scenarios as S
import sstfmt
import toplingdb_b200 as T

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", sorted(S.ZLIB))
@pytest.mark.parametrize("device_inputs", [False, True])
def test_device_reproduces_the_reference_on_compressed_fixtures(case, device_inputs):
    from gpu_harness import run_product
    g = H.load_golden(case)
    p = H.params_from_reference(g)
    files, metas, st = run_product(p, g["inputs"], device_inputs=device_inputs)
    assert [len(f) for f in files] == [len(f) for f in g["outputs"]]
    assert files == g["outputs"]
    for k in H.STAT_KEYS:
        assert getattr(st, k) == g["manifest"]["stats"][k], k


@pytest.mark.parametrize("seed,extra", [(31, {}), (32, dict(block_size=1024, restart_interval=4)), (33, dict(checksum="crc32c", format_version=4)),
                                        (34, dict(index_compression=0))])
def test_seeded_compressed_jobs_match_the_oracle(seed, extra):
    from gpu_harness import run_product
    if not H.have_ref():
        pytest.fail("oracle/_ref missing: run __graft_entry__.build() where /root/reference exists")
    ops, opts = S.ZLIB["zlib_inputs"](seed=seed, n=4000, nruns=5)
    ref = H.run_reference(ops, **dict(opts, **extra))
    p = H.params_from_reference(ref)
    files, metas, st = run_product(p, ref["inputs"])
    assert files == ref["outputs"]
    # a key range: only the blocks in range are inflated
    keys = sorted({ik[:-8] for d in ref["inputs"][:1] for ik, _ in sstfmt.parse_sst(d)["entries"]})
    p.range_start, p.range_end = keys[len(keys) // 3], keys[2 * len(keys) // 3]
    want, _, wst = H.oracle_compact(p, ref["inputs"])
    got, _, gst = run_product(p, ref["inputs"])
    assert got == want and gst.num_input_records == wst.num_input_records


def test_damaged_compressed_block_is_reported():
    from gpu_harness import job_from_params
    g = H.load_golden("zlib_inputs")
    p = H.params_from_reference(g)
    data = bytearray(g["inputs"][0])
    t = sstfmt.parse_sst(bytes(data))
    off, size = next(h for _, h in t["index"] if data[h[0] + h[1]] == 2)
    data[off + size // 2] ^= 0x40  # inside the deflate stream
    job = job_from_params(p)
    job.add_input(bytes(data), level=0, file_number=1)
    for i, d in enumerate(g["inputs"][1:]):
        job.add_input(d, level=0, file_number=2 + i)
    with pytest.raises(T.B200cError) as ei:
        job.run()
    assert ei.value.code == T.native.ERR_CORRUPTION  # the checksum over the stored bytes catches it before the inflater does
    job.close()
    job = job_from_params(p, verify_input_checksums=0)  # without the check the inflater (or the block parser) must still refuse it
    job.add_input(bytes(data), level=0, file_number=1)
    for i, d in enumerate(g["inputs"][1:]):
        job.add_input(d, level=0, file_number=2 + i)
    try:
        job.run()
        got = job.outputs()
        assert got != g["outputs"]  # (a flipped literal can inflate to the announced size: then the bytes differ, nothing crashed)
    except T.B200cError as e:
        assert e.code == T.native.ERR_CORRUPTION
    job.close()


def test_reference_db_compacts_compressed_inputs_through_the_b200_executor():
    if not (os.path.exists(H.REF_BIN) and os.path.exists(H.REF_B200_BIN)):
        pytest.fail("oracle/_ref/ref_compact(_b200) missing: run __graft_entry__.build() where /root/reference exists")
    ops, opts = S.ZLIB["zlib_inputs"](seed=41, n=2500, nruns=4)
    want = H.run_reference(ops, **opts)
    got = H.run_reference(ops, binary=H.REF_B200_BIN, executor="b200", **opts)
    gm, wm = got["manifest"], want["manifest"]
    assert gm["executor"] == "B200Compact" and gm["remote_compact_read_bytes"] > 0
    assert (gm["scan_count"], gm["scan_digest"]) == (wm["scan_count"], wm["scan_digest"])
    assert H.sizes_without_file_number(got["outputs"]) == H.sizes_without_file_number(want["outputs"])
    ge = [e for f in got["outputs"] for e in sstfmt.parse_sst(f)["entries"]]
    we = [e for f in want["outputs"] for e in sstfmt.parse_sst(f)["entries"]]
    assert ge == we
    for k in H.STAT_KEYS:
        assert gm["stats"][k] == wm["stats"][k], k
