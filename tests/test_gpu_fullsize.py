"""BASELINE.json-size jobs (GPU box).  Byte-exact: the CPU oracle compacts the SAME full-size input images (cfg2, cfg3 with
bottommost on / off, cfg5 -- the jobs bench.py times, staged by the same function) and every output file, FileMetaData and the
statistics must be equal; the bench inputs themselves must equal what the oracle's table builder writes for the same entries.
Plus the size-independent properties of the path:

  * conservation: with disjoint keys (cfg2) every input record comes out exactly once; with overlap + tombstones (cfg3)
    num_input = num_output + hidden + obsolete tombstones, as CompactionJobStats accounts them (compaction_job.cc:1290-1322)
  * sortedness / non-overlap of the output files, and the file-size rule (all files but the last reach the target)
  * re-reading: the product's own decoder accepts every output block (checksums verified) and returns the entry count
  * idempotence: compacting the outputs again (same parameters) reproduces them byte for byte
  * oracle spot check: the first output file's first data blocks equal the CPU oracle's output on the matching key range"""
import struct

import pytest

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

import helpers as H
import sstfmt

pytestmark = pytest.mark.gpu

COMMON = dict(output_level=1, max_output_file_size=64 << 20, file_creation_times=[1700000000], first_file_number=1, db_id="full",
              db_session_id="FULLSIZE", db_host_id="b200")


def _stage(k, n_run, vlen, overlap=0.0, del_frac=0.0, seed=5):
    from toplingdb_b200 import synth
    images, kv_bytes = synth.stage_runs(n_run * k, k, vlen, key_base=7, seed=seed, overlap=overlap, del_frac=del_frac, device_index=0)
    return images, kv_bytes


def _run(images, bottommost, output_mem="device"):
    import toplingdb_b200 as T
    job = T.CompactionJob(output_mem=output_mem, bottommost_level=bottommost, **COMMON)
    for i, img in enumerate(images):
        job.add_input(img, level=0, file_number=100 + i)
    job.run()
    return job


def _check_files(job):
    prev = None
    n = job.output_count()
    sizes = []
    for i in range(n):
        m = job.output_meta(i)
        a = bytes(m.smallest_ikey[:m.smallest_ikey_len - 8])
        b = bytes(m.largest_ikey[:m.largest_ikey_len - 8])
        assert a <= b
        assert prev is None or prev < a, "output files overlap / out of order"
        prev = b
        sizes.append(m.data_size)
    assert all(s >= COMMON["max_output_file_size"] for s in sizes[:-1]), "a file was cut before reaching the target size"
    return n


def _outputs_as_device_tensors(job):
    outs = []
    for i in range(job.output_count()):
        m = job.output_meta(i)
        t = torch.empty(m.file_size, dtype=torch.uint8, device="cuda")
        job.output_read_into(i, t)
        outs.append(t)
    return outs


def test_cfg2_full_size_properties_and_idempotence():
    """configs[1]: 8 x 256 MiB runs, 16 B keys / 32 B values, disjoint keys."""
    k, vlen = 8, 32
    n_run = (256 << 20) // (24 + vlen)
    images, _ = _stage(k, n_run, vlen)
    job = _run(images, bottommost=True)
    st = job.stats()
    assert st.num_input_records == k * n_run
    assert st.num_output_records == st.num_input_records  # disjoint keys, no tombstones: nothing may be dropped
    assert st.num_records_replaced == 0 and st.num_expired_deletion_records == 0
    nfiles = _check_files(job)
    assert sum(job.output_meta(i).num_entries for i in range(nfiles)) == st.num_output_records
    outs = _outputs_as_device_tensors(job)
    del images
    job.close()
    torch.cuda.empty_cache()
    # idempotence: the outputs, fed back as (disjoint, ordered) runs, come out byte-identical
    job2 = _run(outs, bottommost=True)
    assert job2.output_count() == nfiles
    assert job2.stats().num_output_records == st.num_output_records
    for i, t in enumerate(outs):
        m = job2.output_meta(i)
        assert m.file_size == t.numel()
        t2 = torch.empty_like(t)
        job2.output_read_into(i, t2)
        assert torch.equal(t, t2), f"file {i} changed when compacted again"
    # oracle spot check: the block trailers of the first file verify against the CPU oracle's checksum
    L = H.oracle()
    full = outs[0].cpu().numpy().tobytes()
    ft = sstfmt.parse_footer(full)
    assert ft["format_version"] == 5 and ft["checksum_type"] == 4
    iblock, _, _ = sstfmt.read_block(full, ft["index"])
    handles = []
    prev = None
    for kk, pp, shared in sstfmt.block_entries(iblock, value_delta=True)[:200]:
        if shared == 0:
            o, pp = sstfmt.varint(iblock, pp)
            sz, pp = sstfmt.varint(iblock, pp)
        else:
            d, pp = sstfmt.varint(iblock, pp)
            sz = prev[1] + sstfmt.zigzag(d)
            o = prev[0] + prev[1] + 5
        prev = (o, sz)
        handles.append(prev)
    assert handles and handles[0][0] == 0
    for o, sz in handles:
        want = L.orc_block_checksum(H.CKSUM["xxh3"], full[o:o + sz], sz, full[o + sz])
        assert want == struct.unpack_from("<I", full, o + sz + 1)[0]
    job2.close()


def test_cfg3_like_conservation_with_overlap_and_tombstones():
    """configs[2] shape at 1/4 size: 16 runs, 30 % of keys in two runs, 10 % tombstones, 256 B values."""
    k, vlen = 16, 256
    n_run = (64 << 20) // (24 + vlen)
    images, _ = _stage(k, n_run, vlen, overlap=0.3, del_frac=0.1, seed=9)
    for bottommost in (True, False):
        job = _run(images, bottommost=bottommost)
        st = job.stats()
        dropped = st.num_records_replaced + st.num_expired_deletion_records
        assert st.num_input_records == st.num_output_records + dropped
        assert st.num_records_replaced > 0
        if bottommost:
            assert st.num_expired_deletion_records > 0
        else:
            assert st.num_expired_deletion_records == 0  # tombstones must survive above the bottom level
        n = _check_files(job)
        assert sum(job.output_meta(i).num_entries for i in range(n)) == st.num_output_records
        job.close()


def test_cfg1_shape_against_the_reference_itself():
    """configs[0] (the reference's own CPU-runnable case): 4 L0 BlockBasedTable SSTs from 1 M random 16-byte keys with
    100-byte values (db_bench fillrandom shape: keys repeat within and across files), compacted to the bottommost level.
    Full size, and the checker is the unmodified reference run right here (oracle/_ref travels to the GPU box): every output
    file byte for byte, CompactionJobStats, FileMetaData."""
    import os
    import random
    import struct
    if not os.path.exists(H.REF_BIN):
        pytest.fail("oracle/_ref/ref_compact missing: run __graft_entry__.build() where /root/reference exists")
    from gpu_harness import run_product
    rnd = random.Random(1)
    n, files = 1_000_000, 4
    ops = H.Ops()
    per = n // files
    for _ in range(files):
        for _ in range(per):
            ops.put(struct.pack(">QQ", 0, rnd.randrange(n)), rnd.randbytes(100))
        ops.flush()
    ref = H.run_reference(ops, target_file_size=64 << 20)
    man = ref["manifest"]
    assert len(ref["inputs"]) == files and man["stats"]["num_input_records"] > 0.9 * n * 0.8
    p = H.params_from_reference(ref)
    got, metas, st = run_product(p, ref["inputs"])
    assert [len(f) for f in got] == [len(f) for f in ref["outputs"]]
    for i, (a, b) in enumerate(zip(got, ref["outputs"])):
        assert a == b, f"output {i} differs at byte {next(j for j in range(len(a)) if a[j] != b[j])}"
    for k in H.STAT_KEYS:
        assert getattr(st, k) == man["stats"][k], k
    for m, want in zip(metas, man["outputs"]):
        assert (m.file_size, m.num_entries, m.num_deletions) == (want["size"], want["num_entries"], want["num_deletions"])


# ---------------------------------------------------------------------------------------------------------------------------------
# Byte-exact against the CPU oracle at the benchmarked sizes (the jobs bench.py times: same staging function, same parameters)
def _bench_params(bottommost):
    from toplingdb_b200 import synth
    return H.Params(bottommost_level=bool(bottommost), creation_time=0, **synth.BENCH_JOB)


def _compare_with_oracle(workload, bottommost, scale=1.0):
    import numpy as np
    from gpu_harness import job_from_params
    from toplingdb_b200 import synth
    images, _ = synth.stage_bench_inputs(workload, rank=0, scale=scale)
    p = _bench_params(bottommost)
    job = job_from_params(p, output_mem="device")
    for i, img in enumerate(images):
        job.add_input(img, level=0, file_number=100 + i)
    job.run()
    st = job.stats()
    n = job.output_count()
    got = []
    for i in range(n):
        m = job.output_meta(i)
        t = torch.empty(m.file_size, dtype=torch.uint8, device="cuda")
        job.output_read_into(i, t)
        got.append((t.cpu().numpy(), m))
    digest = synth.outputs_digest(job)
    job.close()
    host_inputs = [t.cpu().numpy().tobytes() for t in images]
    del images
    torch.cuda.empty_cache()
    ofiles, ometas, ost = H.oracle_compact(p, host_inputs)
    del host_inputs
    assert n == len(ofiles), (n, len(ofiles))
    for i, ((a, m), b, om) in enumerate(zip(got, ofiles, ometas)):
        bb = np.frombuffer(b, dtype=np.uint8)
        assert a.size == bb.size, f"file {i}: {a.size} vs oracle {bb.size} bytes"
        if not np.array_equal(a, bb):
            at = int(np.flatnonzero(a != bb)[0])
            raise AssertionError(f"{workload} bottommost={bottommost}: output {i} differs from the oracle at byte {at} of {a.size}")
        assert (m.file_size, m.num_entries, m.num_deletions, m.raw_key_size, m.raw_value_size, m.num_data_blocks, m.smallest_seqno,
                m.largest_seqno) == (om.file_size, om.num_entries, om.num_deletions, om.raw_key_size, om.raw_value_size, om.num_data_blocks,
                                     om.smallest_seqno, om.largest_seqno), f"FileMetaData of output {i}"
        assert bytes(m.smallest_ikey[:m.smallest_ikey_len]) == bytes(om.smallest[:om.smallest_len])
        assert bytes(m.largest_ikey[:m.largest_ikey_len]) == bytes(om.largest[:om.largest_len])
    for k in H.STAT_KEYS:
        assert getattr(st, k) == getattr(ost, k), k
    assert digest == synth.files_digest(ofiles)  # the digest bench.py prints is the digest of the oracle's files
    return digest


def test_cfg2_full_size_bytes_equal_oracle():
    """configs[1] as bench.py runs it: 8 x 256 MiB, 16 B / 32 B; every output byte against the CPU oracle."""
    d = _compare_with_oracle("cfg2", bottommost=False)
    _check_committed_digest("cfg2", d)


@pytest.mark.parametrize("bottommost", [True, False])
def test_cfg3_full_size_bytes_equal_oracle(bottommost):
    """configs[2]: 16 x 256 MiB, 30 % overlap, 10 % tombstones, 16 B / 256 B; bottommost on and off."""
    d = _compare_with_oracle("cfg3", bottommost=bottommost)
    if bottommost:
        _check_committed_digest("cfg3", d)


def test_cfg5_full_size_bytes_equal_oracle():
    """one of configs[4]'s sub-compactions: 4 x 64 MiB, 16 B / 128 B."""
    d = _compare_with_oracle("cfg5", bottommost=False)
    _check_committed_digest("cfg5", d)


def _check_committed_digest(workload, digest):
    """tests/golden/bench_digests.json holds the oracle's digests of the bench jobs (rank 0, scale 1), written by
    tools/make_bench_digests.py on a GPU box; bench.py compares its own outputs with them.  A stale file must not pass silently."""
    import json
    import os
    path = os.path.join(H.GOLDEN_DIR, "bench_digests.json")
    if os.path.exists(path):
        want = json.load(open(path)).get(f"{workload}:rank0:scale1.0")
        if want is not None:
            assert want == digest, f"tests/golden/bench_digests.json is stale for {workload}: regenerate with tools/make_bench_digests.py"


def _run_kvstream(n_total, k, r, vlen, key_base, seed, overlap=0.0, del_frac=0.0):
    """kv stream (helpers.kvstream format) of run r as synth.stage_runs encodes it, assembled on the host with numpy"""
    import numpy as np
    from toplingdb_b200 import synth
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    values = torch.randint(0, 256, (max(1, n_total * vlen),), dtype=torch.uint8, device=dev, generator=g)
    n, pfx, tr, vref, meta = synth.make_run_columns(n_total, k, r, vlen, key_base, values, overlap, del_frac, seed, dev)
    vl = (meta.to(torch.int64) & ((1 << 27) - 1)).cpu().numpy()
    voff = (vref - values.data_ptr()).cpu().numpy()
    hi = pfx[:, 0].cpu().numpy().view(np.uint64).astype(">u8")
    lo = pfx[:, 1].cpu().numpy().view(np.uint64).astype(">u8")
    trn = tr.cpu().numpy().view(np.uint64).astype("<u8")
    vals = values.cpu().numpy()
    if (vl == vlen).all():
        rec = np.zeros(n, dtype=[("kl", "<u4"), ("vl", "<u4"), ("hi", ">u8"), ("lo", ">u8"), ("tr", "<u8"), ("v", "u1", (vlen,))])
        rec["kl"], rec["vl"], rec["hi"], rec["lo"], rec["tr"] = 24, vlen, hi, lo, trn
        assert (voff % vlen == 0).all()
        rec["v"] = vals.reshape(-1, vlen)[voff // vlen]
        return rec.tobytes()
    out = bytearray()
    hib, lob, trb, valb = hi.tobytes(), lo.tobytes(), trn.tobytes(), vals.tobytes()  # array-level tobytes keeps the declared byte order
    for i in range(n):
        out += struct.pack("<II", 24, int(vl[i])) + hib[8 * i:8 * i + 8] + lob[8 * i:8 * i + 8] + trb[8 * i:8 * i + 8] + valb[voff[i]:voff[i] + vl[i]]
    return bytes(out)


def test_bench_input_images_equal_oracle_built_tables():
    """The bench inputs are written by the product's own encoder (synth.stage_runs -> b200c_job_encode_columns); an encoder / decoder
    symmetric bug would be invisible to every round-trip check.  Here each image must equal, byte for byte, the table the oracle's
    BlockBasedTableBuilder restatement (block_based_table_builder.cc:961-1133) writes for the same entries: one full-size cfg2 run
    (4.8 M entries, 256 MiB) and a small cfg3-shaped run with overlap substitutions and tombstones."""
    import numpy as np
    from toplingdb_b200 import sharding, synth
    cases = [("cfg2", 1.0, 3), ("cfg3", 1.0 / 64, 0), ("cfg3", 1.0 / 64, 15)]
    for workload, scale, r in cases:
        w = synth.WORKLOADS[workload]
        _, n_total = synth.bench_shape(workload, scale)
        key_base = sharding.key_range_base(0, n_total)
        images, _ = synth.stage_runs(n_total, w["k"], w["vlen"], key_base=key_base, seed=2, overlap=w["overlap"], del_frac=w["del_frac"])
        img = images[r].cpu().numpy()
        del images
        torch.cuda.empty_cache()
        kv = _run_kvstream(n_total, w["k"], r, w["vlen"], key_base, 2, w["overlap"], w["del_frac"])
        p = H.Params(output_level=0, creation_time=0, db_id="", db_session_id="", db_host_id="", file_creation_times=[1700000000],
                     first_file_number=1000 + r)
        want = np.frombuffer(H.oracle_build_sst(p, kv), dtype=np.uint8)
        assert img.size == want.size, (workload, r, img.size, want.size)
        assert np.array_equal(img, want), f"{workload} run {r}: staged image differs from the oracle-built table at byte {int(np.flatnonzero(img != want)[0])}"
