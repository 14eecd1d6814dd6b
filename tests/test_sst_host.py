"""The product's host-side table code (toplingdb_b200/csrc/sst_host.cc) compiled for the CPU on its own (tests/native/tail_sim.cc):
 * build_output_tail: given the numbers the device computes per output file, the properties block, metaindex block and footer must be
   the bytes the reference wrote behind the index block (PropertyBlockBuilder meta_blocks.cc:54-175, MetaIndexBuilder :35-49,
   FooterBuilder format.cc:211-259) -- on every committed reference fixture, and on tables with a Bloom filter block;
 * parse_footer / parse_metaindex / parse_properties: what the job reads from its input files, against the Python parser.
CPU only; the same code runs inside libb200c.so on the GPU box (tests/test_gpu_parity.py compares whole files there)."""
import ctypes as C
import os
import subprocess

import pytest

import helpers as H
import scenarios as S
import sstfmt

ROOT = H.ROOT


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("tail") / "tail_sim.so")
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-std=c++17", "-I" + os.path.join(ROOT, "toplingdb_b200", "csrc"),
                           os.path.join(ROOT, "tests", "native", "tail_sim.cc"), "-o", so])
    L = C.CDLL(so)
    L.tail_sim_build.restype = C.c_uint64
    return L


def _u(props, name):
    return sstfmt.prop_u64(props, name)


def _rebuild_tail(shim, data):
    t = sstfmt.parse_sst(data)
    pr, ft = t["properties"], t["footer"]
    io, isz = ft["index"]
    s = lambda k: pr.get(k, b"")
    out = C.create_string_buffer(8192)
    n = shim.tail_sim_build(
        C.c_uint32(ft["checksum_type"]), C.c_uint32(ft["format_version"]), C.c_uint64(_u(pr, "rocksdb.data.size")), C.c_uint64(isz),
        C.c_uint64(_u(pr, "rocksdb.filter.size")), C.c_uint64(_u(pr, "rocksdb.num.filter_entries")), C.c_uint64(_u(pr, "rocksdb.num.entries")),
        C.c_uint64(_u(pr, "rocksdb.deleted.keys")), C.c_uint64(_u(pr, "rocksdb.raw.key.size")), C.c_uint64(_u(pr, "rocksdb.raw.value.size")),
        C.c_uint64(_u(pr, "rocksdb.num.data.blocks")), C.c_int(_u(pr, "rocksdb.index.key.is.user.key")),
        C.c_uint32(_u(pr, "rocksdb.column.family.id")), s("rocksdb.column.family.name"), s("rocksdb.creating.db.identity"),
        s("rocksdb.creating.session.identity"), s("rocksdb.creating.host.identity"), C.c_uint64(_u(pr, "rocksdb.creation.time")),
        C.c_uint64(_u(pr, "rocksdb.oldest.key.time")),
        C.c_uint64(_u(pr, "rocksdb.file.creation.time") if "rocksdb.file.creation.time" in pr else 0),
        C.c_uint64(_u(pr, "rocksdb.original.file.number")), out, C.c_uint64(8192))
    return out.raw[:n], data[io + isz + 5:]


@pytest.mark.parametrize("case", H.golden_cases())
def test_output_tail_is_what_the_reference_wrote(shim, case):
    g = H.load_golden(case)
    for data in g["outputs"] + g["inputs"]:  # the inputs were written by the reference's FlushJob with the same builder
        if sstfmt.parse_sst(data)["properties"].get("rocksdb.compression") != b"NoCompression":
            continue  # (compressed inputs of the zlib fixtures: the tail builder writes the properties of uncompressed tables only)
        got, want = _rebuild_tail(shim, data)
        assert got == want


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("case,opts", [("cfg3_mini", dict(bloom_bits=10)), ("crc32c_small_blocks", dict(bloom_bits=7)),
                                       ("basic_bottommost", dict(format_version=3)), ("snapshots", dict(format_version=4))])
def test_output_tail_with_filter_blocks_and_older_format_versions(shim, case, opts):
    ops, o = S.ALL[case]()
    ref = H.run_reference(ops, **dict(o, **opts))
    for data in ref["outputs"]:
        got, want = _rebuild_tail(shim, data)
        assert got == want


@pytest.mark.parametrize("case", ["basic_bottommost", "crc32c_small_blocks", "same_user_key_across_blocks", "tiny"])
def test_input_tail_parser_agrees_with_the_python_parser(shim, case):
    g = H.load_golden(case)
    for data in g["inputs"] + g["outputs"]:
        f = (C.c_uint64 * 12)()
        assert shim.tail_sim_parse(data, C.c_uint64(len(data)), f) == 0
        t = sstfmt.parse_sst(data)
        pr = t["properties"]
        assert (f[0], f[1]) == t["footer"]["index"]
        assert (f[2], f[3]) == t["metaindex"]["rocksdb.properties"]
        assert [f[4], f[5], f[6], f[7], f[8]] == [_u(pr, "rocksdb.num.entries"), _u(pr, "rocksdb.num.data.blocks"), _u(pr, "rocksdb.raw.key.size"),
                                                 _u(pr, "rocksdb.raw.value.size"), _u(pr, "rocksdb.data.size")]
        assert (f[9], f[10], f[11]) == (t["footer"]["checksum_type"], t["footer"]["format_version"], 0)
    assert shim.tail_sim_parse(b"x" * 100, C.c_uint64(100), (C.c_uint64 * 12)()) != 0  # bad magic
