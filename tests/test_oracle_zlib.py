"""Compressed inputs (kZlibCompression, the one codec this image can pin: zlib is the only compression library present and the reference
is compiled with -DZLIB for it).  The job's input files are written by the unmodified reference with zlib block compression -- data
blocks and index block, UncompressBlockData table/format.cc:511, Zlib_Compress / Zlib_Uncompress util/compression.h:746-924 -- its
output level is uncompressed.  The CPU oracle (which inflates with zlib itself) must reproduce the reference's output files byte for
byte, on the committed fixtures and on live seeded runs; the fixtures really hold compressed blocks of both kinds (inflated and left
uncompressed because they did not shrink)."""
import pytest

import helpers as H
import scenarios as S
import sstfmt


def _block_types(data):
    t = sstfmt.parse_sst(data)
    types = [data[h[0] + h[1]] for _, h in t["index"]]
    io, isz = t["footer"]["index"]
    return types, data[io + isz]


@pytest.mark.parametrize("case", sorted(S.ZLIB))
def test_fixture_inputs_hold_compressed_and_uncompressed_blocks(case):
    g = H.load_golden(case)
    kinds, index_kinds = set(), set()
    for data in g["inputs"]:
        assert sstfmt.parse_sst(data)["properties"]["rocksdb.compression"] == b"Zlib"
        tys, ity = _block_types(data)
        kinds |= set(tys)
        index_kinds.add(ity)
    assert kinds == {0, 2}, kinds       # blocks that shrank and blocks that stayed as they were
    assert 2 in index_kinds             # enable_index_compression: the index block goes through the same WriteBlock
    for data in g["outputs"]:
        assert set(_block_types(data)[0]) == {0} and sstfmt.parse_sst(data)["properties"]["rocksdb.compression"] == b"NoCompression"


@pytest.mark.parametrize("case", sorted(S.ZLIB))
def test_oracle_reproduces_the_reference_on_compressed_fixtures(case):
    g = H.load_golden(case)
    p = H.params_from_reference(g)
    files, metas, st = H.oracle_compact(p, g["inputs"])
    assert [len(f) for f in files] == [len(f) for f in g["outputs"]]
    assert files == g["outputs"]
    for k in H.STAT_KEYS:
        assert getattr(st, k) == g["manifest"]["stats"][k], k


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", sorted(S.ZLIB))
@pytest.mark.parametrize("seed,extra", [(5, {}), (6, dict(index_compression=0)), (7, dict(block_size=1024, restart_interval=4)),
                                        (8, dict(checksum="crc32c", format_version=4))])
def test_oracle_matches_live_reference_on_compressed_inputs(case, seed, extra):
    ops, opts = S.ZLIB[case](seed=seed, n=700, nruns=3)
    ref = H.run_reference(ops, **dict(opts, **extra))
    assert any(2 in _block_types(d)[0] for d in ref["inputs"])
    p = H.params_from_reference(ref)
    files, metas, st = H.oracle_compact(p, ref["inputs"])
    assert files == ref["outputs"]
    for k in H.STAT_KEYS:
        assert getattr(st, k) == ref["manifest"]["stats"][k], k
