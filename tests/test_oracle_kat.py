"""Known-answer vectors the reference's own tests hold for the leaf functions of the path, checked
against the CPU oracle (oracle/compaction_oracle.c).  Sources cited per test."""
import struct

import helpers as H


def _hex_le(v):
    return struct.pack("<I", v).hex().upper()


def test_crc32c_standard_results():
    # util/crc32c_test.cc:67-95 (rfc3720 B.4)
    L = H.oracle()
    assert L.orc_crc32c_value(bytes(32), 32) == 0x8A9136AA
    assert L.orc_crc32c_value(b"\xff" * 32, 32) == 0x62A8AB43
    assert L.orc_crc32c_value(bytes(range(32)), 32) == 0x46DD794E
    assert L.orc_crc32c_value(bytes(31 - i for i in range(32)), 32) == 0x113FDB5C
    data = bytes([0x01, 0xC0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0x14, 0, 0, 0, 0, 0, 0x04, 0,
                  0, 0, 0, 0x14, 0, 0, 0, 0x18, 0x28, 0, 0, 0, 0, 0, 0, 0, 0x02, 0, 0, 0, 0, 0, 0, 0])
    assert L.orc_crc32c_value(data, 48) == 0xD9963A56


def test_crc32c_mask():
    # util/crc32c_test.cc:107-113
    L = H.oracle()
    crc = L.orc_crc32c_value(b"foo", 3)
    assert L.orc_crc32c_mask(crc) != crc


def test_block_checksum_schemas():
    # table/table_test.cc:2303-2389 BuiltinChecksumTest.ChecksumSchemas
    L = H.oracle()
    b1 = b"This is a short block!"
    b2 = b"This is a long block!" * 100
    ct = (0, 1, 7)  # kNoCompression, kSnappyCompression, kZSTD
    want = {
        "crc32c": ("D8EA82A2", ["D28F2549", "052B2843", "46F8F711"], ["583F0355", "2F9B0A57", "ECE7DA1D"],
                   ["943EF0AB", "43A2EDB1", "00E53D63"]),
        "xxh3": ("00000000", ["C294D338", "1B174353", "2D0E20C8"], ["B37FB5E6", "6AFC258D", "5CE54616"],
                 ["FA2D482E", "23AED845", "15B7BBDE"]),
    }
    for name, (empty, w0, w1, w2) in want.items():
        t = H.CKSUM[name]
        assert _hex_le(L.orc_checksum(t, b"", 0)) == empty
        for body, ws in ((b"", w0), (b1, w1), (b2, w2)):
            for c, w in zip(ct, ws):
                assert _hex_le(L.orc_block_checksum(t, body, len(body), c)) == w, (name, len(body), c)
                full = body + bytes([c])
                assert _hex_le(L.orc_checksum(t, full, len(full))) == w


def test_xxh3_all_length_classes_against_block_checksum_identity():
    # XXH3 length classes 0, 1-3, 4-8, 9-16, 17-128, 129-240, >240 and the 1024-byte block boundary:
    # ComputeBuiltinChecksum(data) == ComputeBuiltinChecksumWithLastByte(data[:-1], data[-1]) (table_test.cc:2283-2287)
    L = H.oracle()
    import random
    rnd = random.Random(5)
    for n in list(range(0, 300)) + [1023, 1024, 1025, 2047, 2048, 2049, 4096, 5000]:
        d = rnd.randbytes(n + 1)
        assert L.orc_checksum(4, d, n + 1) == L.orc_block_checksum(4, d[:n], n, d[n])


def test_varint_encoding():
    # util/coding_test.cc Varint32/Varint64 round trips
    import ctypes as C
    L = H.oracle()
    import sstfmt
    for v in [0, 1, 127, 128, 255, 300, 16383, 16384, (1 << 32) - 1, 1 << 32, (1 << 63) + 5, (1 << 64) - 1]:
        buf = C.create_string_buffer(10)
        n = L.orc_put_varint64(buf, C.c_uint64(v))
        got, p = sstfmt.varint(buf.raw, 0)
        assert (got, p) == (v, n)


def test_internal_key_order():
    # db/dbformat_test.cc:55-135: user key ascending, then sequence DEscending, then type descending
    L = H.oracle()

    def less(a, b):
        return bool(L.orc_internal_key_less(a, len(a), b, len(b)))

    k = H.ikey
    assert less(k(b"foo", 100, 1), k(b"foo", 99, 1))
    assert less(k(b"foo", 100, 1), k(b"foo", 100, 0))
    assert less(k(b"bar", 1, 1), k(b"foo", 100, 1))
    assert less(k(b"foo", 1, 1), k(b"foo\x00", 100, 1))
    assert not less(k(b"foo", 5, 1), k(b"foo", 5, 1))
    assert less(k(b"", 5, 1), k(b"\x00", 9, 1))


def test_shortest_separator():
    # db/dbformat_test.cc:97-135 InternalKeyShortSeparator
    import ctypes as C
    L = H.oracle()

    def shorten(s, l):
        b = C.create_string_buffer(s, len(s) + 8)
        n = L.orc_shortest_separator(b, C.c_size_t(len(s)), l, C.c_size_t(len(l)))
        return b.raw[:n]

    k = H.ikey
    MAXS, SEEK = H.MAX_SEQ, 0x16
    assert shorten(k(b"foo", 100, 1), k(b"foo", 99, 1)) == k(b"foo", 100, 1)
    assert shorten(k(b"foo", 100, 1), k(b"foo", 101, 1)) == k(b"foo", 100, 1)
    assert shorten(k(b"foo", 100, 1), k(b"bar", 99, 1)) == k(b"foo", 100, 1)
    assert shorten(k(b"foo", 100, 1), k(b"hello", 200, 1)) == k(b"g", MAXS, SEEK)
    assert shorten(k(b"ABC1AAAAA", 100, 1), k(b"ABC2", 200, 1)) == k(b"ABC1B", MAXS, SEEK)  # skip-byte case
    assert shorten(k(b"AAA1", 100, 1), k(b"AAA2", 200, 1)) == k(b"AAA1", 100, 1)
    assert shorten(k(b"foo", 100, 1), k(b"foobar", 200, 1)) == k(b"foo", 100, 1)
    assert shorten(k(b"foobar", 100, 1), k(b"foo", 200, 1)) == k(b"foobar", 100, 1)


def _run_citer(keys, vals, snapshots=(), bottommost=False):
    p = H.Params(snapshots=list(snapshots), bottommost_level=bottommost)
    out, st = H.oracle_citer(p, H.kvstream(zip(keys, vals)))
    return H.parse_kvstream(out), st


def test_citer_zero_out_sequence_at_bottom_level():
    # db/compaction/compaction_iterator_test.cc:723-731
    k = H.ikey
    out, _ = _run_citer([k(b"a", 1, 1), k(b"b", 2, 1)], [b"v1", b"v2"], snapshots=[1], bottommost=True)
    assert out == [(k(b"a", 0, 1), b"v1"), (k(b"b", 2, 1), b"v2")]


def test_citer_remove_deletion_at_bottom_level():
    # db/compaction/compaction_iterator_test.cc:735-745
    k = H.ikey
    out, _ = _run_citer([k(b"a", 1, 0), k(b"b", 3, 0), k(b"b", 1, 1)], [b"", b"", b""], snapshots=[1], bottommost=True)
    assert out == [(k(b"b", 3, 0), b""), (k(b"b", 0, 1), b"")]


def test_citer_hidden_versions_and_tombstones_no_snapshot():
    # SURVEY.md Appendix B rules 1-3 (compaction_iterator.cc:890-946)
    k = H.ikey
    keys = [k(b"a", 9, 1), k(b"a", 5, 1), k(b"b", 8, 0), k(b"b", 2, 1), k(b"c", 7, 1)]
    vals = [b"a9", b"a5", b"", b"b2", b"c7"]
    out, st = _run_citer(keys, vals, bottommost=False)
    assert out == [(keys[0], b"a9"), (keys[2], b""), (keys[4], b"c7")]
    assert (st.num_records_replaced, st.num_expired_deletion_records) == (2, 0)
    out, st = _run_citer(keys, vals, bottommost=True)
    assert out == [(k(b"a", 0, 1), b"a9"), (k(b"c", 0, 1), b"c7")]
    assert (st.num_records_replaced, st.num_expired_deletion_records) == (2, 1)


def test_citer_snapshot_stripes():
    # one visible version per (user key, snapshot stripe): compaction_iterator.cc:619-629,890-911
    k = H.ikey
    keys = [k(b"a", 30, 1), k(b"a", 25, 1), k(b"a", 15, 1), k(b"a", 12, 0), k(b"a", 5, 1)]
    vals = [b"30", b"25", b"15", b"", b"5"]
    out, _ = _run_citer(keys, vals, snapshots=[10, 20], bottommost=False)
    assert [x[0] for x in out] == [keys[0], keys[2], keys[4]]
    out, _ = _run_citer(keys, vals, snapshots=[10, 20], bottommost=True)
    assert [x[0] for x in out] == [keys[0], keys[2], k(b"a", 0, 1)]


def test_citer_rejects_types_outside_rule_set():
    import pytest
    k = H.ikey
    with pytest.raises(RuntimeError):
        _run_citer([k(b"a", 3, 2)], [b"m"])  # kTypeMerge
