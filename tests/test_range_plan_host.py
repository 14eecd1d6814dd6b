"""toplingdb_b200/csrc/range_plan.h (the host half of b200c_job_plan_ranges / b200c_job_upload_by_ranges: anchors off an index block,
boundaries of equal-byte key ranges, the byte ranges of a file that a key range can touch) compiled for the host.  Checked on
reference-written tables (index_block_restart_interval 1: direct access through the restart array) and on hand-built index blocks with
other restart intervals and both value encodings (the sequential walk), against an independent Python reading of the same blocks."""
import ctypes as C
import os
import random
import struct
import subprocess

import pytest

import helpers as H
import sstfmt

ROOT = H.ROOT


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("rp") / "range_plan_sim.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-I" + os.path.join(ROOT, "toplingdb_b200", "csrc"),
                           os.path.join(ROOT, "tests", "native", "range_plan_sim.cc"), "-o", so])
    return C.CDLL(so)


def _anchors(sim, blk, ndb, fv, user_key, per_file):
    cap = 4096
    keys, lens, by = C.create_string_buffer(16 * cap), (C.c_uint32 * cap)(), (C.c_uint64 * cap)()
    n = sim.plan_sim_anchors(blk, C.c_uint64(len(blk)), C.c_uint64(ndb), fv, int(user_key), per_file, keys, lens, by, cap)
    assert n >= 0
    return [(keys.raw[16 * i:16 * i + lens[i]], by[i]) for i in range(n)]


def _cuts(sim, blk, ndb, fv, user_key, file_len, bounds):
    nb = len(bounds)
    bk = C.create_string_buffer(16 * max(1, nb))
    bl = (C.c_uint32 * max(1, nb))()
    for i, b in enumerate(bounds):
        bk[16 * i:16 * i + len(b)] = b
        bl[i] = len(b)
    cuts, de = (C.c_uint64 * max(1, nb))(), C.c_uint64()
    assert sim.plan_sim_cuts(blk, C.c_uint64(len(blk)), C.c_uint64(ndb), fv, int(user_key), C.c_uint64(file_len), bk, bl, nb, cuts, C.byref(de)) == 0
    return list(cuts)[:nb], de.value


def _expect(index, user_key, per_file, bounds):
    """index: [(separator, (offset, size))] -> anchors, cuts, data_end by the definitions in range_plan.h"""
    seps = [(k if user_key else k[:-8], o + s + 5) for k, (o, s) in index]
    nblk = len(seps)
    step = max(1, nblk // max(per_file, 1))
    anchors, last = [], 0
    for n in range(step, nblk, step):
        k, e = seps[n - 1]
        if len(k) <= 16:
            anchors.append((k, e - last))
            last = e
    data_end = seps[-1][1]
    cuts = []
    for b in bounds:
        hit = [e for k, e in seps if k >= b]
        cuts.append(hit[0] if hit else data_end)
    return anchors, cuts, data_end


@pytest.mark.parametrize("case", [c for c in H.golden_cases() if c != "long_keys"])
def test_anchors_and_cuts_on_reference_written_tables(sim, case):
    g = H.load_golden(case)
    rnd = random.Random(case)
    for data in g["inputs"] + g["outputs"]:
        t = sstfmt.parse_sst(data)
        if len(t["index"]) < 1:
            continue
        blk, _, _ = sstfmt.read_block(data, t["footer"]["index"])
        ndb = sstfmt.prop_u64(t["properties"], "rocksdb.num.data.blocks")
        user_key = sstfmt.prop_u64(t["properties"], "rocksdb.index.key.is.user.key") == 1
        fv = t["footer"]["format_version"]
        ukeys = sorted({ik[:-8] for ik, _ in t["entries"]})
        for per_file in (1, 4, 128):
            bounds = sorted({rnd.choice(ukeys)[:rnd.randint(1, 16)] for _ in range(rnd.randint(0, 5))} | {b"", b"\xff" * 16})
            want_a, want_c, want_e = _expect(t["index"], user_key, per_file, bounds)
            assert _anchors(sim, blk, ndb, fv, user_key, per_file) == want_a
            cuts, de = _cuts(sim, blk, ndb, fv, user_key, len(data), bounds)
            assert (cuts, de) == (want_c, want_e)
            # every block that holds a key of [.., b) ends in front of the cut of b (what the upload order relies on)
            for b, c in zip(bounds, cuts):
                for _, (o, sz) in t["index"]:
                    payload, _, _ = sstfmt.read_block(data, (o, sz))
                    if any(ik[:-8] < b for ik, _, _ in sstfmt.block_entries(payload)):
                        assert o + sz + 5 <= c


def _varint(v):
    out = bytearray()
    while v >= 128:
        out.append((v & 127) | 128)
        v >>= 7
    out.append(v)
    return bytes(out)


def _build_index(seps, handles, restart_interval, value_delta):
    """an index block as BlockBuilder writes it (block_builder.cc:21-32, IndexValue::EncodeTo format.cc:102-118)"""
    buf, restarts, prev_key, prev = bytearray(), [], b"", None
    for i, (k, (o, s)) in enumerate(zip(seps, handles)):
        restart = i % restart_interval == 0
        shared = 0
        if not restart:
            while shared < min(len(k), len(prev_key)) and k[shared] == prev_key[shared]:
                shared += 1
        else:
            restarts.append(len(buf))
        if value_delta and not restart:
            d = s - prev[1]
            val = _varint((d << 1) ^ (d >> 63) if d >= 0 else ((-d) << 1) - 1)
        else:
            val = _varint(o) + _varint(s)
        buf += _varint(shared) + _varint(len(k) - shared)
        if not value_delta:
            buf += _varint(len(val))
        buf += k[shared:] + val
        prev_key, prev = k, (o, s)
    for r in restarts:
        buf += struct.pack("<I", r)
    buf += struct.pack("<I", len(restarts))
    return bytes(buf)


@pytest.mark.parametrize("restart_interval", [1, 2, 5, 16])
@pytest.mark.parametrize("value_delta,user_key", [(True, True), (True, False), (False, True)])
def test_sequential_walk_on_hand_built_index_blocks(sim, restart_interval, value_delta, user_key):
    rnd = random.Random(restart_interval * 7 + value_delta * 3 + user_key)
    nblk = 300
    ukeys = sorted({struct.pack(">QQ", 5, rnd.randrange(1 << 30))[:rnd.randint(12, 16)] for _ in range(nblk * 2)})[:nblk]
    nblk = len(ukeys)
    seps = [k if user_key else k + struct.pack("<Q", (rnd.randrange(1, 1000) << 8) | 1) for k in ukeys]
    handles, off = [], 0
    for _ in range(nblk):
        s = rnd.randint(3000, 4300)
        handles.append((off, s))
        off += s + 5
    blk = _build_index(seps, handles, restart_interval, value_delta)
    index = list(zip(seps, handles))
    fv = 5 if value_delta else 3
    for per_file in (3, 128):
        bounds = sorted({rnd.choice(ukeys) for _ in range(4)} | {ukeys[0][:3]})
        want_a, want_c, want_e = _expect(index, user_key, per_file, bounds)
        assert _anchors(sim, blk, nblk, fv, user_key, per_file) == want_a
        assert _cuts(sim, blk, nblk, fv, user_key, off + 1000, bounds) == (want_c, want_e)


def test_boundaries_split_the_bytes_evenly(sim):
    rnd = random.Random(9)
    anchors = []
    for f in range(6):  # six files, anchors in key order within a file
        ks = sorted(rnd.sample(range(1 << 20), 128))
        anchors += [(struct.pack(">QQ", 1, k), rnd.randint(30000, 40000)) for k in ks]
    total = sum(b for _, b in anchors)
    n = len(anchors)
    keys = b"".join(k.ljust(16, b"\0") for k, _ in anchors)
    lens = (C.c_uint32 * n)(*[len(k) for k, _ in anchors])
    by = (C.c_uint64 * n)(*[b for _, b in anchors])
    for max_ranges, min_bytes in ((4, 0), (8, 0), (64, total // 5), (1, 0)):
        ok, ol = C.create_string_buffer(16 * 64), (C.c_uint32 * 64)()
        nb = sim.plan_sim_boundaries(keys, lens, by, n, C.c_uint64(total), max_ranges, C.c_uint64(min_bytes), ok, ol)
        bounds = [ok.raw[16 * i:16 * i + ol[i]] for i in range(nb)]
        assert bounds == sorted(bounds) and len(set(bounds)) == nb and nb <= max(0, max_ranges - 1)
        if max_ranges == 1:
            assert nb == 0
            continue
        # bytes per range (by the anchors' own accounting) stay within a factor of two of the target
        target = max(total // max_ranges, min_bytes, 1)
        edges = [b""] + bounds + [b"\xff" * 17]
        sizes = [sum(b for k, b in anchors if lo < k <= hi) if lo else sum(b for k, b in anchors if k <= hi) for lo, hi in zip(edges, edges[1:])]
        assert all(s >= target * 0.9 for s in sizes[:-1]), (sizes, target)
        assert max(sizes[:-1]) <= 2 * target + 40000 * 6
