"""toplingdb_b200/csrc/group_rules.h: the CompactionIterator rules for one user key as a serial walk over its versions (the planned device
path for keys that carry a SingleDelete; not wired into the kernels yet).  Compiled for the host (tests/native/group_rules_sim.cc) and
run over whole merged input streams, key by key, against the oracle's iterator -- on every golden scenario (Value / Deletion only, with
snapshots, compaction filters, bottommost or not), on seeded SingleDelete streams, and on the reference's SingleDelete known answers."""
import ctypes as C
import itertools
import json
import os
import random
import struct
import subprocess

import pytest

import helpers as H
import sstfmt

ROOT = H.ROOT
MAXSEQ = (1 << 56) - 1


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("gr") / "group_rules_sim.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-I" + os.path.join(ROOT, "toplingdb_b200", "csrc"),
                           os.path.join(ROOT, "tests", "native", "group_rules_sim.cc"), "-o", so])
    return C.CDLL(so)


def _merged(inputs):
    ents = []
    for r, data in enumerate(inputs):
        for ik, v in sstfmt.parse_sst(data)["entries"]:
            ents.append((ik[:-8], -struct.unpack("<Q", ik[-8:])[0], r, ik, v))
    ents.sort(key=lambda e: (e[0], e[1], e[2]))
    return [(e[3], e[4]) for e in ents]


def _filter_removes(p, value):
    if p.compaction_filter == "remove_empty_value":
        return len(value) == 0
    if p.compaction_filter == "ttl" and p.ttl > 0 and len(value) >= 4:
        return struct.unpack("<I", value[-4:])[0] + p.ttl < p.now
    return False


def walk_stream(sim, p, stream, key_not_exists=None):
    """-> (output entries [(internal key, value)], counters) the way CompactionIterator would emit them"""
    snaps = (C.c_uint64 * max(1, len(p.snapshots)))(*p.snapshots)
    counters = (C.c_uint32 * 4)()
    out, job_has_output = [], False
    walk_stream.examined_key_bytes = 0  # key bytes of the versions the iterator examined (total_input_raw_key_bytes)
    for uk, grp in itertools.groupby(stream, key=lambda e: e[0][:-8]):
        grp = list(grp)
        # identical internal keys (same user key and sequence number in two files) are one version to the iterator's rules: the second
        # is hidden by the first; keep them as separate versions, the hidden rule handles it
        n = len(grp)
        seqs = (C.c_uint64 * n)(*[struct.unpack("<Q", ik[-8:])[0] >> 8 for ik, _ in grp])
        types = (C.c_uint8 * n)(*[ik[-8] for ik, _ in grp])
        verd = (C.c_uint8 * (4 * n))()
        kne = int(p.bottommost_level) if key_not_exists is None else int(key_not_exists(uk))
        rc = sim.group_rules_walk(seqs, types, n, snaps, len(p.snapshots), int(p.bottommost_level),
                                  C.c_uint64(p.earliest_write_conflict_snapshot or MAXSEQ), kne,
                                  int(grp[0][0][-8] == 1 and _filter_removes(p, grp[0][1])), int(not job_has_output), verd, counters)
        assert rc == 0
        for i, (ik, v) in enumerate(grp):
            keep, otype, clear, zero = verd[4 * i:4 * i + 4]
            if not clear & 2:  # kGrSkipped: stepped over inside another version's branch
                walk_stream.examined_key_bytes += len(ik)
            if keep:
                seq = 0 if zero else struct.unpack("<Q", ik[-8:])[0] >> 8
                out.append((uk + struct.pack("<Q", (seq << 8) | otype), b"" if clear & 1 else v))
                job_has_output = True
    return out, list(counters)


@pytest.mark.parametrize("case", [c for c in H.golden_cases()])
def test_group_walk_reproduces_the_iterator_on_every_fixture(sim, case):
    g = H.load_golden(case)
    p = H.params_from_reference(g)
    stream = _merged(g["inputs"])
    want_kv, st = H.oracle_citer(p, H.kvstream(stream))
    got, cnt = walk_stream(sim, p, stream)
    assert got == H.parse_kvstream(want_kv)
    assert cnt[0] == st.num_records_replaced and cnt[1] == st.num_expired_deletion_records and cnt[3] == st.num_record_drop_user


@pytest.mark.parametrize("seed", range(12))
def test_group_walk_with_single_deletes_matches_the_oracle(sim, seed):
    rnd = random.Random(seed)
    stream, seq = [], 5000
    for k in range(600):
        uk = struct.pack(">QQ", 0, k)
        sd_key = rnd.random() < 0.6
        for _ in range(rnd.choice([1, 1, 2, 3, 5, 8])):
            t = rnd.choice([1, 1, 7]) if sd_key else rnd.choice([1, 1, 0])
            stream.append((uk + struct.pack("<Q", (seq << 8) | t), b"" if t != 1 else rnd.randbytes(rnd.randint(0, 12))))
            seq -= rnd.randint(1, 3)
    snaps = sorted(rnd.sample(range(1, 5000), rnd.choice([0, 1, 3, 8])))
    ewcs = rnd.choice([0, 0, rnd.choice(snaps) if snaps else 0])
    p = H.Params(bottommost_level=bool(seed % 2), snapshots=snaps, earliest_write_conflict_snapshot=ewcs)
    want_kv, st = H.oracle_citer(p, H.kvstream(stream))
    got, cnt = walk_stream(sim, p, stream)
    assert got == H.parse_kvstream(want_kv)
    assert cnt[0] == st.num_records_replaced and cnt[1] == st.num_expired_deletion_records
    assert walk_stream.examined_key_bytes == st.total_input_raw_key_bytes


KAT = json.load(open(os.path.join(H.GOLDEN_DIR, "compaction_job_kat.json")))["cases"]


@pytest.mark.parametrize("c", KAT, ids=[c["name"] for c in KAT])
def test_group_walk_on_the_reference_single_delete_known_answers(sim, c):
    def ent(e):
        return H.ikey(e[0].encode(), e[1], e[2]), e[3].encode()
    stream = sorted([ent(e) for f in c["inputs"] for e in f["entries"]], key=lambda e: (e[0][:-8], -int.from_bytes(e[0][-8:], "little")))
    deeper = [(min(e[0] for e in f["entries"]).encode(), max(e[0] for e in f["entries"]).encode()) for f in c["deeper_levels"]]
    p = H.Params(bottommost_level=not deeper, snapshots=c["snapshots"], earliest_write_conflict_snapshot=c["earliest_write_conflict_snapshot"] or 0)
    got, _ = walk_stream(sim, p, stream, key_not_exists=lambda uk: not any(a <= uk <= b for a, b in deeper))
    assert got == [ent(e) for e in c["expected"]]
