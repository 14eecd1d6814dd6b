"""kTypeSingleDeletion on the device (compaction_iterator.cc:635-661, 662-887): tiles are cut at user-key boundaries
when an input holds a SingleDelete, and every key that has one is walked serially with the reference's rules for one key
(csrc/group_rules.h).  Checked against the reference's own CompactionJob known answers, the compiled reference on seeded scenarios
(byte-exact files), the CPU oracle on larger random jobs that span many tiles, and through the executor plugin."""
import json
import os
import random
import struct

import pytest

try:
    import torch  # noqa: F401
except Exception:  # pragma: no cover
    torch = None

import helpers as H
import scenarios as S
import sstfmt

pytestmark = pytest.mark.gpu
SD_KAT = json.load(open(os.path.join(H.GOLDEN_DIR, "compaction_job_kat.json")))["cases"]
VALUE, DELETION, SINGLE_DELETION = 1, 0, 7


def _T():
    import toplingdb_b200 as T
    return T


def ik(user_key, seq, t=VALUE):
    return H.ikey(user_key.encode() if isinstance(user_key, str) else user_key, seq, t)


@pytest.mark.parametrize("c", SD_KAT, ids=[c["name"] for c in SD_KAT])
def test_single_delete_known_answers_of_the_reference(c):
    """compaction_job_test.cc:1032-1388.  The device evaluates KeyNotExistsBeyondOutputLevel as a RunRemote worker does (true exactly at
    the bottommost level, compaction.cc:555-556): cases without deeper levels must give the test's expected records; the others are
    compared with the oracle under the same worker semantics.  A write-conflict snapshot keeps the job off the device."""
    from gpu_harness import run_product
    T = _T()

    def ent(e):
        return ik(e[0], e[1], e[2]), e[3].encode()
    files = [[ent(e) for e in f["entries"]] for f in c["inputs"]]
    inputs = [H.oracle_build_sst(H.Params(), H.kvstream(f)) for f in reversed(files)]
    deeper = c["deeper_levels"]
    p = H.Params(output_level=1, bottommost_level=not deeper, snapshots=c["snapshots"])
    if c["earliest_write_conflict_snapshot"]:
        with pytest.raises(T.B200cError) as ei:
            run_product(p, inputs, earliest_write_conflict_snapshot=c["earliest_write_conflict_snapshot"])
        assert ei.value.code == T.native.ERR_NOT_SUPPORTED
        return
    got_files, _, st = run_product(p, inputs)
    want_files, _, wst = H.oracle_compact(p, inputs)
    assert got_files == want_files
    for k in H.STAT_KEYS:
        assert getattr(st, k) == getattr(wst, k), k
    if not deeper:
        got = [e for f in got_files for e in sstfmt.parse_sst(f)["entries"]]
        assert got == [ent(e) for e in c["expected"]]


@pytest.mark.parametrize("name", ["single_deletes", "single_deletes_nonbottom"])
@pytest.mark.parametrize("seed", [19, 20, 21, 22])
def test_seeded_single_delete_jobs_match_the_reference(name, seed):
    if not os.path.exists(H.REF_BIN):
        pytest.fail("oracle/_ref/ref_compact missing")
    from gpu_harness import run_product
    ops, opts = S.ORACLE_ONLY[name](seed=seed)
    ref = H.run_reference(ops, **opts)
    assert sum(1 for d in ref["inputs"] for ikey, _ in sstfmt.parse_sst(d)["entries"] if ikey[-8] == SINGLE_DELETION) > 100
    p = H.params_from_reference(ref)
    files, metas, st = run_product(p, ref["inputs"])
    assert files == ref["outputs"]
    for k in H.STAT_KEYS:
        assert getattr(st, k) == ref["manifest"]["stats"][k], k
    for m, want in zip(metas, ref["manifest"]["outputs"]):
        assert (m.file_size, m.num_entries, m.num_deletions) == (want["size"], want["num_entries"], want["num_deletions"])


@pytest.mark.parametrize("seed,nruns,nkeys,bottom,nsnap", [(1, 8, 30000, True, 0), (2, 5, 20000, False, 3), (3, 16, 12000, True, 4), (4, 3, 50000, False, 0)])
def test_large_single_delete_jobs_match_the_oracle(seed, nruns, nkeys, bottom, nsnap):
    """many merge tiles, keys with one to six versions (Put / SingleDelete alternating, plus keys with plain Deletes), snapshots in between:
    every tile boundary that would split a key is moved behind it"""
    from gpu_harness import run_product
    rnd = random.Random(seed)
    seq = 1
    hist = []  # (key, seq, type, value) in write order
    live = {}
    for step in range(nkeys * 3):
        k = rnd.randrange(nkeys)
        kb = struct.pack(">QQ", k, (k * 0x9E3779B97F4A7C15) & ((1 << 64) - 1))[: 8 + (k % 9)]  # unique: the first 8 bytes are k
        sd_key = k % 3 != 0  # two thirds of the keys only ever see Put / SingleDelete, the rest Put / Delete
        if sd_key:
            if live.get(k) and rnd.random() < 0.5:
                hist.append((kb, seq, SINGLE_DELETION, b""))
                live[k] = False
            elif not live.get(k):
                hist.append((kb, seq, VALUE, rnd.randbytes(rnd.choice((0, 8, 40)))))
                live[k] = True
            else:
                continue
        else:
            t = DELETION if rnd.random() < 0.3 else VALUE
            hist.append((kb, seq, t, b"" if t == DELETION else rnd.randbytes(16)))
        seq += 1
    # cut the history into runs (oldest first), snapshots at random points
    per = (len(hist) + nruns - 1) // nruns
    runs = []
    for r in range(nruns):
        chunk = hist[r * per:(r + 1) * per]
        newest = {}
        for kb, s, t, v in chunk:  # a flushed memtable holds every version; keep them all (distinct seqs)
            newest.setdefault(kb, []).append((s, t, v))
        run = []
        for kb in sorted(newest):
            for s, t, v in sorted(newest[kb], reverse=True):
                run.append((kb + struct.pack("<Q", (s << 8) | t), v))
        runs.append(run)
    snaps = sorted(rnd.sample(range(1, seq), nsnap)) if nsnap else []
    inputs = [H.oracle_build_sst(H.Params(), H.kvstream(r)) for r in reversed(runs)]  # newest run first
    p = H.Params(bottommost_level=bottom, snapshots=snaps, max_output_file_size=1 << 20, file_creation_times=[3])
    want, _, wst = H.oracle_compact(p, inputs)
    files, _, st = run_product(p, inputs)
    assert [len(f) for f in files] == [len(f) for f in want]
    assert files == want
    for k in H.STAT_KEYS:
        assert getattr(st, k) == getattr(wst, k), k


def test_single_delete_meeting_a_delete_fails_the_job():
    """enforce_single_del_contracts (default true): Status::Corruption (compaction_iterator.cc:779-800)"""
    from gpu_harness import run_product
    T = _T()
    newer = H.oracle_build_sst(H.Params(), H.kvstream([(H.ikey(b"k", 9, SINGLE_DELETION), b"")]))
    older = H.oracle_build_sst(H.Params(), H.kvstream([(H.ikey(b"k", 5, DELETION), b"")]))
    with pytest.raises(T.B200cError) as ei:
        run_product(H.Params(bottommost_level=False), [newer, older])
    assert ei.value.code == T.native.ERR_CORRUPTION


def test_key_with_more_versions_than_the_serial_walk_takes_is_refused():
    from gpu_harness import run_product
    T = _T()
    ents = []
    for s in range(200, 0, -1):
        ents.append((H.ikey(b"hot", s, SINGLE_DELETION if s % 2 == 0 else VALUE), b"" if s % 2 == 0 else b"v"))
    one = H.oracle_build_sst(H.Params(), H.kvstream(ents))
    other = H.oracle_build_sst(H.Params(), H.kvstream([(H.ikey(b"a", 1000, VALUE), b"x")]))
    with pytest.raises(T.B200cError) as ei:
        run_product(H.Params(bottommost_level=True), [other, one])
    assert ei.value.code == T.native.ERR_NOT_SUPPORTED


def test_reference_db_compacts_single_deletes_through_the_b200_executor():
    if not (os.path.exists(H.REF_BIN) and os.path.exists(H.REF_B200_BIN)):
        pytest.fail("oracle/_ref/ref_compact(_b200) missing")
    ops, opts = S.ORACLE_ONLY["single_deletes"](seed=23)
    want = H.run_reference(ops, **opts)
    got = H.run_reference(ops, binary=H.REF_B200_BIN, executor="b200", **opts)
    gm, wm = got["manifest"], want["manifest"]
    assert gm["executor"] == "B200Compact" and gm["remote_compact_read_bytes"] > 0
    assert (gm["scan_count"], gm["scan_digest"]) == (wm["scan_count"], wm["scan_digest"])
    for k in ("num_entries", "num_deletions", "smallestkey", "largestkey"):
        assert [m[k] for m in gm["outputs"]] == [m[k] for m in wm["outputs"]], k
    for k in H.STAT_KEYS:
        assert gm["stats"][k] == wm["stats"][k], k
