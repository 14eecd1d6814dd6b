"""End-to-end known answers of the reference's own CompactionJob tests (db/compaction/compaction_job_test.cc:763-878: Simple,
SimpleDeletion, OutputNothing, SimpleOverwrite, SimpleNonLastLevel -- the ones inside the device rule set: kTypeValue / kTypeDeletion, no
merge operator) run through the whole oracle job: the inputs as BlockBasedTable files, the expected internal keys and values exactly as
the tests spell them (sequence numbers zeroed at the bottommost level, tombstones gone, hidden versions gone)."""
import helpers as H
import sstfmt

VALUE, DELETION = 1, 0


def ik(user_key, seq, t=VALUE):
    return H.ikey(user_key.encode() if isinstance(user_key, str) else user_key, seq, t)


def run(files, bottommost, snapshots=()):
    """files: newest first, each a list of (internal key, value)"""
    inputs = [H.oracle_build_sst(H.Params(), H.kvstream(sorted(f, key=lambda e: (e[0][:-8], -int.from_bytes(e[0][-8:], "little")))))
              for f in files]
    p = H.Params(output_level=1, bottommost_level=bottommost, snapshots=list(snapshots))
    out, _, st = H.oracle_compact(p, inputs)
    return [e for f in out for e in sstfmt.parse_sst(f)["entries"]], st


def test_simple_two_overlapping_files():  # :763-772 with CreateTwoFiles(false) :485-530
    per, matching = 10000, 5000
    files, seq, expected = [], 0, {}
    for i in range(2):
        f = []
        for k in range(per):
            key, val = str(i * matching + k), str(i * per + k).encode()
            seq += 1
            f.append((ik(key, seq), val))
            if i == 1 or k < matching:
                expected[ik(key, 0)] = val
        files.append(f)
    got, st = run(list(reversed(files)), bottommost=True)
    assert got == sorted(expected.items(), key=lambda e: e[0][:-8])
    assert st.num_input_records == 2 * per and st.num_output_records == per + matching


def test_simple_deletion():  # :785-803
    file1 = [(ik("c", 4, DELETION), b""), (ik("c", 3), b"val")]
    file2 = [(ik("b", 2), b"val"), (ik("b", 1), b"val")]
    got, _ = run([file2, file1], bottommost=True)
    assert got == [(ik("b", 0), b"val")]


def test_output_nothing():  # :805-823
    got, st = run([[(ik("a", 2, DELETION), b"")], [(ik("a", 1), b"val")]], bottommost=True)
    assert got == [] and st.num_output_records == 0


def test_simple_overwrite():  # :825-846
    file1 = [(ik("a", 3), b"val2"), (ik("b", 4), b"val3")]
    file2 = [(ik("a", 1), b"val"), (ik("b", 2), b"val")]
    got, _ = run([file2, file1], bottommost=True)
    assert got == [(ik("a", 0), b"val2"), (ik("b", 0), b"val3")]


def test_simple_non_last_level():  # :848-878: level 2 holds older versions, so the output level is not the last one
    file1 = [(ik("a", 5), b"val2"), (ik("b", 6), b"val3")]   # L0
    file2 = [(ik("a", 3), b"val"), (ik("b", 4), b"val")]     # L1
    got, _ = run([file1, file2], bottommost=False)
    assert got == [(ik("a", 5), b"val2"), (ik("b", 6), b"val3")]


# ---- SingleDelete: the reference's CompactionJob tests, extracted by tests/golden/make_compaction_job_kat.py ---------------------------
import json
import os

import pytest

SD_KAT = json.load(open(os.path.join(H.GOLDEN_DIR, "compaction_job_kat.json")))["cases"]


@pytest.mark.parametrize("c", SD_KAT, ids=[c["name"] for c in SD_KAT])
def test_single_delete_known_answers(c):
    """SimpleSingleDelete, SingleDeleteSnapshots, EarliestWriteConflictSnapshot, SingleDeleteZeroSeq, MultiSingleDelete
    (compaction_job_test.cc:1032-1388): SingleDelete/Put pairs cancel only inside one snapshot stripe, a SingleDelete kept for
    write-conflict checking is followed by its Put with the value cleared, leftovers are dropped where the key cannot exist beyond the
    output level.  The tests run a job of the DB itself, so KeyNotExistsBeyondOutputLevel looks at the deeper levels' files."""
    def ent(e):
        return ik(e[0], e[1], e[2]), e[3].encode()
    files = [[ent(e) for e in f["entries"]] for f in c["inputs"]]
    deeper = [(min(e[0] for e in f["entries"]).encode(), max(e[0] for e in f["entries"]).encode()) for f in c["deeper_levels"]]
    inputs = [H.oracle_build_sst(H.Params(), H.kvstream(f)) for f in reversed(files)]  # later AddMockFile = newer level-0 file
    p = H.Params(output_level=1, bottommost_level=not deeper, snapshots=c["snapshots"], deeper_files=deeper,
                 earliest_write_conflict_snapshot=c["earliest_write_conflict_snapshot"] or 0)
    out, _, st = H.oracle_compact(p, inputs)
    got = [e for f in out for e in sstfmt.parse_sst(f)["entries"]]
    assert got == [ent(e) for e in c["expected"]]


def test_remove_single_deletion_at_bottom_level():  # db/compaction/compaction_iterator_test.cc:749-756
    kv = H.kvstream([(ik("a", 1, 7), b""), (ik("b", 2, 7), b"")])
    out, _ = H.oracle_citer(H.Params(bottommost_level=True, snapshots=[1]), kv)
    assert H.parse_kvstream(out) == [(ik("b", 2, 7), b"")]
