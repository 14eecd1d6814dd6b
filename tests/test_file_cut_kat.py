"""The reference's own known-answer tests for the grandparent file-cut rules (db/compaction/compaction_job_test.cc:1755-2150,
CompactionJobDynamicFileSizeTest: CutForMaxCompactionBytes, CutToSkipGrandparentFile, CutToAlignGrandparentBoundary, ...SameKey),
transcribed into tests/golden/file_cut_kat.json.  They run on mock tables (file size = entries x constant), so they are replayed on the
rules alone: the oracle's should_stop_before sequence (orc_file_cut_sim) and the product's rank-based rules (gp_rules.h through
tests/native/gp_rules_sim.cc) must both cut exactly where the reference's expected output files begin, with
level_compaction_dynamic_file_size on and off."""
import bisect
import ctypes as C
import json
import os

import pytest

import helpers as H
from test_gp_rules_host import sim  # noqa: F401  (fixture: the product's rules compiled for the host)

KAT = json.load(open(os.path.join(H.GOLDEN_DIR, "file_cut_kat.json")))["cases"]
IDS = [f"{c['name']}-{mode}" for c in KAT for mode in ("dynamic", "static")]
ARGS = [(c, mode) for c in KAT for mode in ("dynamic", "static")]


def _params(c, mode):
    dyn = mode == "dynamic"
    return H.Params(output_level=1, bottommost_level=False, target_output_file_size=c["target"],
                    max_output_file_size=(2 if dyn else 1) * c["target"], max_compaction_bytes=c["max_compaction_bytes"],
                    level_compaction_dynamic_file_size=dyn, grandparents=[(a.encode(), b.encode(), 10) for a, b in c["grandparents"]])


@pytest.mark.parametrize("c,mode", ARGS, ids=IDS)
def test_oracle_rules_cut_where_the_reference_kat_expects(c, mode):
    files = c[mode]
    stream = [k.encode() for f in files for k in f]
    want = [0] * len(stream)
    pos = 0
    for f in files:
        want[pos] = 1
        pos += len(f)
    L = H.oracle()
    p = _params(c, mode)
    keys = (C.c_char_p * len(stream))(*stream)
    lens = (C.c_uint32 * len(stream))(*[len(k) for k in stream])
    cut = (C.c_uint8 * len(stream))()
    L.orc_file_cut_sim.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    cp = p.c()
    assert L.orc_file_cut_sim(C.byref(cp), len(stream), keys, lens, c["kv_size"], cut) == 0
    assert list(cut) == want


@pytest.mark.parametrize("c,mode", ARGS, ids=IDS)
def test_product_rank_rules_agree_with_the_reference_kat(sim, c, mode):
    files = c[mode]
    stream = [k.encode() for f in files for k in f]
    gps = [(a.encode(), b.encode(), 10) for a, b in c["grandparents"]]
    G = len(gps)
    lo = [bisect.bisect_left(stream, a) for a, _, _ in gps]
    eq = [bisect.bisect_left(stream, b) for _, b, _ in gps]
    hi = [bisect.bisect_right(stream, b) for _, b, _ in gps]
    same = [int(i + 1 < G and gps[i + 1][0] == gps[i][1]) for i in range(G)]
    # one "block" per entry; the flushed size an entry's successor sees = entries of the file so far x kv_size (mock table model)
    first, foff, last = [], [], []
    pos = 0
    for f in files:
        for j in range(len(f)):
            first.append(pos + j)
            foff.append((j + 1) * c["kv_size"])
            last.append(int(j == len(f) - 1))
        pos += len(f)
    u64 = lambda v: (C.c_uint64 * max(1, len(v)))(*v)
    dyn = mode == "dynamic"
    cuts = (C.c_uint64 * (2 * G + 2))()
    n = sim.gp_rules_sim(C.c_uint32(G), u64(lo), u64(eq), u64(hi), u64([10] * G), (C.c_uint8 * max(1, G))(*same), C.c_uint32(int(dyn)),
                         C.c_uint64(c["max_compaction_bytes"]), C.c_uint64(c["target"]), C.c_uint64((2 if dyn else 1) * c["target"]),
                         C.c_uint64(len(stream)), C.c_uint64(len(stream)), u64(first), (C.c_uint32 * len(stream))(*([1] * len(stream))),
                         u64(foff), (C.c_uint8 * len(stream))(*last), cuts, C.c_uint64(2 * G + 2))
    assert n >= 0, f"rules disagree with the expected files at entry {-n - 1}"
    starts = []
    pos = 0
    for f in files[:-1]:
        pos += len(f)
        starts.append(pos)
    size_cuts = [s for s, f in zip(starts, files) if len(f) * c["kv_size"] >= (2 if dyn else 1) * c["target"]]
    assert list(cuts[:n]) == [s for s in starts if s not in size_cuts]
