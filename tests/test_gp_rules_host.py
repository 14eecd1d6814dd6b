"""The product's grandparent cut rules (toplingdb_b200/csrc/gp_rules.h: the reference's key-driven state machine of
CompactionOutputs::UpdateGrandparentBoundaryInfo / ShouldStopBefore, compaction_outputs.cc:133-354, restated on entry RANKS) compiled
for the host and driven over the block layout of finished jobs the way the encoder's stitch walk drives it
(tests/native/gp_rules_sim.cc).  It must cut exactly where the oracle did (synthetic shapes, tests/gp_cases.py) and where the
unmodified reference did (picker-built jobs).  CPU only: this is the host-logic half of test_gpu_grandparents.py."""
import bisect
import ctypes as C
import os
import subprocess

import pytest

import gp_cases
import helpers as H
import scenarios as S
import sstfmt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("gp") / "gp_rules_sim.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-I" + os.path.join(ROOT, "toplingdb_b200", "csrc"),
                           os.path.join(ROOT, "tests", "native", "gp_rules_sim.cc"), "-o", so])
    L = C.CDLL(so)
    L.gp_rules_sim.restype = C.c_int64
    return L


def _layout(files):
    """output files -> (user keys of all entries, blocks [(first entry, count, flushed bytes before, last of file)], file starts)"""
    ukeys, blocks, starts = [], [], []
    for data in files:
        t = sstfmt.parse_sst(data)
        starts.append(len(ukeys))
        hs = [h for _, h in t["index"]]
        for i, h in enumerate(hs):
            payload, _, _ = sstfmt.read_block(data, h)
            ents = list(sstfmt.block_entries(payload))
            blocks.append((len(ukeys), len(ents), h[0], i == len(hs) - 1))
            ukeys += [k[:-8] for k, _, _ in ents]
    return ukeys, blocks, starts


def _run(sim, p, files):
    ukeys, blocks, starts = _layout(files)
    gps = p.grandparents
    G = len(gps)
    lo = [bisect.bisect_left(ukeys, a) for a, _, _ in gps]
    eq = [bisect.bisect_left(ukeys, b) for _, b, _ in gps]
    hi = [bisect.bisect_right(ukeys, b) for _, b, _ in gps]
    same = [int(i + 1 < G and gps[i + 1][0] == gps[i][1]) for i in range(G)]
    u64 = lambda v: (C.c_uint64 * max(1, len(v)))(*v)
    target = p.target_output_file_size or p.max_output_file_size
    cuts = (C.c_uint64 * (2 * G + 2))()
    n = sim.gp_rules_sim(C.c_uint32(G), u64(lo), u64(eq), u64(hi), u64([s for _, _, s in gps]), (C.c_uint8 * max(1, G))(*same),
                         C.c_uint32(int(p.level_compaction_dynamic_file_size)), C.c_uint64(p.max_compaction_bytes or 25 * target),
                         C.c_uint64(target), C.c_uint64(p.max_output_file_size), C.c_uint64(len(ukeys)), C.c_uint64(len(blocks)),
                         u64([b[0] for b in blocks]), (C.c_uint32 * len(blocks))(*[b[1] for b in blocks]),
                         u64([b[2] for b in blocks]), (C.c_uint8 * len(blocks))(*[int(b[3]) for b in blocks]), cuts, C.c_uint64(2 * G + 2))
    assert n >= 0, f"rules disagree with the layout at block {-n - 1}: {blocks[-n - 1]}"
    # file starts the size rule made: the previous file's last block began at or beyond the maximum
    size_cut = {blocks[i + 1][0] for i, b in enumerate(blocks[:-1]) if b[3] and b[2] >= p.max_output_file_size}
    return list(cuts[:n]), [s for s in starts[1:] if s not in size_cut], len(size_cut)


@pytest.mark.parametrize("name", sorted(gp_cases.CASES))
def test_rank_rules_cut_where_the_oracle_does(sim, name):
    p, inputs = gp_cases.build(**gp_cases.CASES[name])
    files, _, _ = H.oracle_compact(p, inputs)
    got, want, _ = _run(sim, p, files)
    assert got == want
    if name != "grandparents_behind_the_stream":
        assert got


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed,n", [(18, 8000), (5, 40000), (6, 30000), (7, 30000)])
def test_rank_rules_cut_where_the_reference_does(sim, seed, n):
    ops, opts = S.grandparent_cuts(n=n, seed=seed)
    ref = H.run_reference(ops, **opts)
    p = H.params_from_reference(ref)
    got, want, _ = _run(sim, p, ref["outputs"])
    assert got == want and got
