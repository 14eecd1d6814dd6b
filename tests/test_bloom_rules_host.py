"""The product's Bloom-filter arithmetic (toplingdb_b200/csrc/bloom_rules.h: XXPH3 of a user key held in the two big-endian key-column
words, FastLocalBloom sizing / line / probe positions) compiled for the host (tests/native/bloom_rules_sim.cc) and checked against
 (a) the reference's own Hash64 known answers (util/hash_test.cc, committed as tests/golden/hash64_kat.json),
 (b) the oracle's byte-wise XXPH3 on random keys of every length 0..16,
 (c) filter blocks the compiled reference wrote (live, when oracle/_ref is present) and the oracle builds.
CPU only: the host-logic half of tests/test_gpu_bloom.py."""
import ctypes as C
import json
import os
import random
import subprocess

import pytest

import helpers as H
import scenarios as S
import sstfmt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("bloom") / "bloom_rules_sim.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-I" + os.path.join(ROOT, "toplingdb_b200", "csrc"),
                           os.path.join(ROOT, "tests", "native", "bloom_rules_sim.cc"), "-o", so])
    L = C.CDLL(so)
    L.bloom_sim_hash.restype = C.c_uint64
    L.bloom_sim_hash.argtypes = [C.c_char_p, C.c_uint32]
    L.bloom_sim_build.restype = C.c_uint64
    return L


def test_hash_known_answers_of_the_reference(sim):
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "hash64_kat.json")))["vectors"]
    assert len(kat) >= 40
    for v in kat:
        d = bytes.fromhex(v["hex"])
        assert sim.bloom_sim_hash(d, len(d)) == v["hash64"], v


def test_hash_matches_the_oracle_for_every_key_length(sim):
    L = H.oracle()
    L.orc_xxph3_64.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint64)]
    rnd = random.Random(5)
    for n in range(17):
        for _ in range(200):
            d = rnd.randbytes(n) if rnd.random() < 0.8 else bytes(rnd.choice([0, 0xff, 0x80]) for _ in range(n))
            out = C.c_uint64()
            assert L.orc_xxph3_64(d, n, C.byref(out)) == 0
            assert sim.bloom_sim_hash(d, n) == out.value, (n, d.hex())


def _filter_of(sim, sst, millibits):
    t = sstfmt.parse_sst(sst)
    off, size = t["metaindex"]["fullfilter.rocksdb.BuiltinBloomFilter"]
    ukeys = [ik[:-8] for ik, _ in t["entries"]]
    keys = b"".join(k.ljust(16, b"\0") for k in ukeys)
    lens = (C.c_uint32 * len(ukeys))(*[len(k) for k in ukeys])
    out = C.create_string_buffer(size + 64)
    entries = C.c_uint64()
    n = sim.bloom_sim_build(keys, lens, C.c_uint64(len(ukeys)), C.c_uint32(millibits), out, C.byref(entries))
    return sst[off:off + size], out.raw[:n], entries.value, sstfmt.prop_u64(t["properties"], "rocksdb.num.filter_entries")


@pytest.mark.parametrize("millibits", [10000, 6500, 3000, 16000, 24000])
def test_filter_bits_match_the_oracle(sim, millibits):
    rnd = random.Random(millibits)
    ents = []
    for i, k in enumerate(sorted({rnd.randbytes(rnd.randint(0, 16)) for _ in range(5000)})):
        for s in range(rnd.choice([1, 1, 2, 3])):
            ents.append((k + (((1000000 - i * 4 - s) << 8) | 1).to_bytes(8, "little"), b"v"))
    sst = H.oracle_build_sst(H.Params(bloom_millibits_per_key=millibits), H.kvstream(ents))
    want, got, n_added, n_prop = _filter_of(sim, sst, millibits)
    assert got == want and n_added == n_prop


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("case,bits", [("basic_bottommost", 10), ("varlen_keys", 10), ("cfg3_mini", 7), ("snapshots", 20)])
def test_filter_bits_match_the_reference(sim, case, bits):
    ops, opts = S.ALL[case]()
    ref = H.run_reference(ops, bloom_bits=bits, **opts)
    for sst in ref["outputs"] + ref["inputs"]:
        want, got, n_added, n_prop = _filter_of(sim, sst, bits * 1000)
        assert got == want and n_added == n_prop
