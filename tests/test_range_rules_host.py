"""toplingdb_b200/csrc/range_rules.h (which data blocks of an input file a sub-compaction's key range can touch; groundwork, not wired into
the kernels yet) compiled for the host and checked on reference-written tables: for random ranges no block that holds a key in range may
be dropped, and the selection is tight -- at most one block in front of and one behind the blocks that really hold keys in range."""
import ctypes as C
import os
import random
import subprocess

import pytest

import helpers as H
import sstfmt

ROOT = H.ROOT


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("rr") / "range_rules_sim.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-I" + os.path.join(ROOT, "toplingdb_b200", "csrc"),
                           os.path.join(ROOT, "tests", "native", "range_rules_sim.cc"), "-o", so])
    return C.CDLL(so)


def _blocks(data):
    t = sstfmt.parse_sst(data)
    user_sep = sstfmt.prop_u64(t["properties"], "rocksdb.index.key.is.user.key") == 1
    seps, keys = [], []
    for sep, h in t["index"]:
        seps.append(sep if user_sep else sep[:-8])
        payload, _, _ = sstfmt.read_block(data, h)
        keys.append([ik[:-8] for ik, _, _ in sstfmt.block_entries(payload)])
    return seps, keys


@pytest.mark.parametrize("case", ["basic_bottommost", "varlen_keys", "same_user_key_across_blocks", "crc32c_small_blocks", "cfg3_mini", "snapshots"])
def test_block_selection_is_safe_and_tight(sim, case):
    g = H.load_golden(case)
    rnd = random.Random(case)
    for data in g["inputs"] + g["outputs"]:
        seps, keys = _blocks(data)
        if any(len(s) > 16 for s in seps):
            continue
        n = len(seps)
        allk = sorted({k for blk in keys for k in blk})
        buf = b"".join(s.ljust(16, b"\0") for s in seps)
        lens = (C.c_uint32 * n)(*[len(s) for s in seps])
        for _ in range(40):
            a, b = sorted(rnd.choice(allk) if rnd.random() < 0.7 else rnd.randbytes(rnd.randint(0, 16)) for _ in range(2))
            has_s, has_e = rnd.random() < 0.85, rnd.random() < 0.85
            out = (C.c_uint8 * n)()
            sim.range_rules_select(buf, lens, n, int(has_s), a, len(a), int(has_e), b, len(b), out)
            need = [any((not has_s or k >= a) and (not has_e or k < b) for k in blk) for blk in keys]
            for i in range(n):
                assert out[i] or not need[i], (case, i, a, b)  # safe: nothing in range is dropped
            picked = [i for i in range(n) if out[i]]
            needed = [i for i in range(n) if need[i]]
            if needed:
                assert picked[0] >= needed[0] - 1 and picked[-1] <= needed[-1] + 1  # tight: one boundary block per side at most
            else:
                assert len(picked) <= 2
