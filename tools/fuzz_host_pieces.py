"""Random checks of the product code that also compiles for the host (the same headers / sources the device kernels and libb200c.so use):
  gp     toplingdb_b200/csrc/gp_rules.h     grandparent cut rules on ranks      vs the oracle's file boundaries (tests/gp_cases.py shapes)
  tail   toplingdb_b200/csrc/sst_host.cc    properties / metaindex / footer     vs tails the compiled reference wrote (random table options)
  group  toplingdb_b200/csrc/group_rules.h  per-key CompactionIterator walk     vs the oracle's iterator (random SingleDelete streams)
`python tools/fuzz_host_pieces.py [seconds per piece] [seed]`.  Recorded run: 60 random grandparent configurations,
40 798 tails, 600 SingleDelete streams -- no mismatch."""
import ctypes as C
import os
import random
import struct
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gp_cases  # noqa: E402
import helpers as H  # noqa: E402
import scenarios as S  # noqa: E402
import test_gp_rules_host as TGP  # noqa: E402
import test_group_rules_host as TGR  # noqa: E402
import test_sst_host as TT  # noqa: E402


def build(src):
    so = os.path.join(tempfile.mkdtemp(prefix="b200c_fuzz_"), "x.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-I" + os.path.join(ROOT, "toplingdb_b200", "csrc"),
                           os.path.join(ROOT, "tests", "native", src), "-o", so])
    return C.CDLL(so)


def fuzz_gp(budget, rnd):
    L = build("gp_rules_sim.cc")
    L.gp_rules_sim.restype = C.c_int64
    runs = bad = 0
    t0 = time.time()
    while time.time() - t0 < budget:
        kw = dict(seed=rnd.randrange(100, 10 ** 6), nruns=rnd.choice([2, 3, 5]), n=rnd.choice([3000, 8000]), vlen=rnd.choice([0, 16, 200]),
                  ngp=rnd.choice([2, 7, 30, 90]), dynamic=rnd.random() < 0.7, target=rnd.choice([64 << 10, 256 << 10, 1 << 20]),
                  max_compaction_bytes=rnd.choice([0, 0, 300 << 10, 2 << 20]), snapshots=rnd.random() < 0.4, share=rnd.choice([0, 0.4, 0.95]),
                  short_keys=rnd.random() < 0.3, first=rnd.choice(["inside", "before", "behind"]), gp_size=rnd.choice([None, 5 << 10, 100 << 10, 3 << 20]))
        p, inputs = gp_cases.build(**kw)
        files, _, _ = H.oracle_compact(p, inputs)
        try:
            got, want, _ = TGP._run(L, p, files)
            ok = got == want
        except AssertionError:
            ok = False
        runs += 1
        if not ok:
            bad += 1
            print("gp MISMATCH", kw, flush=True)
    return runs, bad


def fuzz_tail(budget, rnd):
    if not H.have_ref():
        return 0, 0
    L = build("tail_sim.cc")
    L.tail_sim_build.restype = C.c_uint64
    names = [n for n in S.ALL if n != "long_keys"]
    runs = bad = 0
    t0 = time.time()
    while time.time() - t0 < budget:
        name, seed = rnd.choice(names), rnd.randrange(1000, 10 ** 6)
        ops, opts = S.ALL[name](seed=seed)
        fv = rnd.choice([3, 4, 5, 5])
        table = dict(block_size=rnd.choice([512, 4096, 16384]), restart_interval=rnd.choice([1, 4, 16]), format_version=fv,
                     checksum=rnd.choice(["xxh3", "crc32c"]))
        if fv == 5 and rnd.random() < 0.5:
            table["bloom_bits"] = rnd.choice([3, 10, 17.5])
        ref = H.run_reference(ops, **dict(opts, **table))
        for data in ref["outputs"] + ref["inputs"]:
            got, want = TT._rebuild_tail(L, data)
            runs += 1
            if got != want:
                bad += 1
                print("tail MISMATCH", name, seed, table, flush=True)
    return runs, bad


def fuzz_group(budget, rnd):
    L = build("group_rules_sim.cc")
    runs = bad = 0
    t0 = time.time()
    while time.time() - t0 < budget:
        stream, seq = [], 9000
        for k in range(200):
            uk, sd = struct.pack(">QQ", 0, k), rnd.random() < 0.7
            for _ in range(rnd.choice([1, 2, 3, 4, 6, 10])):
                t = rnd.choice([1, 1, 7, 7]) if sd else rnd.choice([1, 0, 0, 1])
                stream.append((uk + struct.pack("<Q", (seq << 8) | t), b"" if t != 1 else rnd.randbytes(rnd.randint(0, 6))))
                seq -= rnd.randint(1, 4)
        snaps = sorted(rnd.sample(range(1, 9000), rnd.choice([0, 1, 2, 5, 20])))
        ewcs = rnd.choice([0, 0, rnd.choice(snaps) if snaps else 0, rnd.randrange(1, 9000)])
        if ewcs and snaps and ewcs < snaps[0]:
            ewcs = snaps[0]  # earliest_write_conflict_snapshot >= earliest_snapshot by construction
        p = H.Params(bottommost_level=rnd.random() < 0.5, snapshots=snaps, earliest_write_conflict_snapshot=ewcs,
                     compaction_filter=rnd.choice(["none", "none", "remove_empty_value"]))
        want_kv, st = H.oracle_citer(p, H.kvstream(stream))
        got, cnt = TGR.walk_stream(L, p, stream)
        runs += 1
        if got != H.parse_kvstream(want_kv) or cnt[0] != st.num_records_replaced or cnt[1] != st.num_expired_deletion_records:
            bad += 1
            print("group MISMATCH", flush=True)
    return runs, bad


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
    rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    total_bad = 0
    for name, fn in (("gp", fuzz_gp), ("tail", fuzz_tail), ("group", fuzz_group)):
        runs, bad = fn(budget, rnd)
        total_bad += bad
        print(f"{name}: runs {runs} mismatches {bad}")
    return 1 if total_bad else 0


if __name__ == "__main__":
    sys.exit(main())
