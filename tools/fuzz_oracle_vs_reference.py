"""Random cross-check of the CPU oracle against the compiled reference (oracle/_ref): scenarios of tests/scenarios.py re-seeded at random,
with random table options (block size, restart interval, format_version 3-5, checksum), a Bloom filter policy at random bits per key,
random target file sizes, SingleDelete scenarios and zlib-compressed inputs included.  Every run compares all output files byte for byte and the job statistics.
`python tools/fuzz_oracle_vs_reference.py [seconds] [seed]`; the run recorded in DESIGN.md: 1500 s, seed 777 -> 16 244 jobs, 0 mismatches.
`python tools/fuzz_oracle_vs_reference.py [seconds] [seed] picker`: jobs the DB's own picker builds (DB::CompactRange: grandparents
attached) with 1-8 sub-compactions and a random filter policy, every sub-compaction range checked on its own (files + statistics);
recorded run: 800 s, seed 99 -> 1 105 jobs, 0 mismatches."""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
import scenarios as S  # noqa: E402


def picker_jobs(budget, rnd):
    import sstfmt
    from test_oracle_subcompactions import SUB_STATS, file_parts
    runs = bad = 0
    t0 = time.time()
    while time.time() - t0 < budget:
        seed, n = rnd.randrange(1, 10 ** 6), rnd.choice([6000, 15000, 30000, 50000])
        ops, opts = S.grandparent_cuts(n=n, seed=seed)
        extra = dict(max_subcompactions=rnd.choice([1, 1, 2, 4, 8]))
        if rnd.random() < 0.4:
            extra["bloom_bits"] = rnd.choice([5, 10, 15.5])
        ref = H.run_reference(ops, **dict(opts, **extra))
        ranges = H.subcompaction_ranges(ref)
        props = [sstfmt.parse_sst(o)["properties"] for o in ref["outputs"]]
        k, ok = 0, True
        for start, end, rs in ranges:
            p = H.params_from_reference(ref)
            p.range_start, p.range_end = start, end
            nf = len(H.oracle_compact(p, ref["inputs"])[0])
            p.file_creation_times = [sstfmt.prop_u64(q, "rocksdb.file.creation.time") for q in props[k:k + nf]] or [0]
            files, _, st = H.oracle_compact(p, ref["inputs"])
            want = ref["outputs"][k:k + nf]
            ok = ok and len(want) == len(files) and all(file_parts(a) == file_parts(b) for a, b in zip(files, want))
            ok = ok and (len(ranges) == 1 or all(getattr(st, key) == rs[key] for key in SUB_STATS))
            k += nf
        runs += 1
        if not ok or k != len(ref["outputs"]):
            bad += 1
            print("MISMATCH", seed, n, extra, flush=True)
    print("runs", runs, "mismatches", bad)
    return 1 if bad else 0


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 777)
    if len(sys.argv) > 3 and sys.argv[3] == "picker":
        return picker_jobs(budget, rnd)
    names = [n for n in S.ALL if n != "long_keys"] + ["single_deletes", "single_deletes_nonbottom"] + list(S.ZLIB)
    runs = bad = 0
    t0 = time.time()
    while time.time() - t0 < budget:
        name, seed = rnd.choice(names), rnd.randrange(1000, 100000)
        fn = S.ALL.get(name) or S.ORACLE_ONLY.get(name) or S.ZLIB[name]
        try:
            ops, opts = fn(seed=seed)
        except TypeError:
            continue
        fv = rnd.choice([3, 4, 5, 5, 5])
        table = dict(block_size=rnd.choice([256, 512, 1024, 4096, 8192, 32768]), restart_interval=rnd.choice([1, 2, 4, 16, 64]),
                     format_version=fv, checksum=rnd.choice(["xxh3", "crc32c"]))
        if fv == 5 and rnd.random() < 0.5:
            table["bloom_bits"] = rnd.choice([1, 4.5, 10, 12.5, 20, 33])
        if "target_file_size" in opts and rnd.random() < 0.5:
            opts["target_file_size"] = rnd.choice([8 << 10, 20 << 10, 64 << 10, 300 << 10])
        if name in S.ZLIB and rnd.random() < 0.3:
            opts["index_compression"] = 0
        opts = dict(opts, **table)
        ref = H.run_reference(ops, **opts)
        p = H.params_from_reference(ref)
        files, _, st = H.oracle_compact(p, ref["inputs"])
        runs += 1
        if files != ref["outputs"] or any(getattr(st, k) != ref["manifest"]["stats"][k] for k in H.STAT_KEYS):
            bad += 1
            print("MISMATCH", name, seed, opts, flush=True)
    print("runs", runs, "mismatches", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
