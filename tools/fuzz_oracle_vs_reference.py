"""Random cross-check of the CPU oracle against the compiled reference (oracle/_ref): scenarios of tests/scenarios.py re-seeded at random,
with random table options (block size, restart interval, format_version 3-5, checksum), a Bloom filter policy at random bits per key,
random target file sizes, SingleDelete scenarios included.  Every run compares all output files byte for byte and the job statistics.
`python tools/fuzz_oracle_vs_reference.py [seconds] [seed]`; the run recorded in DESIGN.md: 1500 s, seed 777 -> 16 244 jobs, 0 mismatches."""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
import scenarios as S  # noqa: E402


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 777)
    names = [n for n in S.ALL if n != "long_keys"] + ["single_deletes", "single_deletes_nonbottom"]
    runs = bad = 0
    t0 = time.time()
    while time.time() - t0 < budget:
        name, seed = rnd.choice(names), rnd.randrange(1000, 100000)
        fn = S.ALL.get(name) or S.ORACLE_ONLY[name]
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
