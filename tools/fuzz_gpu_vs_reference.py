"""Random cross-check of the CUDA path against the compiled reference (oracle/_ref) ON THE GPU BOX: scenarios of tests/scenarios.py re-seeded
at random -- the plain ones, SingleDelete scripts, zlib-compressed inputs -- with random table options (block size, restart interval,
format_version 3-5, checksum), a Bloom filter policy at random bits per key, random target file sizes; host or device-resident inputs.
Every job: all output files byte for byte and the job statistics against the reference's own run; every third job is additionally cut
into key ranges by b200c_job_plan_ranges and run as concurrent sub-jobs over the shared inputs, each range against the CPU oracle.
`python tools/fuzz_gpu_vs_reference.py [seconds] [seed]`; recorded runs: profiles/README.md."""
import os
import random
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
import scenarios as S  # noqa: E402


def main():
    import torch  # noqa: F401
    from gpu_harness import job_from_params, run_product
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 4711
    rnd = random.Random(seed0)
    names = [n for n in S.ALL if n != "long_keys"] + ["single_deletes", "single_deletes_nonbottom"] + list(S.ZLIB) * 2
    runs = bad = subruns = rejected = 0
    kinds = {}
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
        if fv == 5 and rnd.random() < 0.4:
            table["bloom_bits"] = rnd.choice([4.5, 10, 12.5, 20])
        if "target_file_size" in opts and rnd.random() < 0.5:
            opts["target_file_size"] = rnd.choice([8 << 10, 20 << 10, 64 << 10, 300 << 10])
        if name in S.ZLIB and rnd.random() < 0.3:
            opts["index_compression"] = 0
        opts = dict(opts, **table)
        ref = H.run_reference(ops, **opts)
        p = H.params_from_reference(ref)
        dev_in = rnd.random() < 0.4
        ok = True
        try:
            files, _, st = run_product(p, ref["inputs"], device_inputs=dev_in)
            ok = files == ref["outputs"] and all(getattr(st, k) == ref["manifest"]["stats"][k] for k in H.STAT_KEYS)
        except Exception as e:  # noqa: BLE001
            if getattr(e, "code", None) == 5:  # B200C_ERR_NOT_SUPPORTED: outside the device rule set (e.g. a SingleDelete key with > 64 versions)
                rejected += 1
                continue
            ok = False
            print("ERROR", name, seed, opts, repr(e)[:300], flush=True)
        runs += 1
        kinds[name] = kinds.get(name, 0) + 1
        if ok and runs % 3 == 0 and ref["inputs"]:
            # the same job as concurrent key-range sub-jobs over shared inputs, every range against the CPU oracle
            parent = job_from_params(p)
            for i, d in enumerate(ref["inputs"]):
                parent.add_input(d, level=0, file_number=i)
            try:
                bounds = parent.plan_ranges(rnd.choice([2, 3, 5]), min_range_bytes=rnd.choice([1 << 10, 16 << 10, 64 << 10]))
                ranges = list(zip([None] + bounds, bounds + [None]))
                subs = [parent.sub_job(range_start=a, range_end=b) for a, b in ranges]
                errs = []

                def go(j):
                    try:
                        j.run()
                    except Exception as e:  # noqa: BLE001
                        errs.append(e)
                ths = [threading.Thread(target=go, args=(j,)) for j in subs]
                for th in ths:
                    th.start()
                for th in ths:
                    th.join()
                if errs:
                    raise errs[0]
                for (a, b), j in zip(ranges, subs):
                    p.range_start, p.range_end = a, b
                    want, _, wst = H.oracle_compact(p, ref["inputs"])
                    got = j.outputs()
                    if got != want or j.stats().num_output_records != wst.num_output_records:
                        ok = False
                    j.close()
                p.range_start = p.range_end = None
                subruns += len(ranges)
            except Exception as e:  # noqa: BLE001
                ok = False
                print("ERROR(sub-jobs)", name, seed, opts, repr(e)[:300], flush=True)
            parent.close()
        if not ok:
            bad += 1
            print("MISMATCH", name, seed, opts, "device_inputs" if dev_in else "host_inputs", flush=True)
    print("seed", seed0, "runs", runs, "sub-job ranges", subruns, "rejected as NOT_SUPPORTED", rejected, "mismatches", bad, "by scenario", kinds)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
