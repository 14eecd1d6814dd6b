#!/usr/bin/env python
"""Regenerates tests/golden/<case>/ from the UNMODIFIED reference compiled into oracle/_ref
(oracle/Makefile `ref`).  Run in the build container (needs /root/reference to have built _ref):

    python tests/golden/make_golden.py

Each case directory holds the reference's own input SSTs (made by its FlushJob), its output SSTs (made
by CompactionJob::Run -> ProcessKeyValueCompaction) and the manifest (job parameters + CompactionJobStats).
The fixtures are what pins the CPU oracle — and through it the CUDA path — when oracle/_ref is absent."""
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers as H  # noqa: E402
import scenarios as S  # noqa: E402

SMALL = dict(  # keep committed fixtures small (tens of KB each)
    basic_bottommost=dict(n=300, nruns=3), nonbottom_tombstones=dict(n=200), snapshots=dict(n=150),
    snapshots_nonbottom=dict(n=150), varlen_keys=dict(n=200), long_keys=dict(n=100), crc32c_small_blocks=dict(n=300),
    same_user_key_across_blocks={}, tiny={}, all_deleted={}, cfg2_mini=dict(per_run=250), cfg3_mini=dict(per_run=60),
    output_level0={}, filter_empty_value=dict(n=150), filter_empty_value_nonbottom=dict(n=150),
    ttl_filter=dict(n=150), ttl_filter_nonbottom=dict(n=150), zlib_inputs=dict(n=400, nruns=3), zlib_inputs_nonbottom=dict(n=300, nruns=3))


def main():
    assert H.have_ref(), "oracle/_ref/ref_compact missing: run `make -C oracle ref` where /root/reference exists"
    only = set(sys.argv[1:])  # optional: regenerate just the named cases (the others keep their committed bytes)
    for name, fn in list(S.ALL.items()) + list(S.ZLIB.items()):
        if only and name not in only:
            continue
        ops, opts = fn(**SMALL[name])
        d = os.path.join(HERE, name)
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
        ref = H.run_reference(ops, workdir=None, **opts)
        man = ref["manifest"]
        for kind in ("inputs", "outputs"):
            os.makedirs(os.path.join(d, kind))
            for meta, data in zip(man[kind], ref[kind]):
                with open(os.path.join(d, kind, meta["name"].lstrip("/")), "wb") as f:
                    f.write(data)
        man["stats"].pop("wall_micros"), man["stats"].pop("elapsed_micros"), man["stats"].pop("cpu_micros")
        json.dump(man, open(os.path.join(d, "manifest.json"), "w"), indent=1, sort_keys=True)
        print(f"{name}: {len(ref['inputs'])} inputs ({sum(map(len, ref['inputs']))} B) -> "
              f"{len(ref['outputs'])} outputs ({sum(map(len, ref['outputs']))} B)")


if __name__ == "__main__":
    main()
