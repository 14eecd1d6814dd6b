"""Seeded jobs with synthetic grandparent files (files one level below the output level), for the grandparent-aware
output-cut rules of CompactionOutputs::ShouldStopBefore (db/compaction/compaction_outputs.cc:231-354).  The CPU oracle
(pinned against the compiled reference in test_oracle_grandparents.py) is the checker; the shapes below are chosen so that
every branch of UpdateGrandparentBoundaryInfo / ShouldStopBefore is taken by at least one case:
  gaps between files, files that share a boundary key (largest_i == smallest_{i+1}), boundary keys that are / are not
  present in the merged stream, keys shorter than the stream's keys, several versions of a boundary key (snapshots),
  first key inside / in front of / behind the grandparents, dynamic file size on and off, a small max_compaction_bytes."""
import random
import struct

import helpers as H


def _key(k, mode):
    if mode == 0:
        return struct.pack(">QQ", 0, k)
    return struct.pack(">QQ", k >> 3, (k * 0x9E3779B97F4A7C15) & ((1 << 64) - 1))


def build(seed, nruns=4, n=12000, vlen=40, ngp=12, dynamic=True, target=256 << 10, max_compaction_bytes=0, snapshots=False,
          share=0.4, short_keys=False, first="inside", gp_size=None):
    rnd = random.Random(1000 + seed)
    universe = n * nruns
    mode = seed % 2
    runs, seq, all_keys = [], 1, set()
    for _ in range(nruns):
        ks = sorted(rnd.sample(range(universe // 2), n))
        run = []
        for k in ks:
            t = 0 if rnd.random() < 0.1 else 1
            run.append((_key(k, mode) + struct.pack("<Q", (seq << 8) | t), b"" if t == 0 else rnd.randbytes(vlen)))
            seq += 1
            all_keys.add(k)
        run.sort(key=lambda e: e[0][:-8])
        runs.append(run)
    inputs = [H.oracle_build_sst(H.Params(), H.kvstream(r)) for r in reversed(runs)]  # newest run first
    present = sorted(all_keys)
    # boundary keys: a mix of keys that exist in the stream and keys that fall between two of them
    lo_k, hi_k = present[0], present[-1]
    if first == "before":
        lo_k = present[len(present) // 10]
    elif first == "behind":
        lo_k, hi_k = hi_k + 10, hi_k + 10 + universe
    cuts = set()
    while len(cuts) < 2 * ngp:
        k = rnd.choice(present) if rnd.random() < 0.6 and first != "behind" else rnd.randrange(lo_k, hi_k + 1)
        if lo_k <= k <= hi_k:
            cuts.add(_key(k, mode))
    cuts = sorted(cuts)  # byte order (the hashed mode does not keep the integer order)
    if first == "inside":
        cuts[0] = min(cuts[0], _key(max(0, present[0] - 1), mode) if rnd.random() < 0.5 else min(_key(k, mode) for k in present[:64]))
    gps = []
    i = 0
    while i + 1 < len(cuts):
        ka, kb = cuts[i], cuts[i + 1]
        if short_keys and rnd.random() < 0.5:
            kb = kb[:rnd.randint(12, 15)]  # a proper prefix: sorts in front of the 16-byte key
            if kb < ka:
                kb = cuts[i + 1]
        size = gp_size if gp_size is not None else rnd.choice([8 << 10, 40 << 10, 200 << 10, 1 << 20])
        gps.append((ka, kb, size))
        if rnd.random() < share:  # the next file starts at this file's largest key
            i += 1
        else:
            i += 2
    gps = [g for j, g in enumerate(gps) if j == 0 or g[0] >= gps[j - 1][1]]
    snaps = sorted(rnd.sample(range(1, seq), 6)) if snapshots else []
    p = H.Params(output_level=1, bottommost_level=False, max_output_file_size=2 * target, target_output_file_size=target,
                 file_creation_times=[7, 8, 9], snapshots=snaps, grandparents=gps, level_compaction_dynamic_file_size=dynamic,
                 max_compaction_bytes=max_compaction_bytes)
    return p, inputs


CASES = {
    "dynamic_mixed": dict(seed=2),
    "dynamic_hashed_keys": dict(seed=1, ngp=20),
    "shared_boundaries_snapshots": dict(seed=3, share=0.9, snapshots=True, nruns=6, n=6000),
    "short_boundary_keys": dict(seed=4, short_keys=True, ngp=16),
    "static_file_size": dict(seed=5, dynamic=False, max_compaction_bytes=1500 << 10, gp_size=600 << 10),
    "small_max_compaction_bytes": dict(seed=6, max_compaction_bytes=700 << 10, gp_size=300 << 10),
    "first_key_before_grandparents": dict(seed=7, first="before"),
    "grandparents_behind_the_stream": dict(seed=8, first="behind", ngp=4),
    "many_small_grandparents": dict(seed=9, ngp=150, gp_size=60 << 10, n=20000),
    "many_large_grandparents": dict(seed=10, ngp=60, gp_size=2 << 20, target=128 << 10),
    "tiny_blocks_worth": dict(seed=11, vlen=4, n=30000, ngp=40, target=64 << 10),
    "large_values": dict(seed=12, vlen=3000, n=2000, ngp=25, target=1 << 20),
}
