"""Named compaction scenarios (write scripts for the compiled reference + job options).
Used by tests/golden/make_golden.py (committed fixtures) and by the ref/oracle/GPU parity tests."""
import random
import struct

from helpers import Ops


def key16(i, hi=0):
    return struct.pack(">QQ", hi, i)


def scramble(i, n_bits=40):
    """fixed odd multiplier scramble of a dense range"""
    return (i * 0x9E3779B97F4A7C15) & ((1 << n_bits) - 1)


def _runs(ops, runs, rnd, vlen, del_frac=0.0, snapshot_after=()):
    for r, keys in enumerate(runs):
        for k in keys:
            if del_frac and rnd.random() < del_frac:
                ops.delete(k)
            else:
                ops.put(k, rnd.randbytes(vlen) if not callable(vlen) else rnd.randbytes(vlen(rnd)))
        ops.flush()
        if r in snapshot_after:
            ops.snapshot()


def basic_bottommost(n=500, nruns=4, seed=1, vlen=100):
    rnd = random.Random(seed)
    ops = Ops()
    runs = [sorted(key16(k) for k in rnd.sample(range(n * 4), n)) for _ in range(nruns)]
    _runs(ops, runs, rnd, vlen, del_frac=0.1)
    return ops, dict(target_file_size=32 << 10)


def nonbottom_tombstones(n=400, nruns=3, seed=2):
    rnd = random.Random(seed)
    ops = Ops()
    # base data pushed to L6 and spanning every key the job will see -> job is not bottommost and
    # KeyNotExistsBeyondOutputLevel() is false for all its keys (same as on a dcompact worker)
    base = [key16(0)] + sorted(key16(k) for k in rnd.sample(range(1, n * 4), n)) + [key16(1 << 40)]
    for k in base:
        ops.put(k, rnd.randbytes(20))
    ops.flush()
    ops.compact_all_to(6)
    runs = [sorted(key16(k) for k in rnd.sample(range(1, n * 4), n)) for _ in range(nruns)]
    _runs(ops, runs, rnd, 64, del_frac=0.25)
    return ops, dict(target_file_size=48 << 10)


def snapshots(n=300, nruns=5, seed=3):
    rnd = random.Random(seed)
    ops = Ops()
    runs = [sorted(key16(k) for k in rnd.sample(range(n * 2), n)) for _ in range(nruns)]
    _runs(ops, runs, rnd, 40, del_frac=0.2, snapshot_after=(0, 2, 3))
    return ops, dict(target_file_size=64 << 10)


def snapshots_nonbottom(n=300, nruns=4, seed=4):
    rnd = random.Random(seed)
    ops = Ops()
    for k in [key16(0), key16(1 << 40)]:
        ops.put(k, b"base")
    ops.flush()
    ops.compact_all_to(6)
    runs = [sorted(key16(k) for k in rnd.sample(range(1, n * 2), n)) for _ in range(nruns)]
    _runs(ops, runs, rnd, 40, del_frac=0.3, snapshot_after=(1,))
    return ops, dict(target_file_size=64 << 10)


def varlen_keys(n=400, nruns=3, seed=5, maxlen=16):
    rnd = random.Random(seed)
    ops = Ops()
    pool = sorted({rnd.randbytes(rnd.randint(1, maxlen))[: rnd.randint(1, maxlen)] for _ in range(n * 3)} |
                  {b"a", b"a\x00", b"a\x00\x00", b"ab", b"\xff" * maxlen, b"\x00"})
    runs = [sorted(rnd.sample(pool, n)) for _ in range(nruns)]
    _runs(ops, runs, rnd, lambda r: r.randint(0, 300), del_frac=0.1)
    return ops, dict(target_file_size=40 << 10)


def long_keys(n=200, nruns=3, seed=6):
    return varlen_keys(n, nruns, seed, maxlen=48)


def crc32c_small_blocks(n=600, nruns=3, seed=7):
    rnd = random.Random(seed)
    ops = Ops()
    runs = [sorted(key16(k) for k in rnd.sample(range(n * 3), n)) for _ in range(nruns)]
    _runs(ops, runs, rnd, 24, del_frac=0.05)
    return ops, dict(target_file_size=16 << 10, checksum="crc32c", block_size=1024, restart_interval=4)


def same_user_key_across_blocks(seed=8):
    """many snapshot-separated versions of few keys: adjacent data blocks share a user key, so the index
    keeps full internal keys (index_builder.h:175-180)"""
    rnd = random.Random(seed)
    ops = Ops()
    keys = [key16(k) for k in range(6)]
    for _ in range(60):
        for k in keys:
            ops.put(k, rnd.randbytes(90))
        ops.flush()
        ops.snapshot()
    return ops, dict(target_file_size=20 << 10)


def tiny(seed=9):
    ops = Ops()
    ops.put(key16(7), b"v")
    ops.flush()
    return ops, dict()


def all_deleted(seed=10):
    ops = Ops()
    for k in range(50):
        ops.put(key16(k), b"x" * 10)
    ops.flush()
    for k in range(50):
        ops.delete(key16(k))
    ops.flush()
    return ops, dict()


def cfg2_mini(per_run=2000, nruns=8, seed=11, vlen=32):
    """SURVEY.md §8d cfg2 in miniature: disjoint, fully interleaved scrambled keys, no drops"""
    rnd = random.Random(seed)
    ops = Ops()
    total = per_run * nruns
    allk = sorted(scramble(i) for i in range(total))
    runs = [[key16(k) for k in allk[r::nruns]] for r in range(nruns)]
    _runs(ops, runs, rnd, vlen)
    return ops, dict(target_file_size=256 << 10, output_level=1)


def cfg3_mini(per_run=1500, nruns=16, seed=12, vlen=256):
    """cfg3 in miniature: 30 % of user keys in >= 2 runs, 10 % tombstones"""
    rnd = random.Random(seed)
    ops = Ops()
    universe = list(range(per_run * nruns))
    runs = []
    for r in range(nruns):
        ks = set(rnd.sample(universe, int(per_run * 0.7)))
        while len(ks) < per_run:
            ks.add(rnd.choice(universe[: per_run * 2]))
        runs.append(sorted(key16(scramble(k)) for k in ks))
    _runs(ops, runs, rnd, vlen, del_frac=0.1)
    return ops, dict(target_file_size=512 << 10)


def output_level0(seed=13):
    """output level 0 never cuts files (compaction_outputs.cc:272)"""
    ops, _ = basic_bottommost(n=300, nruns=3, seed=seed)
    return ops, dict(target_file_size=8 << 10, output_level=0)


def _filter_runs(ops, rnd, n, nruns, snapshot_after=()):
    """runs where ~30 % of the values are empty: RemoveEmptyValueCompactionFilter turns the NEWEST version of a user key into a
    tombstone when its value is empty (older empty versions are left to the ordinary rules)"""
    for r in range(nruns):
        for k in sorted(rnd.sample(range(1, n * 2), n)):
            x = rnd.random()
            if x < 0.1:
                ops.delete(key16(k))
            elif x < 0.4:
                ops.put(key16(k), b"")
            else:
                ops.put(key16(k), rnd.randbytes(rnd.randint(1, 60)))
        ops.flush()
        if r in snapshot_after:
            ops.snapshot()


def filter_empty_value(n=400, nruns=4, seed=14):
    """bottommost: converted tombstones are dropped unless a snapshot pins them"""
    rnd = random.Random(seed)
    ops = Ops()
    _filter_runs(ops, rnd, n, nruns, snapshot_after=(1,))
    return ops, dict(target_file_size=24 << 10, filter="remove_empty_value")


def filter_empty_value_nonbottom(n=400, nruns=3, seed=15):
    """not bottommost: converted tombstones are written out as kTypeDeletion entries"""
    rnd = random.Random(seed)
    ops = Ops()
    for k in [key16(0), key16(1 << 40)]:
        ops.put(k, b"base")
    ops.flush()
    ops.compact_all_to(6)
    _filter_runs(ops, rnd, n, nruns)
    return ops, dict(target_file_size=24 << 10, filter="remove_empty_value")


def ttl_filter(n=400, nruns=4, seed=16, nonbottom=False):
    """DBWithTTL (utilities/ttl): every value carries its write time; TtlCompactionFilter turns the newest version of a user key
    into a tombstone when write time + ttl < now.  Runs written at different clock readings, compaction at a later one."""
    rnd = random.Random(seed)
    ops = Ops()
    t0 = 1_700_000_000
    if nonbottom:
        ops.set_time(t0)
        for k in [key16(0), key16(1 << 40)]:
            ops.put(k, b"base")
        ops.flush()
        ops.compact_all_to(6)
    for r in range(nruns):
        ops.set_time(t0 + 400 * r)  # ttl 1000, compaction at t0 + 1500: runs 0 and 1 are stale, 2 and 3 are fresh
        for k in sorted(rnd.sample(range(1, n * 2), n)):
            if rnd.random() < 0.1:
                ops.delete(key16(k))
            else:
                ops.put(key16(k), rnd.randbytes(rnd.randint(0, 50)))
        ops.flush()
        if r == 1:
            ops.snapshot()
    ops.set_time(t0 + 1500)
    return ops, dict(target_file_size=24 << 10, ttl=1000)


def ttl_filter_nonbottom(n=400, nruns=4, seed=17):
    return ttl_filter(n, nruns, seed, nonbottom=True)


def grandparent_cuts(n=8000, seed=18, nruns=3):
    """The job as the DB's own picker builds it (DB::CompactRange): L0 -> L1 with the overlapping L2 files as grandparents, so
    CompactionOutputs::ShouldStopBefore also cuts output files at grandparent boundaries (compaction_outputs.cc:294-351).
    Oracle-only: the device path does not implement these cut rules yet (the executor keeps such jobs on the CPU)."""
    rnd = random.Random(seed)
    ops = Ops()
    for k in sorted(rnd.sample(range(1, n * 4), n)):
        ops.put(key16(k), rnd.randbytes(60))
    ops.flush()
    ops.compact_all_to(2)
    for _ in range(nruns):
        for k in sorted(rnd.sample(range(1, n * 4), n // 2)):
            ops.put(key16(k), rnd.randbytes(60))
        ops.flush()
    return ops, dict(mode="range", target_file_size=512 << 10, setup_file_size=128 << 10)


# scenarios only the CPU oracle is checked on (rules the device path rejects)
ORACLE_ONLY = dict(grandparent_cuts=grandparent_cuts)

ALL = dict(basic_bottommost=basic_bottommost, nonbottom_tombstones=nonbottom_tombstones, snapshots=snapshots,
           snapshots_nonbottom=snapshots_nonbottom, varlen_keys=varlen_keys, long_keys=long_keys,
           crc32c_small_blocks=crc32c_small_blocks, same_user_key_across_blocks=same_user_key_across_blocks,
           tiny=tiny, all_deleted=all_deleted, cfg2_mini=cfg2_mini, cfg3_mini=cfg3_mini, output_level0=output_level0,
           filter_empty_value=filter_empty_value, filter_empty_value_nonbottom=filter_empty_value_nonbottom,
           ttl_filter=ttl_filter, ttl_filter_nonbottom=ttl_filter_nonbottom)


def single_deletes(n=400, nruns=5, seed=19, nonbottom=False, with_snapshots=True):
    """SingleDelete (compaction_iterator.cc:662-887): Put / SingleDelete pairs across runs, repeated pairs on one key, SingleDeletes
    whose Put lives in an older snapshot stripe or below the output level, dangling SingleDeletes.  Delete is never mixed with
    SingleDelete on one key (that breaks the SingleDelete contract and fails the job).  Outside the device rule set: oracle-side only."""
    rnd = random.Random(seed)
    ops = Ops()
    if nonbottom:
        for k in [key16(0), key16(1 << 40)]:
            ops.put(k, b"base")
        ops.flush()
        ops.compact_all_to(6)
    sd_keys = set(rnd.sample(range(1, n * 2), n // 2))  # keys that only ever see Put / SingleDelete
    live = set()
    for r in range(nruns):
        for k in sorted(rnd.sample(range(1, n * 2), n)):
            if k in sd_keys:
                if k in live and rnd.random() < 0.6:
                    ops.single_delete(key16(k))
                    live.discard(k)
                    if rnd.random() < 0.3:  # written again in the same run: Put SD Put
                        ops.put(key16(k), rnd.randbytes(rnd.randint(0, 30)))
                        live.add(k)
                elif k not in live:
                    if rnd.random() < 0.1:
                        ops.single_delete(key16(k))  # nothing to delete: a dangling SingleDelete
                    else:
                        ops.put(key16(k), rnd.randbytes(rnd.randint(0, 30)))
                        live.add(k)
            elif rnd.random() < 0.15:
                ops.delete(key16(k))
            else:
                ops.put(key16(k), rnd.randbytes(rnd.randint(0, 30)))
        ops.flush()
        if with_snapshots and r in (1, 3):
            ops.snapshot()
    return ops, dict(target_file_size=48 << 10)


ORACLE_ONLY["single_deletes"] = single_deletes
ORACLE_ONLY["single_deletes_nonbottom"] = lambda **kw: single_deletes(nonbottom=True, **kw)


_WORDS = [b"compaction", b"level", b"block", b"table", b"restart", b"varint", b"checksum", b"footer", b"index", b"merge", b"snapshot", b"tombstone"]


def _texty(rnd, n):
    out = bytearray()
    while len(out) < n:
        out += rnd.choice(_WORDS) + b" "
    return bytes(out[:n])


def zlib_inputs(n=1500, nruns=4, seed=21, nonbottom=False):
    """Inputs written with kZlibCompression (flush compression), outputs uncompressed: what a job reads when the upper levels are
    compressed and its output level is not.  Values are mostly word soup (blocks compress and are stored as varint32 size + raw
    deflate), some runs of random bytes (blocks that do not shrink by 12.5 % stay uncompressed: block_based_table_builder.cc:1150-1275),
    some empty; the index blocks are compressed too (enable_index_compression).  Registered apart from ALL: the generic per-scenario
    tests rebuild inputs with the uncompressed table builder."""
    rnd = random.Random(seed)
    ops = Ops()
    if nonbottom:
        for k in [key16(0), key16(1 << 40)]:
            ops.put(k, b"base")
        ops.flush()
        ops.compact_all_to(6)
    for r in range(nruns):
        keys = sorted(rnd.sample(range(1, n * 3), n))
        noisy = None
        for i, k in enumerate(keys):
            if i % 200 == 0:
                noisy = rnd.random() < 0.3  # a stretch of incompressible values now and then
            if rnd.random() < 0.1:
                ops.delete(key16(k))
            elif rnd.random() < 0.05:
                ops.put(key16(k), b"")
            elif noisy:
                ops.put(key16(k), rnd.randbytes(rnd.randint(20, 120)))
            else:
                ops.put(key16(k), _texty(rnd, rnd.randint(1, 400)))
        ops.flush()
        if r == 1:
            ops.snapshot()
    return ops, dict(target_file_size=96 << 10, input_compression="zlib")


ZLIB = dict(zlib_inputs=zlib_inputs, zlib_inputs_nonbottom=lambda **kw: zlib_inputs(nonbottom=True, **kw))
