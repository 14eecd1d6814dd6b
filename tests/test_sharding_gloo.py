"""N>1 host logic on CPU: world_size-2 gloo process group, the boundary all-gather of toplingdb_b200/sharding.py and the
job-to-rank planner.  (The data path itself has no collective: ranks run independent sub-compactions,
db/compaction/compaction_job.cc:476-700.)"""
import os
import socket
import struct
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _ikey(ordinal, seq=1, t=1):
    return struct.pack(">QQ", 0, ordinal) + struct.pack("<Q", (seq << 8) | t)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, q):
    import torch.distributed as dist
    from toplingdb_b200 import sharding as sh
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 1000
        base = sh.key_range_base(rank, n)
        if mode == "ok":
            mine = sh.Boundary(_ikey(base), _ikey(base + n - 1), 3, n)
        elif mode == "overlap":  # rank 1 starts inside rank 0's range
            lo = base if rank == 0 else 500
            mine = sh.Boundary(_ikey(lo), _ikey(lo + n - 1), 3, n)
        elif mode == "same_user_key":  # rank 1's first user key == rank 0's last user key (different seq): still an overlap
            mine = sh.Boundary(_ikey(0, 9), _ikey(999, 9), 1, n) if rank == 0 else sh.Boundary(_ikey(999, 5), _ikey(3000), 1, n)
        elif mode == "empty_rank":  # rank 0 produced nothing (everything deleted)
            mine = sh.Boundary() if rank == 0 else sh.Boundary(_ikey(base), _ikey(base + 5), 1, 6)
        try:
            got = sh.exchange_boundaries(mine)
            q.put((rank, "ok", [(b.smallest, b.largest, b.n_files, b.n_entries) for b in got]))
        except ValueError as e:
            q.put((rank, "err", str(e)))
    finally:
        dist.destroy_process_group()


def _split_worker(rank, world, port, q):
    """one job, its sub-compaction ranges spread over the ranks; the CPU oracle stands in for the device (the host logic around it is
    what runs on every rank of the real thing)"""
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    import sstfmt
    from toplingdb_b200 import sharding as sh
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = H.load_golden("cfg3_mini")
        p = H.params_from_reference(g)
        keys = sorted({ik[:-8] for d in g["inputs"] for ik, _ in sstfmt.parse_sst(d)["entries"]})
        bounds = [keys[len(keys) * i // 5] for i in range(1, 5)]
        ranges = sh.subcompaction_ranges(bounds)
        entries, nfiles, first, last = [], 0, b"", b""
        for i in sh.ranges_of_rank(len(ranges), world, rank):
            p.range_start, p.range_end = ranges[i]
            files, metas, st = H.oracle_compact(p, g["inputs"])
            for f in files:
                es = sstfmt.parse_sst(f)["entries"]
                entries += es
            nfiles += len(files)
        if entries:
            first, last = entries[0][0], entries[-1][0]
        table = sh.exchange_boundaries(sh.Boundary(first, last, nfiles, len(entries)))
        q.put((rank, "ok", ([(b.smallest, b.largest, b.n_files, b.n_entries) for b in table], entries)))
    except ValueError as e:
        q.put((rank, "err", str(e)))
    finally:
        dist.destroy_process_group()


def _run(mode, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    return res


@pytest.mark.timeout(300)
def test_boundary_all_gather_world2():
    res = _run("ok")
    assert [r[1] for r in res] == ["ok", "ok"]
    assert res[0][2] == res[1][2]  # every rank sees the same table, in rank order
    table = res[0][2]
    assert table[0][0] == _ikey(0) and table[1][0] == _ikey(2024) and table[0][2:] == (3, 1000)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("mode", ["overlap", "same_user_key"])
def test_overlapping_subcompactions_are_refused_on_every_rank(mode):
    res = _run(mode)
    assert [r[1] for r in res] == ["err", "err"]
    assert "overlap" in res[0][2]


@pytest.mark.timeout(300)
def test_rank_with_no_output_is_skipped():
    res = _run("empty_rank")
    assert [r[1] for r in res] == ["ok", "ok"]
    assert res[0][2][0] == (b"", b"", 0, 0)


@pytest.mark.timeout(300)
def test_one_job_split_into_key_ranges_over_two_ranks():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    import sstfmt
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_split_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=240) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == ["ok", "ok"]
    assert res[0][2][0] == res[1][2][0]  # same boundary table on both ranks
    g = H.load_golden("cfg3_mini")
    whole = [e for f in g["outputs"] for e in sstfmt.parse_sst(f)["entries"]]
    assert res[0][2][1] + res[1][2][1] == whole  # rank order = key order; nothing lost, nothing twice
    assert sum(b[3] for b in res[0][2][0]) == len(whole)


def test_range_planning_helpers():
    from toplingdb_b200 import sharding as sh
    assert sh.subcompaction_ranges([]) == [(None, None)]
    assert sh.subcompaction_ranges([b"b", b"d"]) == [(None, b"b"), (b"b", b"d"), (b"d", None)]
    with pytest.raises(ValueError):
        sh.subcompaction_ranges([b"d", b"b"])
    for n, w in [(5, 2), (8, 8), (3, 4), (0, 2), (7, 3)]:
        got = [list(sh.ranges_of_rank(n, w, r)) for r in range(w)]
        assert [i for g in got for i in g] == list(range(n))
        assert max(len(g) for g in got) - min(len(g) for g in got) <= 1


def test_boundary_record_roundtrip_and_limits():
    from toplingdb_b200 import sharding as sh
    b = sh.Boundary(b"abc" + b"\1" * 8, b"zzzzzzzzzzzzzzzz" + b"\2" * 8, 7, 1 << 40)
    assert len(b.pack()) == 64 and sh.Boundary.unpack(b.pack()) == b
    assert sh.Boundary.unpack(sh.Boundary().pack()) == sh.Boundary()
    with pytest.raises(ValueError):
        sh.Boundary(b"x" * 25, b"y").pack()
    with pytest.raises(ValueError):
        sh.check_disjoint([sh.Boundary(_ikey(5), _ikey(4), 1, 2)])


def test_job_planner_is_deterministic_and_balanced():
    from toplingdb_b200 import sharding as sh
    sizes = [900, 100, 400, 400, 300, 300, 50, 850]
    plan = sh.assign_jobs(sizes, 4)
    assert sorted(i for p in plan for i in p) == list(range(len(sizes)))
    loads = [sum(sizes[i] for i in p) for p in plan]
    assert max(loads) <= 1.2 * sum(sizes) / 4 + max(sizes) * 0.0 + 100
    assert plan == sh.assign_jobs(list(sizes), 4)
    assert sh.assign_jobs([], 2) == [[], []]
    assert sh.assign_jobs([5], 3) == [[0], [], []]
