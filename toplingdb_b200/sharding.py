"""Multi-GPU sharding of compaction work: one process per GPU, every rank runs INDEPENDENT sub-compactions.

The reference splits a big job into key-range sub-compactions that share nothing on the data path
(db/compaction/compaction_job.cc:476-640 GenSubcompactionBoundaries, :661-700 one thread per SubcompactionState) and
distributes whole jobs to dcompact workers (db/compaction/compaction_executor.h:153-185).  Both map to the same thing
here: a rank owns whole jobs; nothing but the (smallest, largest) internal keys of what a rank produced ever crosses
ranks.  That exchange is the one collective of the path: a 64-byte-per-rank all-gather (NCCL on the GPU box, gloo in
the CPU tests) after which every rank checks the level's invariant that output files of different sub-compactions do not
overlap (what VersionBuilder's CheckConsistency would refuse at install time, db/version_builder.cc:380-420).

Host logic only: nothing here touches the data path, and it never falls back to computing anything on the CPU."""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

MAX_IKEY = 24  # device path: user key <= 16 B + 8 B trailer
_REC = 2 * (1 + MAX_IKEY) + 14  # smallest + largest (length-prefixed) + n_files(u32) + n_entries(u64) + pad = 64 B
assert _REC == 64


@dataclass
class Boundary:
    """what one rank reports about the outputs of the sub-compaction(s) it ran in this step"""
    smallest: bytes = b""  # smallest internal key of its first output file ("" = produced nothing)
    largest: bytes = b""   # largest internal key of its last output file
    n_files: int = 0
    n_entries: int = 0

    def pack(self) -> bytes:
        if len(self.smallest) > MAX_IKEY or len(self.largest) > MAX_IKEY:
            raise ValueError("internal key longer than the device path supports")
        b = bytes([len(self.smallest)]) + self.smallest.ljust(MAX_IKEY, b"\0")
        b += bytes([len(self.largest)]) + self.largest.ljust(MAX_IKEY, b"\0")
        b += struct.pack("<IQ", self.n_files, self.n_entries) + b"\0\0"
        return b

    @staticmethod
    def unpack(b: bytes) -> "Boundary":
        assert len(b) == _REC
        ls = b[0]
        ll = b[1 + MAX_IKEY]
        nf, ne = struct.unpack_from("<IQ", b, 2 * (1 + MAX_IKEY))
        return Boundary(bytes(b[1:1 + ls]), bytes(b[2 + MAX_IKEY:2 + MAX_IKEY + ll]), nf, ne)


def key_range_base(rank: int, units_per_rank: int, gap: int = 1024) -> int:
    """first key ordinal of the range rank `rank` owns: ranges are disjoint and ordered by rank, like the reference's
    sub-compaction boundaries (compaction_job.cc:476)"""
    return rank * (units_per_rank + gap)


def assign_jobs(job_bytes: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time assignment of independent jobs (by input bytes) to ranks; returns job indices per rank.
    Deterministic: every rank computes the same plan from the same list, so no collective is needed to agree on it."""
    order = sorted(range(len(job_bytes)), key=lambda i: (-job_bytes[i], i))
    load = [0] * world
    plan: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda q: (load[q], q))
        plan[r].append(i)
        load[r] += job_bytes[i]
    for p in plan:
        p.sort()
    return plan


def _user_key(ikey: bytes) -> bytes:
    return ikey[:-8]


def check_disjoint(bounds: Sequence[Boundary]) -> None:
    """ranks are ordered by key range: the largest user key of rank i must sort strictly before the smallest user key
    of the next rank that produced anything (two files of one sorted level never share a user key,
    compaction_outputs.cc:226-253 cuts only between user keys)."""
    prev: Optional[Tuple[int, Boundary]] = None
    for r, b in enumerate(bounds):
        if b.n_files == 0:
            continue
        if _user_key(b.smallest) > _user_key(b.largest):
            raise ValueError(f"rank {r}: smallest key sorts after largest key")
        if prev is not None and not (_user_key(prev[1].largest) < _user_key(b.smallest)):
            raise ValueError(f"sub-compaction outputs of ranks {prev[0]} and {r} overlap")
        prev = (r, b)


def exchange_boundaries(mine: Boundary, device=None, group=None) -> List[Boundary]:
    """all-gather of the 64-byte boundary record (the only collective of the compaction path).  `device` is the CUDA
    device for an NCCL group, None for gloo.  Returns every rank's record, in rank order, after check_disjoint()."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out = [mine]
    else:
        world = dist.get_world_size(group)
        t = torch.frombuffer(bytearray(mine.pack()), dtype=torch.uint8)
        if device is not None:
            t = t.to(device)
        allt = torch.empty(world * _REC, dtype=torch.uint8, device=t.device)
        dist.all_gather_into_tensor(allt, t, group=group)
        raw = bytes(allt.cpu().numpy().tobytes())
        out = [Boundary.unpack(raw[i * _REC:(i + 1) * _REC]) for i in range(world)]
    check_disjoint(out)
    return out


def job_boundary(job) -> Boundary:
    """Boundary of a finished toplingdb_b200.CompactionJob (reads only output metadata)."""
    n = job.output_count()
    if n == 0:
        return Boundary()
    first, last = job.output_meta(0), job.output_meta(n - 1)
    ents = sum(job.output_meta(i).num_entries for i in range(n))
    return Boundary(bytes(first.smallest_ikey[:first.smallest_ikey_len]), bytes(last.largest_ikey[:last.largest_ikey_len]), n, ents)


def subcompaction_ranges(boundaries: Sequence[bytes]) -> List[Tuple[Optional[bytes], Optional[bytes]]]:
    """n sorted boundary user keys -> the n + 1 key ranges [start, end) of one job's sub-compactions, first and last one
    unbounded, exactly how CompactionJob::Prepare turns `boundaries_` into SubcompactionStates (compaction_job.cc:264-281).
    Each range is one `CompactionJob(range_start=..., range_end=...)` over the SAME input files; ranges are independent, so
    `assign_jobs` spreads them over the ranks and `exchange_boundaries` checks the stitched level afterwards."""
    bs = list(boundaries)
    if any(bs[i] >= bs[i + 1] for i in range(len(bs) - 1)):
        raise ValueError("sub-compaction boundaries must be strictly ascending")
    starts: List[Optional[bytes]] = [None] + bs
    ends: List[Optional[bytes]] = bs + [None]
    return list(zip(starts, ends))


def ranges_of_rank(n_ranges: int, world: int, rank: int) -> range:
    """Contiguous block of one job's sub-compaction ranges for `rank`: rank order = key order, which is what `check_disjoint`
    verifies after the boundary exchange and what lets rank r's output files be installed between those of r - 1 and r + 1."""
    per, extra = divmod(n_ranges, world)
    lo = rank * per + min(rank, extra)
    return range(lo, lo + per + (1 if rank < extra else 0))
