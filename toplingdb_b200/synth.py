"""Synthetic sorted runs of the BASELINE.json shapes, generated on the device and pre-staged as BlockBasedTable images
by the library's own encoder (b200c_job_encode_columns).  SURVEY.md §8d: user keys are 16-byte big-endian (hi, lo);
run r of a k-way job holds every k-th key of the job's sorted key space (disjoint, fully interleaved runs); values are
seeded pseudo-random bytes; seq = 1 + global index."""
import torch

from .native import CompactionJob

KEY_MULT = 0x1000193  # monotone spreading of the dense index over the high key word


def _u64(t):
    return t.to(torch.int64)


def make_run_columns(n_total, k, r, vlen, key_base, values, overlap=0.0, del_frac=0.0, seed=0, device="cuda"):
    """Columns of run r (of k).  Returns (n, pfx[n,2] int64, tr[n] int64, vref[n] int64, meta[n] int32).

    overlap > 0: that fraction of the run's user keys is replaced by the key of the same rank in run (r+1) % k, so the
    key also occurs in a second run (the lower run index carries the higher sequence number = newer version).
    del_frac: fraction of entries written as kTypeDeletion with an empty value."""
    idx = torch.arange(r, n_total, k, device=device, dtype=torch.int64)
    n = idx.numel()
    kidx = idx.clone()
    g = torch.Generator(device=device)
    g.manual_seed(seed * 1000 + r)
    if overlap > 0 and k > 1:
        sel = torch.rand(n, device=device, generator=g) < overlap
        kidx = torch.where(sel, idx - r + ((r + 1) % k), kidx)  # same rank, neighbouring run's key
        kidx = torch.clamp(kidx, max=n_total - 1)
    hi = (key_base + kidx) * KEY_MULT
    lo = (kidx * 0x9E3779B97F4A7C15 + 0x7F4A7C15) ^ (kidx << 17)  # any function of the index; low word of the key
    pfx = torch.stack([hi, lo], dim=1).contiguous()
    # newer runs (lower r) get higher sequence numbers; unique per (user key, seq)
    seq = 1 + idx + (k - 1 - r) * n_total
    typ = torch.ones(n, device=device, dtype=torch.int64)
    vl = torch.full((n,), vlen, device=device, dtype=torch.int64)
    if del_frac > 0:
        dele = torch.rand(n, device=device, generator=g) < del_frac
        typ = torch.where(dele, torch.zeros_like(typ), typ)
        vl = torch.where(dele, torch.zeros_like(vl), vl)
    tr = (seq << 8) | typ
    vref = values.data_ptr() + idx * vlen
    meta = ((16 << 27) | vl).to(torch.int32)
    if overlap > 0:  # substitution can create equal neighbours inside the run: keep it strictly sorted
        keep = torch.ones(n, dtype=torch.bool, device=device)
        keep[1:] = kidx[1:] != kidx[:-1]
        pfx, tr, vref, meta = pfx[keep].contiguous(), tr[keep].contiguous(), vref[keep].contiguous(), meta[keep].contiguous()
        n = int(keep.sum().item())
    return n, pfx, tr.contiguous(), vref.contiguous(), meta


def stage_runs(n_total, k, vlen, key_base=0, seed=1, overlap=0.0, del_frac=0.0, device_index=0, checksum="xxh3"):
    """Builds the k input SST images of one job in device memory.  Returns (images[list of uint8 CUDA tensors],
    input_kv_bytes, keepalive).  The value arena must outlive the images only until they are encoded."""
    dev = torch.device("cuda", device_index)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    values = torch.randint(0, 256, (max(1, n_total * vlen),), dtype=torch.uint8, device=dev, generator=g)
    images, kv_bytes = [], 0
    for r in range(k):
        n, pfx, tr, vref, meta = make_run_columns(n_total, k, r, vlen, key_base, values, overlap, del_frac, seed, dev)
        job = CompactionJob(device=device_index, output_level=0, output_mem="device", checksum=checksum,
                            file_creation_times=[1700000000], first_file_number=1000 + r)
        job.encode_columns(n, pfx, tr, vref, meta)
        assert job.output_count() == 1
        m = job.output_meta(0)
        kv_bytes += m.raw_key_size + m.raw_value_size
        img = torch.empty(m.file_size, dtype=torch.uint8, device=dev)
        job.output_read_into(0, img)
        images.append(img)
        job.close()
        del pfx, tr, vref, meta
    torch.cuda.synchronize(dev)
    del values
    return images, kv_bytes


from .synth_workloads import BENCH_JOB, WORKLOADS  # noqa: E402,F401


def bench_shape(workload, scale=1.0):
    """(entries per run, entries per job) of a bench workload"""
    w = WORKLOADS[workload]
    n_run = int(w["run_bytes"] * scale) // (24 + w["vlen"])
    return n_run, n_run * w["k"]


def stage_bench_inputs(workload, rank=0, scale=1.0, device_index=0):
    """the k input images bench.py compacts on rank `rank` (device resident).  Returns (images, input_kv_bytes)."""
    from . import sharding
    w = WORKLOADS[workload]
    _, n_total = bench_shape(workload, scale)
    key_base = sharding.key_range_base(rank, n_total)  # disjoint, ordered key ranges per rank = independent sub-compactions
    return stage_runs(n_total, w["k"], w["vlen"], key_base=key_base, seed=2 + rank, overlap=w["overlap"], del_frac=w["del_frac"],
                      device_index=device_index)


def outputs_digest(job):
    """sha256 over (file size, file bytes) of every output of a finished job, in order (host side, untimed)"""
    import hashlib
    h = hashlib.sha256()
    for i in range(job.output_count()):
        m = job.output_meta(i)
        if job.params.output_mem == 1:  # device images
            t = torch.empty(m.file_size, dtype=torch.uint8, device=torch.device("cuda", job.params.device))
            job.output_read_into(i, t)
            b = t.cpu().numpy()
        else:
            b = memoryview(job.output_bytes(i))
        h.update(m.file_size.to_bytes(8, "little"))
        h.update(b)
    return h.hexdigest()


def files_digest(files):
    """the same digest over a list of bytes objects (what the oracle returns)"""
    import hashlib
    h = hashlib.sha256()
    for f in files:
        h.update(len(f).to_bytes(8, "little"))
        h.update(f)
    return h.hexdigest()
