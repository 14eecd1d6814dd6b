"""Test helpers: ctypes view of the CPU oracle (oracle/liboracle.so), the ops-script writer and the
driver for the compiled reference (oracle/_ref/ref_compact).  Test infrastructure only."""
import ctypes as C
import json
import os
import shutil
import struct
import subprocess
import tempfile

import sstfmt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_compact")
REF_B200_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_compact_b200")  # same driver + the product's CompactionExecutor plugin
MAX_SEQ = (1 << 56) - 1
CKSUM = {"none": 0, "crc32c": 1, "xxh3": 4}
FILTERS = {"none": 0, "remove_empty_value": 1, "ttl": 2}


class OrcGrandparent(C.Structure):
    _fields_ = [("smallest", C.c_char_p), ("smallest_len", C.c_uint32), ("largest", C.c_char_p), ("largest_len", C.c_uint32),
                ("file_size", C.c_uint64)]


class OrcParams(C.Structure):
    _fields_ = [
        ("output_level", C.c_int32), ("bottommost_level", C.c_int32), ("max_output_file_size", C.c_uint64),
        ("block_size", C.c_uint32), ("block_size_deviation", C.c_uint32), ("block_restart_interval", C.c_uint32),
        ("index_block_restart_interval", C.c_uint32), ("format_version", C.c_uint32), ("checksum_type", C.c_uint32),
        ("snapshots", C.POINTER(C.c_uint64)), ("num_snapshots", C.c_uint32), ("column_family_id", C.c_uint32),
        ("column_family_name", C.c_char_p), ("db_id", C.c_char_p), ("db_session_id", C.c_char_p),
        ("db_host_id", C.c_char_p), ("creation_time", C.c_uint64), ("oldest_key_time", C.c_uint64),
        ("file_creation_times", C.POINTER(C.c_uint64)), ("num_file_creation_times", C.c_uint32),
        ("first_file_number", C.c_uint64), ("compaction_filter", C.c_uint32), ("ttl", C.c_int32), ("now", C.c_int64),
        ("grandparents", C.POINTER(OrcGrandparent)), ("num_grandparents", C.c_uint32),
        ("level_compaction_dynamic_file_size", C.c_uint32), ("max_compaction_bytes", C.c_uint64),
        ("target_output_file_size", C.c_uint64),
        ("range_start", C.c_char_p), ("range_start_len", C.c_uint32), ("has_range_start", C.c_uint32),
        ("range_end", C.c_char_p), ("range_end_len", C.c_uint32), ("has_range_end", C.c_uint32),
        ("bloom_millibits_per_key", C.c_uint32), ("earliest_write_conflict_snapshot", C.c_uint64),
        ("key_not_exists_mode", C.c_uint32), ("deeper_files", C.POINTER(OrcGrandparent)), ("num_deeper_files", C.c_uint32),
    ]


class OrcFileMeta(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("file_number", "file_size", "smallest_seqno", "largest_seqno", "num_entries",
                                          "num_deletions", "raw_key_size", "raw_value_size", "num_data_blocks")] + [
        ("smallest_len", C.c_uint32), ("largest_len", C.c_uint32), ("smallest", C.c_uint8 * 256),
        ("largest", C.c_uint8 * 256)]


class OrcStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("num_input_records", "num_output_records", "num_input_deletion_records",
                                          "num_records_replaced", "num_expired_deletion_records",
                                          "total_input_raw_key_bytes", "total_input_raw_value_bytes",
                                          "num_optimized_del_drop_obsolete", "num_record_drop_user")]


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
        L = C.CDLL(ORACLE_SO)
        L.orc_last_error.restype = C.c_char_p
        L.orc_xxh3_64.restype = C.c_uint64
        L.orc_xxh3_64.argtypes = [C.c_char_p, C.c_size_t]
        for f in ("orc_crc32c_value",):
            getattr(L, f).restype = C.c_uint32
            getattr(L, f).argtypes = [C.c_char_p, C.c_size_t]
        L.orc_crc32c_mask.restype = C.c_uint32
        L.orc_crc32c_mask.argtypes = [C.c_uint32]
        L.orc_block_checksum.restype = C.c_uint32
        L.orc_block_checksum.argtypes = [C.c_uint32, C.c_char_p, C.c_size_t, C.c_uint8]
        L.orc_checksum.restype = C.c_uint32
        L.orc_checksum.argtypes = [C.c_uint32, C.c_char_p, C.c_size_t]
        L.orc_result_file.restype = C.POINTER(C.c_uint8)
        L.orc_shortest_separator.restype = C.c_size_t
        _oracle = L
    return _oracle


class Params:
    """Python-side job description; .c() gives the ctypes struct (keeps referenced arrays alive)."""

    def __init__(self, **kw):
        self.output_level = 1
        self.bottommost_level = True
        self.max_output_file_size = 64 << 20
        self.block_size = 4096
        self.block_size_deviation = 10
        self.block_restart_interval = 16
        self.index_block_restart_interval = 1
        self.format_version = 5
        self.checksum = "xxh3"
        self.snapshots = []
        self.column_family_id = 0
        self.column_family_name = "default"
        self.db_id = "b200c-test-db"
        self.db_session_id = "B200CSESSION00000000"
        self.db_host_id = "b200"
        self.creation_time = 1700000000
        self.oldest_key_time = 0
        self.file_creation_times = [1700000001]
        self.first_file_number = 100
        self.compaction_filter = "none"  # or "remove_empty_value" / "ttl" (with ttl seconds and the clock reading `now`)
        self.ttl = 0
        self.now = 0
        self.grandparents = []  # [(smallest user key, largest user key, file size)]: files one level below the output
        self.level_compaction_dynamic_file_size = True
        self.max_compaction_bytes = 0  # 0: 25 x target
        self.target_output_file_size = 0  # 0: max_output_file_size
        self.range_start = None  # sub-compaction key range: start <= user key < end (None: unbounded)
        self.range_end = None
        self.bloom_millibits_per_key = 0  # NewBloomFilterPolicy(bits) * 1000; 0 = no filter block
        self.earliest_write_conflict_snapshot = 0  # 0 = kMaxSequenceNumber
        self.deeper_files = None  # [(smallest user key, largest user key)] of the levels below the output level: the DB's own (non-worker)
        #                           KeyNotExistsBeyondOutputLevel; None = worker semantics (true only at the bottommost level)
        for k, v in kw.items():
            assert hasattr(self, k), k
            setattr(self, k, v)

    def c(self):
        p = OrcParams()
        p.output_level = self.output_level
        p.bottommost_level = int(self.bottommost_level)
        p.max_output_file_size = self.max_output_file_size
        p.block_size = self.block_size
        p.block_size_deviation = self.block_size_deviation
        p.block_restart_interval = self.block_restart_interval
        p.index_block_restart_interval = self.index_block_restart_interval
        p.format_version = self.format_version
        p.checksum_type = CKSUM[self.checksum]
        self._snaps = (C.c_uint64 * max(1, len(self.snapshots)))(*self.snapshots)
        p.snapshots = C.cast(self._snaps, C.POINTER(C.c_uint64))
        p.num_snapshots = len(self.snapshots)
        p.column_family_id = self.column_family_id
        p.column_family_name = self.column_family_name.encode()
        p.db_id = self.db_id.encode()
        p.db_session_id = self.db_session_id.encode()
        p.db_host_id = self.db_host_id.encode()
        p.creation_time = self.creation_time
        p.oldest_key_time = self.oldest_key_time
        self._fct = (C.c_uint64 * max(1, len(self.file_creation_times)))(*self.file_creation_times)
        p.file_creation_times = C.cast(self._fct, C.POINTER(C.c_uint64))
        p.num_file_creation_times = len(self.file_creation_times)
        p.first_file_number = self.first_file_number
        p.compaction_filter = FILTERS[self.compaction_filter]
        p.ttl = self.ttl
        p.now = self.now
        self._gp = (OrcGrandparent * max(1, len(self.grandparents)))()
        for i, (a, b, sz) in enumerate(self.grandparents):
            self._gp[i].smallest, self._gp[i].smallest_len = a, len(a)
            self._gp[i].largest, self._gp[i].largest_len = b, len(b)
            self._gp[i].file_size = sz
        p.grandparents = C.cast(self._gp, C.POINTER(OrcGrandparent))
        p.num_grandparents = len(self.grandparents)
        p.level_compaction_dynamic_file_size = int(self.level_compaction_dynamic_file_size)
        tgt = self.target_output_file_size or self.max_output_file_size
        p.target_output_file_size = tgt
        p.bloom_millibits_per_key = self.bloom_millibits_per_key
        p.earliest_write_conflict_snapshot = self.earliest_write_conflict_snapshot
        if self.deeper_files is not None:
            self._deeper = (OrcGrandparent * max(1, len(self.deeper_files)))()
            for i, (a, b) in enumerate(self.deeper_files):
                self._deeper[i].smallest, self._deeper[i].smallest_len = a, len(a)
                self._deeper[i].largest, self._deeper[i].largest_len = b, len(b)
            p.key_not_exists_mode = 1
            p.deeper_files = C.cast(self._deeper, C.POINTER(OrcGrandparent))
            p.num_deeper_files = len(self.deeper_files)
        if self.range_start is not None:
            p.range_start, p.range_start_len, p.has_range_start = self.range_start, len(self.range_start), 1
        if self.range_end is not None:
            p.range_end, p.range_end_len, p.has_range_end = self.range_end, len(self.range_end), 1
        p.max_compaction_bytes = self.max_compaction_bytes or tgt * 25
        return p


def ikey(user_key: bytes, seq: int, vtype: int = 1) -> bytes:
    return user_key + struct.pack("<Q", (seq << 8) | vtype)


def kvstream(entries):
    """entries: iterable of (internal_key, value) -> kv stream bytes"""
    out = bytearray()
    for k, v in entries:
        out += struct.pack("<II", len(k), len(v)) + k + v
    return bytes(out)


def parse_kvstream(b):
    out = []
    p = 0
    while p + 8 <= len(b):
        kl, vl = struct.unpack_from("<II", b, p)
        p += 8
        out.append((bytes(b[p:p + kl]), bytes(b[p + kl:p + kl + vl])))
        p += kl + vl
    return out


def _take(ptr, n):
    return C.string_at(ptr, n) if n else b""


def oracle_sst_to_kv(data: bytes):
    L = oracle()
    out = C.POINTER(C.c_uint8)()
    n = C.c_size_t()
    cnt = C.c_uint64()
    rc = L.orc_sst_to_kvstream(data, C.c_size_t(len(data)), C.byref(out), C.byref(n), C.byref(cnt))
    assert rc == 0, (rc, L.orc_last_error())
    b = _take(out, n.value)
    L.orc_free(out)
    return b, cnt.value


def oracle_build_sst(params: Params, kv: bytes) -> bytes:
    L = oracle()
    out = C.POINTER(C.c_uint8)()
    n = C.c_size_t()
    cp = params.c()
    rc = L.orc_build_sst(C.byref(cp), kv, C.c_size_t(len(kv)), C.byref(out), C.byref(n))
    assert rc == 0, (rc, L.orc_last_error())
    b = _take(out, n.value)
    L.orc_free(out)
    return b


def oracle_citer(params: Params, kv: bytes):
    L = oracle()
    out = C.POINTER(C.c_uint8)()
    n = C.c_size_t()
    st = OrcStats()
    cp = params.c()
    rc = L.orc_compaction_iterator(C.byref(cp), kv, C.c_size_t(len(kv)), C.byref(out), C.byref(n), C.byref(st))
    if rc != 0:
        raise RuntimeError((rc, L.orc_last_error()))
    b = _take(out, n.value)
    L.orc_free(out)
    return b, st


def oracle_compact(params: Params, inputs):
    """inputs: list of SST file bytes, newest L0 run first.  Returns (files, metas, stats)."""
    L = oracle()
    cp = params.c()
    n = len(inputs)
    bufs = [C.create_string_buffer(x, len(x)) for x in inputs]
    ptrs = (C.c_void_p * max(1, n))(*[C.cast(b, C.c_void_p) for b in bufs])
    lens = (C.c_uint64 * max(1, n))(*[len(x) for x in inputs])
    res = C.c_void_p()
    rc = L.orc_compact(C.byref(cp), n, ptrs, lens, C.byref(res))
    if rc != 0:
        raise RuntimeError((rc, L.orc_last_error()))
    files, metas = [], []
    for i in range(L.orc_result_num_files(res)):
        ln = C.c_uint64()
        ptr = L.orc_result_file(res, i, C.byref(ln))
        files.append(_take(ptr, ln.value))
        m = OrcFileMeta()
        L.orc_result_meta(res, i, C.byref(m))
        metas.append(m)
    st = OrcStats()
    L.orc_result_stats(res, C.byref(st))
    L.orc_result_free(res)
    return files, metas, st


# ---------------------------------------------------------------- ops scripts + compiled reference
class Ops:
    """Write script for oracle/_ref/ref_compact (format: oracle/ops_format.md)."""

    def __init__(self):
        self.b = bytearray(b"B2OPS\0\0\1")

    def put(self, k, v):
        self.b += struct.pack("<BII", 1, len(k), len(v)) + k + v

    def delete(self, k):
        self.b += struct.pack("<BI", 2, len(k)) + k

    def single_delete(self, k):
        self.b += struct.pack("<BI", 7, len(k)) + k

    def flush(self):
        self.b += b"\x03"

    def snapshot(self):
        self.b += b"\x04"

    def compact_all_to(self, level):
        self.b += bytes([5, level])

    def set_time(self, seconds):
        self.b += struct.pack("<BI", 6, seconds)

    def bytes(self):
        return bytes(self.b) + b"\x00"


def have_ref():
    return os.path.exists(REF_BIN)


def run_reference(ops: Ops, workdir=None, binary=None, **opts):
    """Run the compiled reference on an ops script.  Returns dict(manifest, inputs[bytes], outputs[bytes])."""
    binary = binary or REF_BIN
    own = workdir is None
    if own:
        workdir = tempfile.mkdtemp(prefix="b200c_ref_")
    try:
        with open(os.path.join(workdir, "ops.bin"), "wb") as f:
            f.write(ops.bytes())
        args = [binary, os.path.join(workdir, "ops.bin"), os.path.join(workdir, "w")] + [f"{k}={v}" for k, v in opts.items()]
        subprocess.check_call(args, stdout=subprocess.DEVNULL)
        man = json.load(open(os.path.join(workdir, "w", "manifest.json")))
        ins = [open(os.path.join(workdir, "w", "inputs" + m["name"]), "rb").read() for m in man["inputs"]]
        outs = [open(os.path.join(workdir, "w", "outputs" + m["name"]), "rb").read() for m in man["outputs"]]
        return dict(manifest=man, inputs=ins, outputs=outs)
    finally:
        if own:
            shutil.rmtree(workdir, ignore_errors=True)


def params_from_reference(ref) -> Params:
    """Job parameters that reproduce a reference run: table options from the manifest, identity / clock /
    file-number fields read back from the reference's own output files (they are inputs to the job:
    TableBuilderOptions, db/compaction/compaction_job.cc:2258-2331)."""
    man = ref["manifest"]
    p = Params(output_level=man["output_level"], bottommost_level=man["bottommost_level"],
               max_output_file_size=man.get("max_output_file_size", man["target_file_size"]), block_size=man["block_size"],
               block_restart_interval=man["restart_interval"], format_version=man["format_version"],
               checksum=man["checksum"], snapshots=man["snapshots"], compaction_filter=man.get("compaction_filter", "none"), ttl=man.get("ttl", 0), now=man.get("now", 0),
               grandparents=[(bytes.fromhex(g["smallestkey"]), bytes.fromhex(g["largestkey"]), g["size"]) for g in man.get("grandparents", [])],
               level_compaction_dynamic_file_size=man.get("level_compaction_dynamic_file_size", True),
               max_compaction_bytes=man.get("max_compaction_bytes", 0), target_output_file_size=man.get("target_output_file_size", 0),
               bloom_millibits_per_key=man.get("bloom_millibits_per_key", 0))
    if ref["outputs"]:
        props = [sstfmt.parse_sst(o)["properties"] for o in ref["outputs"]]
        p0 = props[0]
        p.db_id = p0["rocksdb.creating.db.identity"].decode()
        p.db_session_id = p0["rocksdb.creating.session.identity"].decode()
        p.db_host_id = p0["rocksdb.creating.host.identity"].decode()
        p.column_family_name = p0["rocksdb.column.family.name"].decode()
        p.column_family_id = sstfmt.prop_u64(p0, "rocksdb.column.family.id")
        p.creation_time = sstfmt.prop_u64(p0, "rocksdb.creation.time")
        p.oldest_key_time = sstfmt.prop_u64(p0, "rocksdb.oldest.key.time")
        p.file_creation_times = [sstfmt.prop_u64(q, "rocksdb.file.creation.time") for q in props]
        p.first_file_number = sstfmt.prop_u64(p0, "rocksdb.original.file.number")
        nums = [sstfmt.prop_u64(q, "rocksdb.original.file.number") for q in props]
        # concurrent sub-compactions draw their file numbers from one counter, so only a single-range job numbers consecutively
        assert len(man.get("subcompactions", [])) > 1 or nums == list(range(nums[0], nums[0] + len(nums))), nums
    return p


def subcompaction_ranges(ref):
    """Key ranges [start, end) of the sub-compactions of a reference run (max_subcompactions > 1).  The reference reports per
    sub-compaction only statistics (SubcompactionJobInfo).  `total_input_raw_key_bytes` of sub-compaction i counts the keys its
    CompactionIterator consumed behind the ClippingIterator (compaction_job.cc:1495-1519,1676-1700) and grows with the range end, so
    the end of each range is found by bisection over the user keys of the input with the oracle's own accounting; the caller then
    checks files and the other statistics of every range.  Returns [(start or None, end or None, stats dict)]."""
    subs = ref["manifest"]["subcompactions"]
    if len(subs) < 2:
        return [(None, None, sb) for sb in subs]
    ukeys = sorted({ik[:-8] for data in ref["inputs"] for ik, _ in sstfmt.parse_sst(data)["entries"]})
    p = params_from_reference(ref)
    bounds, start = [], None
    for sb in subs[:-1]:
        want = sb["total_input_raw_key_bytes"]

        def consumed(end):
            p.range_start, p.range_end = start, end
            return oracle_compact(p, ref["inputs"])[2].total_input_raw_key_bytes

        lo, hi = 0, len(ukeys) - 1  # smallest end key whose range consumes at least `want`
        while lo < hi:
            mid = (lo + hi) // 2
            if consumed(ukeys[mid]) >= want:
                hi = mid
            else:
                lo = mid + 1
        assert consumed(ukeys[lo]) == want, "no user-key boundary reproduces the sub-compaction's input"
        bounds.append(ukeys[lo])
        start = ukeys[lo]
    return list(zip([None] + bounds, bounds + [None], subs))


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def golden_cases():
    return sorted(d for d in os.listdir(GOLDEN_DIR) if os.path.exists(os.path.join(GOLDEN_DIR, d, "manifest.json")))


def load_golden(name):
    d = os.path.join(GOLDEN_DIR, name)
    man = json.load(open(os.path.join(d, "manifest.json")))
    ins = [open(os.path.join(d, "inputs", m["name"].lstrip("/")), "rb").read() for m in man["inputs"]]
    outs = [open(os.path.join(d, "outputs", m["name"].lstrip("/")), "rb").read() for m in man["outputs"]]
    return dict(manifest=man, inputs=ins, outputs=outs)


STAT_KEYS = ("num_input_records", "num_output_records", "num_records_replaced", "num_expired_deletion_records",
             "num_input_deletion_records", "total_input_raw_key_bytes", "total_input_raw_value_bytes")


def sizes_without_file_number(outputs):
    """file sizes minus the width of rocksdb.original.file.number: on the RunRemote branch that property is the executor's own job-unique
    number (the DB renames the file to a number it allocates afterwards, compaction_job.cc:1022-1034), so its varint may be wider than
    the one the local path writes -- everything else in the files is compared byte for byte"""
    return [len(o) - len(sstfmt.parse_sst(o)["properties"]["rocksdb.original.file.number"]) for o in outputs]
