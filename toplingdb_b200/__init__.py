"""toplingdb_b200 — B200-native implementation of ToplingDB's compaction hot path
(BlockBasedTable decode -> k-way merge with the compaction-iterator rules -> BlockBasedTable encode) behind the
C ABI of include/b200c.h.  The Python layer is plumbing for tests and bench.py (ctypes over libb200c.so, torch for
device buffers); the product is the CUDA library and the C++ plugin mirror in toplingdb_b200/plugin/."""
from .native import (B200cError, CompactionJob, FileMeta, JobStats, block_checksums, device_count, lib, load_library)  # noqa: F401
