import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, numpy as np
import sstfmt
from toplingdb_b200 import synth, sharding
from toplingdb_b200.native import CompactionJob
w = synth.WORKLOADS["cfg2"]
_, n_total = synth.bench_shape("cfg2", 1.0)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(2)
values = torch.randint(0, 256, (n_total * 32,), dtype=torch.uint8, device=dev, generator=g)
n, pfx, tr, vref, meta = synth.make_run_columns(n_total, 8, 0, 32, 0, values, 0.0, 0.0, 2, dev)
ref = None
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    job = CompactionJob(device=0, output_level=0, output_mem="device", checksum="xxh3", file_creation_times=[1700000000], first_file_number=1000)
    job.encode_columns(n, pfx, tr, vref, meta)
    m = job.output_meta(0)
    img = torch.empty(m.file_size, dtype=torch.uint8, device=dev)
    job.output_read_into(0, img)
    job.close()
    a = img.cpu().numpy()
    if ref is None:
        ref = a; print("it", it, "size", a.size, "blocks", m.num_data_blocks, "data", m.data_size, "index", m.index_size); continue
    if a.size != ref.size or not np.array_equal(a, ref):
        k = min(a.size, ref.size)
        d = np.flatnonzero(a[:k] != ref[:k])
        print("it", it, "DIFF size", a.size, "vs", ref.size, "first diff at", int(d[0]) if d.size else k, "ndiff", d.size, "blocks", m.num_data_blocks, "data_size", m.data_size, "index_size", m.index_size)
    else:
        print("it", it, "same")
