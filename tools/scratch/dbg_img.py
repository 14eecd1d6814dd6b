import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, numpy as np
import helpers as H
import test_gpu_fullsize as F
from toplingdb_b200 import synth, sharding
workload, scale, r = "cfg3", 1.0/64, 0
w = synth.WORKLOADS[workload]
_, n_total = synth.bench_shape(workload, scale)
images, _ = synth.stage_runs(n_total, w["k"], w["vlen"], key_base=0, seed=2, overlap=w["overlap"], del_frac=w["del_frac"])
img = images[r].cpu().numpy().tobytes()
kv_img, cnt = H.oracle_sst_to_kv(img)
kv = F._run_kvstream(n_total, w["k"], r, w["vlen"], 0, 2, w["overlap"], w["del_frac"])
a = H.parse_kvstream(kv_img); b = H.parse_kvstream(kv)
print("entries image", len(a), "kv", len(b))
nd = 0
for i, (x, y) in enumerate(zip(a, b)):
    if x != y:
        nd += 1
        if nd <= 5:
            print(i, x[0].hex(), len(x[1]), "|", y[0].hex(), len(y[1]), x[1][:8].hex(), y[1][:8].hex())
print("differences", nd)
