import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from toplingdb_b200 import synth
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
if mode == "after_cfg1":
    import pytest
    pytest.main(["-q", "-x", "tests/test_gpu_fullsize.py::test_cfg1_shape_against_the_reference_itself"])
if mode == "after_props":
    import pytest
    pytest.main(["-q", "-x", "tests/test_gpu_fullsize.py::test_cfg2_full_size_properties_and_idempotence"])
for i in range(3):
    try:
        images, _ = synth.stage_bench_inputs("cfg2", rank=0)
        print("stage ok", i, [int(t.numel()) for t in images][:3])
        del images
    except Exception as e:
        print("stage FAILED", i, e)
        break
