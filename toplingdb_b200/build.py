"""Builds toplingdb_b200/libb200c.so (the C-ABI library of include/b200c.h) in-tree with nvcc for sm_100a.
nvcc cross-compiles without a GPU, so this also runs in the CPU-only build container."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libb200c.so")
OBJ = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CUFLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xptxas", "-v"]
SOURCES = ["decode.cu", "merge.cu", "encode.cu", "api.cu", "sst_host.cc"]


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def build_native(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "b200c.h"))
    objs, jobs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.rsplit(".", 1)[0] + ".o")
        objs.append(obj)
        if force or _newer([src] + headers, obj):
            jobs.append([NVCC] + ARCH + CUFLAGS + ["-x", "cu", "-c", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), r.stdout))
        return r.stdout

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        logs = list(ex.map(run, jobs))
    if verbose:
        sys.stdout.write("\n".join(logs))
    if jobs or force or not os.path.exists(OUT):
        with open(os.path.join(OBJ, "ptxas.log"), "w") as f:
            f.write("\n".join(logs))
        run([NVCC] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", OUT] + objs + ["-lcudart"])
    return OUT


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
