"""NUMA placement of a rank's host buffers next to its GPU.

On the 8-GPU boxes GPU0-3 hang off NUMA node 0 and GPU4-7 off node 1; a pinned staging buffer that lives on the other node makes every
H2D / D2H byte cross the socket interconnect, and with 8 ranks copying at once that link -- not PCIe -- is what the end-to-end
throughput runs into (round-1 SCALE: efficiency 0.35 at 8 GPUs).  Before a rank allocates pinned memory it therefore
  * restricts itself to the CPUs of its GPU's node (if the cgroup allows any of them), and
  * sets its memory policy to that node (set_mempolicy, preferred), so that cudaHostAlloc's pages come from it.
Plumbing only (the plugin does the same in C++, plugin/b200_compaction_executor.cc)."""
import ctypes
import os

_SYS_set_mempolicy = 238  # x86_64
MPOL_DEFAULT, MPOL_PREFERRED, MPOL_BIND = 0, 1, 2


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def parse_cpulist(s):
    cpus = set()
    for part in (s or "").split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_pci_address(index):
    """'0000:1b:00.0' of CUDA device `index` (None when it cannot be determined)"""
    try:
        import pynvml as N
        N.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = index
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if index < len(ids) and ids[index].isdigit():
                phys = int(ids[index])
        h = N.nvmlDeviceGetHandleByIndex(phys)
        bus = N.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        return bus.lower()[-12:]  # NVML prints an 8-digit domain
    except Exception:
        pass
    try:
        import torch
        p = torch.cuda.get_device_properties(index)
        return "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    except Exception:
        return None


def gpu_numa_node(index):
    addr = gpu_pci_address(index)
    if not addr:
        return None
    v = _read(f"/sys/bus/pci/devices/{addr}/numa_node")
    try:
        n = int(v)
    except (TypeError, ValueError):
        return None
    return n if n >= 0 else None


def set_preferred_node(node):
    """memory policy of the calling thread: allocate on `node` first (falls back to other nodes instead of failing)"""
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        nbits = max(64, node + 2)
        words = (nbits + 63) // 64
        mask = (ctypes.c_ulong * words)()
        mask[node // 64] = 1 << (node % 64)
        rc = libc.syscall(_SYS_set_mempolicy, MPOL_PREFERRED, mask, ctypes.c_ulong(nbits + 1))
        return rc == 0
    except Exception:
        return False


def bind_to_gpu_node(index):
    """Call before allocating pinned host memory for GPU `index`.  Returns a dict describing what was done (for the bench line)."""
    info = {"gpu": index, "node": None, "cpus": None, "mempolicy": False}
    node = gpu_numa_node(index)
    if node is None:
        return info
    info["node"] = node
    cpus = parse_cpulist(_read(f"/sys/devices/system/node/node{node}/cpulist"))
    try:
        allowed = os.sched_getaffinity(0)
        mine = sorted(cpus & allowed)
        if mine:
            os.sched_setaffinity(0, mine)
            info["cpus"] = len(mine)
    except OSError:
        pass
    info["mempolicy"] = set_preferred_node(node)
    return info
