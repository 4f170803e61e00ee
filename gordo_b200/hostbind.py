"""
Bind a rank (process) to the host resources next to its GPU.

The fleet response path is PCIe- and host-DRAM-bound (gordo_b200/serving.py): pinned buffers must
live on the NUMA node the GPU's PCIe root hangs off, and the host threads that touch them must run
there, or every DMA crosses the inter-socket link.  On the 8 x B200 boxes GPUs 0-3 sit on node 0
(CPUs 0-31,64-95) and GPUs 4-7 on node 1 (`nvidia-smi topo -m`).

``bind_to_gpu(index, local_rank, local_world)``: CPU affinity = this rank's slice of the node's
CPUs (physical cores split evenly between the ranks that share the node, hyper-thread siblings kept
together), memory policy = prefer that node.  Best effort: every step degrades to a no-op (and is
reported) when sysfs / NVML / the syscall is unavailable.
"""
import ctypes
import os
from typing import Dict, List, Optional


def cpu_quota() -> Optional[float]:
    """CPUs' worth of time the cgroup grants this container (cpu.max / cfs quota), None when unlimited."""
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:                     # cgroup v2
            q, per = f.read().split()[:2]
        return None if q == "max" else float(q) / float(per)
    except Exception:
        pass
    try:                                                              # cgroup v1
        with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f:
            q = float(f.read())
        with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f:
            per = float(f.read())
        return None if q <= 0 else q / per
    except Exception:
        return None


def effective_cpus() -> int:
    """Host threads worth starting: the affinity mask capped by the cgroup CPU quota (a GPU lease is a
    container: it can see 128 CPUs and still be granted the time of a handful)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    q = cpu_quota()
    if q:
        ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))     # the quota is shared by the ranks of the box
        n = min(n, max(1, int(q / ranks + 0.5)))
    return max(1, n)


def _parse_cpulist(text: str) -> List[int]:
    out: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            out += list(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out


def gpu_numa_node(index: int) -> Optional[int]:
    """NUMA node of CUDA device ``index`` from sysfs (pci bus id via torch), None if unknown."""
    try:
        import torch
        p = torch.cuda.get_device_properties(index)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def gpu_cpu_affinity(index: int) -> Optional[List[int]]:
    """The CPUs NVML reports as local to the GPU (same list `nvidia-smi topo -m` prints)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        import torch
        uuid = torch.cuda.get_device_properties(index).uuid
        h = None
        for i in range(pynvml.nvmlDeviceGetCount()):
            hh = pynvml.nvmlDeviceGetHandleByIndex(i)
            u = pynvml.nvmlDeviceGetUUID(hh)
            u = u.decode() if isinstance(u, bytes) else u
            if str(uuid) in u:
                h = hh
                break
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [w * 64 + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        return cpus or None
    except Exception:
        return None


def node_cpus(node: int) -> Optional[List[int]]:
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            return _parse_cpulist(f.read())
    except Exception:
        return None


def _siblings(cpu: int) -> List[int]:
    try:
        with open(f"/sys/devices/system/cpu/cpu{cpu}/topology/thread_siblings_list") as f:
            return _parse_cpulist(f.read())
    except Exception:
        return [cpu]


def _set_preferred_node(node: int) -> bool:
    """set_mempolicy(MPOL_PREFERRED, {node}) through the raw syscall (no libnuma in the image)."""
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        MPOL_PREFERRED, SYS_set_mempolicy = 1, 238            # x86_64
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask)))
        return rc == 0
    except Exception:
        return False


def slice_cpus(cpus: List[int], part: int, parts: int) -> List[int]:
    """Rank ``part`` of ``parts`` gets an even share of the physical cores of ``cpus`` + their siblings."""
    allowed = set(cpus)
    cores, seen = [], set()
    for c in sorted(cpus):
        if c in seen:
            continue
        sib = [s for s in _siblings(c) if s in allowed] or [c]
        seen.update(sib)
        cores.append(sib)
    if parts <= 1 or len(cores) < parts:
        return sorted(cpus)
    lo, hi = len(cores) * part // parts, len(cores) * (part + 1) // parts
    return sorted(c for core in cores[lo:hi] for c in core)


def bind_to_gpu(index: int, local_rank: int = 0, local_world: int = 1, n_gpus_visible: Optional[int] = None) -> Dict:
    """Returns what was done: {"node", "cpus", "n_cpus", "mempolicy", "source"}."""
    info: Dict = {"node": None, "cpus": None, "n_cpus": None, "mempolicy": False, "source": None, "cpu_quota": cpu_quota()}
    try:
        avail = sorted(os.sched_getaffinity(0))
    except Exception:
        return info
    node = gpu_numa_node(index)
    cpus = node_cpus(node) if node is not None else None
    src = "sysfs"
    if not cpus:
        cpus = gpu_cpu_affinity(index)
        src = "nvml"
        if cpus and node is None:
            # which node do these CPUs belong to?
            for k in range(16):
                nc = node_cpus(k)
                if nc and set(cpus) <= set(nc):
                    node = k
                    break
    if not cpus:
        info["source"] = "none"
        info["n_cpus"] = len(avail)
        return info
    cpus = [c for c in cpus if c in set(avail)] or avail
    # ranks sharing the node: those whose GPU has the same node (assume local ranks map to GPU indices)
    share = [r for r in range(local_world) if (gpu_numa_node(r) if src == "sysfs" else None) == node] \
        if src == "sysfs" else list(range(local_world))
    if src != "sysfs" and local_world > 1:
        # NVML path: ranks whose GPU reports the same CPU list share the node
        share = [r for r in range(local_world) if (gpu_cpu_affinity(r) or []) == (gpu_cpu_affinity(index) or [])] or [local_rank]
    if local_rank not in share:
        share = sorted(set(share + [local_rank]))
    mine = slice_cpus(cpus, share.index(local_rank), len(share))
    try:
        os.sched_setaffinity(0, mine)
        info["cpus"] = mine
    except Exception:
        info["cpus"] = None
    info["n_cpus"] = len(mine)
    info["node"] = node
    info["source"] = src
    if node is not None:
        info["mempolicy"] = _set_preferred_node(node)
    return info
