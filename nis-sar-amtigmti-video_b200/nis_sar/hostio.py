"""Host side of the drop-in boundary: page-locked result buffers and CPU / NUMA placement of a rank.

The reference's functions return freshly allocated complex128 numpy arrays; the device holds
complex64.  One image of the bench frame is 1.07 GB on the host, so end to end the step is a PCIe transfer with a
0.3 ms widening kernel in front of it; what this module controls is WHERE that transfer lands:

* ``to_host_c128(t, out=None)``: widen on the device, one asynchronous D2H copy into page-locked memory.  Without
  ``out`` the block comes from torch's caching host allocator (re-used once the caller drops the previous result --
  no ``cudaHostAlloc`` per call in steady state); with ``out`` (``pinned_empty``) the caller owns a persistent buffer
  and nothing is allocated at all.
* ``bind_rank_to_numa(local_rank, ...)``: one process per GPU on a two-socket host -- pin the process (and with it the
  first-touch placement of every pinned block it allocates afterwards) to the NUMA node of its GPU, or, when the
  platform reports every GPU on one node, spread the ranks over the nodes so that eight result streams do not share one
  memory controller.  Measured numbers: DESIGN.md section 5.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import device as dev
from ._lib import NisError


def pinned_empty(shape, dtype=np.complex128, order="C") -> np.ndarray:
    """Page-locked numpy array (kept alive by the torch tensor it views).  ``order="F"`` gives the layout of the
    reference's ``img.T`` (sar_ati_dcpa_sim_csa.py:396)."""
    shape = tuple(int(s) for s in shape)
    tdt = {np.dtype(np.complex128): torch.complex128, np.dtype(np.complex64): torch.complex64,
           np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32}[np.dtype(dtype)]
    if order == "F":
        return torch.empty(shape[::-1], dtype=tdt, pin_memory=True).numpy().T
    return torch.empty(shape, dtype=tdt, pin_memory=True).numpy()


def to_host_c128(t: torch.Tensor, out: np.ndarray | None = None) -> np.ndarray:
    """complex64 CUDA tensor -> complex128 numpy array of the same shape (C order)."""
    if not t.is_cuda or t.dtype != torch.complex64:
        raise NisError("to_host_c128: expected a complex64 CUDA tensor")
    wide = dev.widen_c32(t.contiguous())
    if out is None:
        host = torch.empty(wide.shape, dtype=torch.complex128, pin_memory=True)
    else:
        if out.dtype != np.complex128 or tuple(out.shape) != tuple(wide.shape) or not out.flags.c_contiguous:
            raise NisError(f"to_host_c128: out must be a C-contiguous complex128 array of shape {tuple(wide.shape)}")
        host = torch.from_numpy(out)          # pinned if it came from pinned_empty(): the copy below is then one DMA
    host.copy_(wide, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return host.numpy() if out is None else out


# ------------------------------------------------------------------------------------------ placement
def _cpulist(text: str) -> list[int]:
    cpus: list[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def numa_nodes() -> dict[int, list[int]]:
    """{node: cpus} from sysfs (empty dict when the platform does not expose it)."""
    base = "/sys/devices/system/node"
    out = {}
    try:
        for name in sorted(os.listdir(base)):
            if name.startswith("node") and name[4:].isdigit():
                with open(os.path.join(base, name, "cpulist")) as fh:
                    cpus = _cpulist(fh.read())
                if cpus:
                    out[int(name[4:])] = cpus
    except OSError:
        return {}
    return out


def gpu_numa_node(device_index: int) -> int:
    """NUMA node of a CUDA device from its PCI address (-1: unknown)."""
    bdf = None
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{int(pr.pci_domain_id):04x}:{int(pr.pci_bus_id):02x}:{int(pr.pci_device_id):02x}.0"
    except Exception:
        try:
            import pynvml
            pynvml.nvmlInit()
            bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device_index)).busId
            if isinstance(bdf, bytes):
                bdf = bdf.decode()
            bdf = str(bdf).lower()
            if len(bdf.split(":")[0]) == 8:          # NVML prints an 8-digit domain; sysfs uses 4
                bdf = bdf[4:]
        except Exception:
            return -1
    try:
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            return int(fh.read().strip())
    except (OSError, ValueError):
        return -1


def bind_rank_to_numa(local_rank: int, world: int, device_index: int | None = None, policy: str = "auto") -> dict:
    """Restrict this process to the CPUs of one NUMA node (first-touch then places its pinned buffers there).
    policy "gpu": the GPU's own node; "spread": node = local_rank * n_nodes // world; "auto": "gpu" when the GPUs of the
    job report different nodes, else "spread"; "none": leave the affinity alone.  Returns what was done (for the bench
    record); never raises on platforms without NUMA information."""
    nodes = numa_nodes()
    info = {"policy": policy, "nodes": len(nodes), "node": None, "cpus": None}
    if policy == "none" or len(nodes) < 2:
        return info
    try:
        allowed = set(os.sched_getaffinity(0))
    except AttributeError:
        return info
    gpu_nodes = [gpu_numa_node(i) for i in range(min(world, torch.cuda.device_count()))]
    if policy == "auto":
        policy = "gpu" if len({n for n in gpu_nodes if n >= 0}) > 1 else "spread"
    dev_i = local_rank if device_index is None else device_index
    if policy == "gpu":
        node = gpu_nodes[dev_i] if dev_i < len(gpu_nodes) else -1
    else:
        node = sorted(nodes)[local_rank * len(nodes) // max(world, 1)]
    info.update(policy=policy, gpu_nodes=gpu_nodes)
    if node not in nodes:
        return info
    cpus = sorted(allowed & set(nodes[node]))
    if not cpus:
        return info
    os.sched_setaffinity(0, cpus)
    info.update(node=node, cpus=len(cpus))
    return info
