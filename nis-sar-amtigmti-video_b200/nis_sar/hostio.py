"""Host side of the drop-in boundary: page-locked result buffers and CPU / NUMA placement of a rank.

The reference's functions return freshly allocated complex128 numpy arrays; the device holds
complex64.  One image of the bench frame is 1.07 GB on the host, so end to end the step is a PCIe transfer with a
0.3 ms widening kernel in front of it; what this module controls is WHERE that transfer lands:

* ``to_host_c128(t, out=None)``: two routes.  With enough host cores (``host_threads() >= 6``) the NARROW form crosses
  PCIe -- half the bytes -- in 4 MiB chunks and worker threads widen float -> double straight into the result while
  the DMA engine runs (``nis_d2h_widen``, csrc/hostcopy.cpp): 19 -> 12-14 ms for the 8192 x 8192 image.  Otherwise:
  widen on the device, one asynchronous D2H copy of the wide form.  Without ``out`` the block comes from torch's caching
  host allocator (re-used once the caller drops the previous result -- no ``cudaHostAlloc`` per call in steady state);
  with ``out`` (``pinned_empty`` or any C-contiguous complex128 array) nothing is allocated at all.
* ``to_device_c64(h, device)``: the same for inputs -- complex128 numpy -> complex64 on the device, narrowed on the host
  cores into the transfer ring (``nis_h2d_narrow``) instead of a pageable copy of the wide form + a device kernel.
  ``NIS_HOST_THREADS`` sets the worker count (0: always convert on the device).
* ``bind_rank_to_numa(local_rank, ...)``: one process per GPU on a two-socket host -- pin the process (and with it the
  first-touch placement of every pinned block it allocates afterwards) to the NUMA node of its GPU, or, when the
  platform reports every GPU on one node, spread the ranks over the nodes so that eight result streams do not share one
  memory controller.  Measured numbers: DESIGN.md section 5.
"""
from __future__ import annotations

import os

import numpy as np
import torch

import ctypes as C

from . import _lib, device as dev
from ._lib import NisError

_MIN_THREADS_FOR_HOST_ROUTE = 6      # measured: 4 threads tie with the wide DMA, 8 win by 20 %, 12-16 by 35 %
_MIN_ELEMENTS_FOR_HOST_ROUTE = 1 << 21


def host_threads() -> int:
    """Worker threads of the host-converted transfers: ``NIS_HOST_THREADS``, else the cores this process may run on
    divided among the ranks of the node (``LOCAL_WORLD_SIZE``), two left for the thread that drives the DMA and the
    runtime's own threads (spinning workers on an over-subscribed host cost 2-3 x: tools/e2e_route_probe.py), at most 12."""
    env = os.environ.get("NIS_HOST_THREADS")
    if env is not None:
        try:
            return max(0, min(32, int(env)))
        except ValueError:
            pass
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
    return max(0, min(12, avail // ranks - 2))     # measured: 8-12 workers saturate the host memory path


# what the last to_host_c128 / to_device_c64 call did (bench.py reports the bytes that really crossed PCIe)
last_transfer = {"d2h": None, "h2d": None}


def _host_route(n_elements: int) -> int:
    t = host_threads()
    return t if (t >= _MIN_THREADS_FOR_HOST_ROUTE and n_elements >= _MIN_ELEMENTS_FOR_HOST_ROUTE) else 0


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
    t = t.contiguous()
    if out is not None and (out.dtype != np.complex128 or tuple(out.shape) != tuple(t.shape) or not out.flags.c_contiguous):
        raise NisError(f"to_host_c128: out must be a C-contiguous complex128 array of shape {tuple(t.shape)}")
    threads = _host_route(t.numel())
    if threads:
        host = torch.empty(t.shape, dtype=torch.complex128, pin_memory=True).numpy() if out is None else out
        di = dev._dev_index(t.device)
        with torch.cuda.device(di):
            rc = _lib.load().nis_d2h_widen(_lib.context(di), dev._ptr(t), C.c_void_p(host.ctypes.data), t.numel(), threads,
                                           C.c_void_p(dev._stream_ptr(di)))
        _lib.check(rc, "nis_d2h_widen")
        last_transfer["d2h"] = {"route": "complex64 over PCIe, widened by host threads (nis_d2h_widen)", "threads": threads,
                                "pcie_bytes": 8 * t.numel(), "host_bytes": 16 * t.numel()}
        return host
    wide = dev.widen_c32(t)
    host = torch.empty(wide.shape, dtype=torch.complex128, pin_memory=True) if out is None else torch.from_numpy(out)
    host.copy_(wide, non_blocking=True)       # one DMA when the destination is page-locked
    torch.cuda.current_stream(t.device).synchronize()
    last_transfer["d2h"] = {"route": "widened on the device, complex128 over PCIe", "threads": 0,
                            "pcie_bytes": 16 * t.numel(), "host_bytes": 16 * t.numel()}
    return host.numpy() if out is None else out


def to_device_c64(h: np.ndarray, device) -> torch.Tensor:
    """C-contiguous complex128 (or complex64) numpy array -> complex64 CUDA tensor of the same shape."""
    h = np.ascontiguousarray(h)
    if h.dtype == np.complex64:
        return torch.from_numpy(h).to(device)
    if h.dtype != np.complex128:
        h = h.astype(np.complex128)
    threads = _host_route(h.size)
    if not threads:
        return dev.narrow_c128(torch.from_numpy(h).to(device))
    x = torch.empty(h.shape, dtype=torch.complex64, device=device)
    di = dev._dev_index(x.device)
    with torch.cuda.device(di):
        rc = _lib.load().nis_h2d_narrow(_lib.context(di), C.c_void_p(h.ctypes.data), dev._ptr(x), h.size, threads,
                                        C.c_void_p(dev._stream_ptr(di)))
    _lib.check(rc, "nis_h2d_narrow")
    last_transfer["h2d"] = {"route": "narrowed by host threads, complex64 over PCIe (nis_h2d_narrow)", "threads": threads,
                            "pcie_bytes": 8 * h.size, "host_bytes": 16 * h.size}
    return x


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
