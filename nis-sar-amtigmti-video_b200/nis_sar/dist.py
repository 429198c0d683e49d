"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed for the plumbing).

The path shards along independent units; a collective appears only where the
algorithm has a real exchange step:

  echo, few scatterers / many pulses   pulse blocks            no collective (optional all-gather of rows)
  echo, dense scatterer scenes         scatterer shards        sum all-reduce / reduce of the partial echoes
  CSA, VideoSAR frame sequences        whole frames            none
  HRWS / ATI receive channels          one channel per rank    neighbour exchange of one SLC, then K3 per pair

Every function takes the per-rank compute step as a callable, so the same partition / collective logic
runs under NCCL with the CUDA operators (``nis_sar.device``) and, in the CPU test-suite, under gloo.
Complex tensors travel as their float32 (re, im) view: NCCL reduces real dtypes.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np
import torch
import torch.distributed as dist


def world(group=None):
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def block_range(n: int, rank: int, nranks: int):
    """Contiguous block partition of range(n): the first n % nranks ranks get one extra unit."""
    base, extra = divmod(n, nranks)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _as_real(t: torch.Tensor) -> torch.Tensor:
    return torch.view_as_real(t) if t.is_complex() else t


# ------------------------------------------------------------------------------------------ echo
def echo_pulse_blocks(compute: Callable[[int, int, torch.Tensor], None], raw: torch.Tensor, gather: bool = False,
                      group=None) -> torch.Tensor:
    """Pulse-block sharding: rank r fills rows [p0, p1) of ``raw`` ([P, S] complex64, same shape on every
    rank) by calling ``compute(p0, p1, raw)``.  With ``gather`` the blocks are exchanged so that every
    rank ends with the whole aperture (needed only when one rank must focus the full frame)."""
    rank, n = world(group)
    P = raw.shape[0]
    p0, p1 = block_range(P, rank, n)
    if p1 > p0:
        compute(p0, p1, raw)
    if gather and n > 1:
        # equal-sized padded blocks for all_gather_into_tensor
        blk = -(-P // n)
        send = torch.zeros((blk,) + tuple(raw.shape[1:]), dtype=raw.dtype, device=raw.device)
        send[: p1 - p0] = raw[p0:p1]
        recv = torch.empty((n * blk,) + tuple(raw.shape[1:]), dtype=raw.dtype, device=raw.device)
        dist.all_gather_into_tensor(_as_real(recv), _as_real(send), group=group)
        for r in range(n):
            q0, q1 = block_range(P, r, n)
            raw[q0:q1] = recv[r * blk: r * blk + (q1 - q0)]
    return raw


def echo_scatterer_shards(compute: Callable[[int, int], torch.Tensor], num_scatterers: int, dst: int | None = None,
                          group=None, comm: "CComm | None" = None) -> torch.Tensor:
    """Scatterer sharding (config 3: 1e5 scatterers): rank r synthesises the partial echo of scatterers
    [t0, t1) over the whole [P, S] grid with ``compute(t0, t1)`` and the partial sums are added across
    ranks -- ``dst`` None: all-reduce (every rank gets the echo); else reduce to rank ``dst``.  ``comm``: run the
    reduction through the library's own entry point (``nis_echo_reduce``) instead of torch.distributed.
    fp32 summation order differs from the one-GPU run, so results agree to rounding, not bit-wise."""
    rank, n = world(group)
    t0, t1 = block_range(num_scatterers, rank, n)
    part = compute(t0, t1)
    if n > 1:
        if comm is not None:
            comm.echo_reduce(part, root=-1 if dst is None else dst)
        elif dst is None:
            dist.all_reduce(_as_real(part), op=dist.ReduceOp.SUM, group=group)
        else:
            dist.reduce(_as_real(part), dst=dst, op=dist.ReduceOp.SUM, group=group)
    return part


class CComm:
    """The library's own communicator (``nis_comm_*``, include/nis_sar.h): NCCL bound inside libnis_sar.so, driven through
    the C ABI on torch's current stream.  torch.distributed is only the control plane that carries the 128-byte id from
    rank 0 to the others -- a reference maintainer without torch.distributed would send it over MPI or a file.
    Collective: every rank of ``group`` constructs it."""

    def __init__(self, device=None, group=None):
        import ctypes as C
        from . import _lib
        self.rank, self.n = world(group)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        lib = _lib.load()
        ident = (C.c_uint8 * 128)()
        if self.rank == 0:
            _lib.check(lib.nis_comm_unique_id(C.cast(ident, C.c_void_p)), "nis_comm_unique_id")
        box = [bytes(ident)]
        if self.n > 1:
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = (C.c_uint8 * 128).from_buffer_copy(box[0])
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.nis_comm_init(self.rank, max(self.n, 1), C.cast(ident, C.c_void_p), C.byref(self._h)),
                       "nis_comm_init")

    def _stream(self):
        import ctypes as C
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def echo_reduce(self, raw: torch.Tensor, root: int = -1) -> torch.Tensor:
        """Sum the partial echoes ``raw`` (complex64, same shape on every rank) in place: all ranks (root = -1) or root."""
        import ctypes as C
        from . import _lib
        if raw.dtype != torch.complex64 or not raw.is_contiguous():
            raise _lib.NisError("CComm.echo_reduce: raw must be a contiguous complex64 tensor")
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().nis_echo_reduce(self._h, C.c_void_p(raw.data_ptr()), raw.numel(), int(root), self._stream()),
                       "nis_echo_reduce")
        return raw

    def slc_exchange(self, slc_local: torch.Tensor) -> torch.Tensor | None:
        """Ring shift for channel pairing: returns channel rank+1's image (None on the last rank)."""
        import ctypes as C
        from . import _lib
        if slc_local.dtype != torch.complex64 or not slc_local.is_contiguous():
            raise _lib.NisError("CComm.slc_exchange: slc_local must be a contiguous complex64 tensor")
        recv = torch.empty_like(slc_local) if self.rank < self.n - 1 else None
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().nis_slc_exchange(self._h, C.c_void_p(slc_local.data_ptr()),
                                                    C.c_void_p(recv.data_ptr() if recv is not None else 0),
                                                    slc_local.numel(), self._stream()), "nis_slc_exchange")
        return recv

    def close(self):
        from . import _lib
        if getattr(self, "_h", None) is not None and self._h.value:
            torch.cuda.synchronize(self.device)
            _lib.load().nis_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------ peer memory over NVLink / NVSwitch
class _CudaArray:
    """``__cuda_array_interface__`` carrier: lets torch view a raw device pointer as a tensor (no copy, no ownership)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


_TYPESTR = {torch.complex64: "<c8", torch.float32: "<f4", torch.float64: "<f8", torch.complex128: "<c16"}


class SharedBuffer:
    """One buffer per rank, each mapped into every rank (CUDA IPC): ``local`` is this rank's tensor, ``views[r]`` aliases
    rank r's buffer.  Kernels launched by this rank load from / reduce into ``views[r]`` directly over NVLink -- the
    transfer overlaps the arithmetic access by access instead of preceding it as a copy.  Collective: every rank of the
    group constructs it with the same shape; ``close()`` (collective as well) unmaps and frees."""

    def __init__(self, shape, dtype=torch.complex64, device=None, group=None):
        import ctypes as C
        from . import _lib
        self.group = group
        rank, n = world(group)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        lib = _lib.load()
        nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        self._ptr = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        with torch.cuda.device(self.device):
            _lib.check(lib.nis_peer_alloc(nbytes, C.byref(self._ptr), C.cast(handle, C.c_void_p)), "nis_peer_alloc")
            self.local = torch.as_tensor(_CudaArray(self._ptr.value, shape, _TYPESTR[dtype]), device=self.device)
            handles = [None] * n
            devs = [None] * n
            if n > 1:
                dist.all_gather_object(handles, bytes(handle), group=group)
                dist.all_gather_object(devs, int(self.device.index), group=group)
            else:
                devs = [int(self.device.index)]
            # accumulate="atomic" into a peer's rows is defined only where the link performs system-scope atomics natively
            # (NVLink does; PCIe peers may not): callers check this and fall back to the NCCL reduce (echo_scatterer_shards)
            self.native_atomics = all(lib.nis_peer_native_atomics(int(self.device.index), int(d)) == 1 for d in devs)
            self._opened = {}
            self.views = []
            for r in range(n):
                if r == rank:
                    self.views.append(self.local)
                    continue
                p = C.c_void_p()
                hb = (C.c_uint8 * 64).from_buffer_copy(handles[r])
                _lib.check(lib.nis_peer_open(C.cast(hb, C.c_void_p), C.byref(p)), "nis_peer_open")
                self._opened[r] = p
                self.views.append(torch.as_tensor(_CudaArray(p.value, shape, _TYPESTR[dtype]), device=self.device))

    def close(self):
        from . import _lib
        lib = _lib.load()
        peer_barrier(self.device, self.group, host_sync=True)
        with torch.cuda.device(self.device):
            for p in self._opened.values():
                lib.nis_peer_close(p)
            self._opened = {}
            self.views = []
            peer_barrier(self.device, self.group, host_sync=True)     # everybody has unmapped before anybody frees
            if self._ptr:
                lib.nis_peer_free(self._ptr)
                self._ptr = None
            self.local = None


_barrier_token: dict = {}


def peer_barrier(device, group=None, host_sync: bool = False) -> None:
    """Producer -> consumer hand-off between ranks that read / reduce into each other's HBM.  Under NCCL this is a
    one-element all-reduce enqueued on the current stream: work enqueued afterwards starts only when every rank's earlier
    work on its stream has finished -- no host synchronisation, the CPU keeps enqueueing.  ``host_sync`` (and any other
    backend): device synchronise + host barrier, for teardown paths that free memory."""
    n = world(group)[1]
    if host_sync or n == 1 or dist.get_backend(group) != "nccl":
        torch.cuda.synchronize(device)
        if n > 1:
            dist.barrier(group=group)
        return
    key = (torch.device(device).index, id(group))
    tok = _barrier_token.get(key)
    if tok is None:
        tok = _barrier_token[key] = torch.zeros(1, dtype=torch.float32, device=device)
    dist.all_reduce(tok, group=group)


def echo_scatterer_shards_p2p(compute_into: Callable[[int, int, int, int, torch.Tensor], None], num_scatterers: int,
                              shared: SharedBuffer) -> tuple:
    """Scatterer sharding with the reduction fused into the synthesis kernel (reduce-scatter semantics): rank b owns the
    pulse block b of ``shared`` (a SharedBuffer of shape [P, S] complex64).  Every rank synthesises ITS scatterers
    [t0, t1) for ALL pulse blocks and the kernel's epilogue adds each block straight into the owner's HBM with
    fire-and-forget reductions over NVLink -- ``compute_into(t0, t1, p0, p1, dst)`` must accumulate atomically into rows
    [p0, p1) of ``dst`` (``dev.echo_accumulate(..., out=dst, pulse_range=(p0, p1), accumulate="atomic")``).  No partial
    [P, S] echo is materialised and no separate collective runs.  Returns (p0, p1): the fully reduced rows of
    ``shared.local`` this rank holds."""
    rank, n = world(shared.group)
    if not getattr(shared, "native_atomics", True):
        from ._lib import NisError
        raise NisError("echo_scatterer_shards_p2p: a peer link of this job does not perform atomics natively "
                       "(cudaDevP2PAttrNativeAtomicSupported = 0); use echo_scatterer_shards (NCCL reduce)")
    P = shared.local.shape[0]
    t0, t1 = block_range(num_scatterers, rank, n)
    shared.local.zero_()
    peer_barrier(shared.device, shared.group)
    for k in range(n):
        b = (rank + k) % n                      # start with the local block, then walk the ring: spreads the NVLink load
        p0, p1 = block_range(P, b, n)
        if p1 > p0 and t1 > t0:
            compute_into(t0, t1, p0, p1, shared.views[b])
    peer_barrier(shared.device, shared.group)
    return block_range(P, rank, n)


def pair_products_p2p(shared: SharedBuffer, products: Callable[[torch.Tensor, torch.Tensor], dict],
                      first: torch.Tensor | None = None):
    """HRWS-style chain without the staging copy: every rank has focused its channel into ``shared.local``; the fused
    DPCA/ATI kernel of rank k then reads channel k+1 directly from rank k+1's HBM over NVLink (8 of its 16 input bytes per
    pixel), so the exchange overlaps the products instead of preceding them.  Returns the products of the pair (k, k+1),
    None on the last rank.  ``first`` replaces ``shared.local`` as the pair's first image when the two members of a pair
    are focused from different pulse windows of a channel (the one-pulse DPCA shift, sar_ati_dcpa_sim_csa.py:402-403:
    rank k keeps focus(raw_k[1:]) private and publishes focus(raw_k[:-1]) for rank k-1)."""
    rank, n = world(shared.group)
    peer_barrier(shared.device, shared.group)      # every channel is focused before anyone reads it
    a = shared.local if first is None else first
    out = products(a, shared.views[rank + 1]) if rank < n - 1 else None
    peer_barrier(shared.device, shared.group)      # nobody overwrites its channel while a neighbour still reads it
    return out


# ------------------------------------------------------------------------------------------- CSA
def frame_indices(num_frames: int, group=None) -> range:
    """Round-robin frame ownership for VideoSAR sequences: rank r focuses frames r, r+n, r+2n, ..."""
    rank, n = world(group)
    return range(rank, num_frames, n)


def focus_frames(focus: Callable[[int], torch.Tensor], num_frames: int, group=None) -> dict:
    """Frame-parallel focusing: no data-path collective.  Returns {frame index: focused SLC} for the
    frames this rank owns."""
    return {f: focus(f) for f in frame_indices(num_frames, group)}


# ------------------------------------------------------------------------------- channel pairing
def exchange_with_next(slc_local: torch.Tensor, group=None) -> torch.Tensor | None:
    """One receive channel per rank: returns the SLC of channel rank+1 (None on the last rank) so that
    rank k can form the DPCA/ATI pair (k, k+1).  A ring shift of one [N_rg, N_az] complex64 image
    (134 MB at 4096^2) over NVLink: isend to rank-1, irecv from rank+1."""
    rank, n = world(group)
    if n == 1:
        return None
    ops = []
    recv = None
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, _as_real(slc_local.contiguous()), rank - 1, group=group))
    if rank < n - 1:
        recv = torch.empty_like(slc_local)
        ops.append(dist.P2POp(dist.irecv, _as_real(recv), rank + 1, group=group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return recv


def pair_products(slc_local: torch.Tensor, products: Callable[[torch.Tensor, torch.Tensor], dict], group=None):
    """HRWS-style chain: exchange with the neighbour, then run the fused DPCA/ATI stage on the pair
    (k, k+1) held by rank k.  The detection threshold is per reference channel, so no global reduction."""
    other = exchange_with_next(slc_local, group)
    return None if other is None else products(slc_local, other)


def max_over_ranks(value: float, device, group=None) -> float:
    """Timing reduction used by bench.py: device-measured milliseconds, maximum over ranks."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if world(group)[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
