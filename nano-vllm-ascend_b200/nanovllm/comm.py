"""Peer-memory communicator for the fused tensor-parallel all-reduce (C1).

One cudaMalloc'd region per rank, mapped into every other rank with CUDA IPC (NVLink P2P):
  [ partial buffer 0 | partial buffer 1 | gather buffer | flags (uint32[16]) | state (uint32[4]) ]
The row-parallel GEMMs write their partial products straight into partial buffer `parity`; the
fused kernel reads every rank's copy: one-shot (b200_allreduce_add_rmsnorm, every rank pulls all
copies) for small messages, two-shot (b200_allreduce_add_rmsnorm_twoshot, each rank reduces a slice
of the rows and pushes the sums into every rank's gather buffer) for larger ones. Set-up uses
torch.distributed only to exchange the 64-byte IPC handles.

Form selection (``mode``): one-shot while (world-1) x message <= 4 MB, two-shot beyond (both validated on
hardware inside the engine, eager and CUDA-graph decode, r02), NCCL for shapes neither takes.
``B200_ALLREDUCE`` forces one form for every shape it can take -- a measurement aid for A/B runs:
``nccl`` | ``oneshot`` | ``twoshot`` | ``nvls`` | ``nvls_fused`` | ``nvls_one``. The NVLS values make the region a
torch symmetric-memory allocation (plumbing: cuMemCreate + handle exchange + cuMulticastBindMem done by
``torch.distributed._symmetric_memory``), which yields the same per-rank peer pointers PLUS one multicast
address for the whole region; messages then go through b200_allreduce_add_rmsnorm_nvls[_fused]
(in-switch reduce + in-switch broadcast). NVLS status (r02): the kernels pass their 2-GPU tests and the
eager TP=2 engine test; the CUDA-graph TP=2 engine test timed out, so it is never selected by default.
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.distributed as dist

from . import _C


class _RawCudaBuffer:
    """Exposes a raw device pointer to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 3, "strides": None}


class PeerAllReduce:
    FLAG_BYTES = 256

    def __init__(self, device: torch.device, max_bytes: int = 4 << 20):
        lib = _C.load()
        self.lib = lib
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = device
        self.max_bytes = int(max_bytes)
        total = 3 * self.max_bytes + 2 * self.FLAG_BYTES
        self.multicast_ptr = 0       # multicast address of the region (NVLS), 0 = not available
        self._symm = None
        self.forced = os.environ.get("B200_ALLREDUCE", "auto").lower()
        if self.forced not in ("auto", "nccl", "oneshot", "twoshot", "nvls", "nvls_fused", "nvls_one"):
            raise ValueError(f"B200_ALLREDUCE={self.forced!r}: expected "
                             "auto|nccl|oneshot|twoshot|nvls|nvls_fused|nvls_one")
        if self.forced in ("nvls", "nvls_fused", "nvls_one"):
            self._map_symmetric(total)
        else:
            self._map_ipc(total)
        arr = ctypes.c_void_p * self.world
        self._bufs = [arr(*[p + par * self.max_bytes for p in self.peer_ptrs]) for par in (0, 1)]
        self._gather = arr(*[p + 2 * self.max_bytes for p in self.peer_ptrs])
        self._flags = arr(*[p + 3 * self.max_bytes for p in self.peer_ptrs])
        self._state = self.local_ptr + 3 * self.max_bytes + self.FLAG_BYTES
        self._parity = 0
        self.pending = None          # (tensor, parity) written by the last deferred row-parallel GEMM
        dist.barrier()

    def _map_ipc(self, total: int) -> None:
        lib = self.lib
        ptr = ctypes.c_void_p()
        _C.check(lib.b200_comm_alloc(total, ctypes.byref(ptr)), "b200_comm_alloc")
        self.local_ptr = ptr.value
        handle = ctypes.create_string_buffer(64)
        _C.check(lib.b200_comm_get_handle(self.local_ptr, handle), "b200_comm_get_handle")
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw))
        self.peer_ptrs = []
        for r in range(self.world):
            if r == self.rank:
                self.peer_ptrs.append(self.local_ptr)
            else:
                p = ctypes.c_void_p()
                _C.check(lib.b200_comm_open_handle(handles[r], ctypes.byref(p)), "b200_comm_open_handle")
                self.peer_ptrs.append(p.value)
        self.local = torch.as_tensor(_RawCudaBuffer(self.local_ptr, total), device=self.device)

    def _map_symmetric(self, total: int) -> None:
        """Region allocated and exchanged by torch symmetric memory; raises if the platform cannot
        bind it to a multicast object (no silent downgrade: the caller asked for NVLS)."""
        import torch.distributed._symmetric_memory as symm_mem
        self.local = symm_mem.empty(total, dtype=torch.uint8, device=self.device)
        self._symm = symm_mem.rendezvous(self.local, dist.group.WORLD)
        self.local.zero_()
        torch.cuda.synchronize(self.device)
        self.local_ptr = self.local.data_ptr()
        self.peer_ptrs = [int(p) for p in self._symm.buffer_ptrs]
        if self.peer_ptrs[self.rank] != self.local_ptr:
            raise RuntimeError("symmetric memory: local buffer pointer mismatch")
        self.multicast_ptr = int(self._symm.multicast_ptr or 0)
        if not self.multicast_ptr:
            raise RuntimeError("B200_ALLREDUCE asks for NVLS but this platform offers no multicast "
                               "(NVSwitch / driver support missing)")

    # one-shot: every rank pulls (world-1) full copies over NVLink; beyond a few MB of incoming
    # traffic the row-sliced two-shot kernel (reduce a slice, push the sums) moves 2(world-1)/world of a
    # message instead. Measured: Qwen3-0.6B TP=2 (0.5 MB in) +9 % step throughput over NCCL;
    # Qwen3-32B TP=8 (18 MB in per call) one-shot is bandwidth-bound at ~26 us per call.
    ONE_SHOT_MAX_INCOMING = 4 << 20

    def mode(self, rows: int, hidden: int, itemsize: int) -> str | None:
        """"one" / "two" / "nvls" / "nvls_fused" when a fused kernel takes this message, None for NCCL."""
        nbytes = rows * hidden * itemsize
        if self.forced == "nccl" or nbytes > self.max_bytes or hidden % 8 != 0 or hidden > 8192:
            return None
        sliced_ok = rows >= self.world           # the row-sliced forms need a row per rank
        if self.forced == "nvls_one":
            return "nvls_one" if rows <= 148 * 8 else None
        if self.forced in ("nvls", "nvls_fused"):
            return self.forced if sliced_ok else "one"
        if self.forced == "twoshot":
            return "two" if sliced_ok else "one"
        if self.forced == "oneshot" or (self.world - 1) * nbytes <= self.ONE_SHOT_MAX_INCOMING:
            return "one"
        return "two" if sliced_ok else None

    def fits(self, rows: int, hidden: int, itemsize: int) -> bool:
        return self.mode(rows, hidden, itemsize) is not None

    def partial_buffer(self, rows: int, hidden: int, dtype: torch.dtype) -> torch.Tensor:
        """View of this rank's next partial buffer; the GEMM writes its output here."""
        parity = self._parity
        self._parity ^= 1
        n = rows * hidden * dtype.itemsize
        view = self.local[parity * self.max_bytes:parity * self.max_bytes + n].view(dtype).view(rows, hidden)
        self.pending = (view, parity)
        return view

    def allreduce_add_rmsnorm(self, out, residual, weight, eps, mode: str | None = None):
        """out <- RMSNorm(sum_r partial_r + residual) * weight; residual updated in place.
        `out` may be any [rows, hidden] tensor or None (a fresh one is made; the partial buffer itself
        must not be used: peers may still be reading it). `mode` forces a form (tests); every rank
        must pass the same value. In the "nvls_fused" form the result lives in the symmetric region
        (returned view; `out` is ignored) and `residual` is current only in this rank's row slice."""
        view, parity = self.pending
        self.pending = None
        rows, hidden = view.shape
        from .ops import _dt, _stream
        mode = mode or self.mode(rows, hidden, view.dtype.itemsize)
        if mode == "nvls_fused":
            mc = self.multicast_ptr
            n = rows * hidden * view.dtype.itemsize
            _C.check(self.lib.b200_allreduce_add_rmsnorm_nvls_fused(
                residual.data_ptr(), weight.data_ptr(), float(eps), rows, hidden,
                mc + parity * self.max_bytes, mc + 2 * self.max_bytes, self._flags, self._state,
                self.world, self.rank, _dt(view), _stream()), "b200_allreduce_add_rmsnorm_nvls_fused")
            out = self.local[2 * self.max_bytes:2 * self.max_bytes + n].view(view.dtype).view(rows, hidden)
            return out, residual
        if out is None:
            out = torch.empty_like(view)
        if mode == "nvls_one":
            _C.check(self.lib.b200_allreduce_add_rmsnorm_nvls_oneshot(
                out.data_ptr(), residual.data_ptr(), weight.data_ptr(), float(eps), rows, hidden,
                self.multicast_ptr + parity * self.max_bytes, self._flags, self._state, self.world,
                self.rank, _dt(view), _stream()), "b200_allreduce_add_rmsnorm_nvls_oneshot")
        elif mode == "nvls":
            mc = self.multicast_ptr
            _C.check(self.lib.b200_allreduce_add_rmsnorm_nvls(
                out.data_ptr(), residual.data_ptr(), weight.data_ptr(), float(eps), rows, hidden,
                mc + parity * self.max_bytes, mc + 2 * self.max_bytes,
                self.local_ptr + 2 * self.max_bytes, self._flags, self._state, self.world, self.rank,
                _dt(view), _stream()), "b200_allreduce_add_rmsnorm_nvls")
        elif mode == "two":
            _C.check(self.lib.b200_allreduce_add_rmsnorm_twoshot(
                out.data_ptr(), residual.data_ptr(), weight.data_ptr(), float(eps), rows, hidden,
                self._bufs[parity], self._gather, self._flags, self._state, self.world, self.rank,
                _dt(view), _stream()), "b200_allreduce_add_rmsnorm_twoshot")
        else:
            _C.check(self.lib.b200_allreduce_add_rmsnorm(
                out.data_ptr(), residual.data_ptr(), weight.data_ptr(), float(eps), rows, hidden,
                self._bufs[parity], self._flags, self._state, self.world, self.rank, _dt(view),
                _stream()), "b200_allreduce_add_rmsnorm")
        return out, residual

    def close(self):
        if self._symm is not None:        # torch owns the symmetric allocation and its mappings
            self._symm = None
            self.local = None
            return
        for r, p in enumerate(self.peer_ptrs):
            if r != self.rank:
                self.lib.b200_comm_close_handle(p)
        self.local = None
        self.lib.b200_comm_free(self.local_ptr)


_ACTIVE: PeerAllReduce | None = None


def get_comm() -> PeerAllReduce | None:
    return _ACTIVE


def set_comm(comm: PeerAllReduce | None) -> None:
    global _ACTIVE
    _ACTIVE = comm
