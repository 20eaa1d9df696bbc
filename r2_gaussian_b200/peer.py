"""NVLink peer-memory exchange for the Gaussian-sharded projector: `PeerReducer` sums one float32 buffer per rank
into every rank's output with a single kernel launch (r2x_peer_allreduce_sum) instead of an NCCL all-reduce.

One process per GPU of ONE node (torchrun); `torch.distributed` is only used at construction, to swap the CUDA IPC
handles of the buffers.  The sum is taken in rank order: all ranks hold bitwise identical results.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from ._lib import check, load

MAX_PEERS = 16


class _DevArray:
    """Zero-copy view of raw device memory for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr: int, n: int, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class PeerReducer:
    def __init__(self, numel: int, device, group=None, timeout_s: float = 20.0):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerReducer needs an initialised torch.distributed process group")
        self.lib = load()
        self.device = torch.device(device)
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > MAX_PEERS:
            raise RuntimeError(f"PeerReducer supports at most {MAX_PEERS} ranks")
        self.numel = int(numel)
        self.timeout_cycles = int(max(timeout_s, 0.001) * 1.9e9)   # SM clock cycles the kernel waits for a peer
        self.stride = (self.numel * 4 + 255) // 256 * 256
        nbytes = 2 * self.stride + 256
        with torch.cuda.device(self.device):
            base = C.c_void_p()
            check(self.lib.r2x_peer_alloc(nbytes, C.byref(base)), "r2x_peer_alloc")
            self._base = base.value
            handle = (C.c_ubyte * 64)()
            check(self.lib.r2x_ipc_export(self._base, handle), "r2x_ipc_export")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            self._opened, bases = [], []
            for p, h in enumerate(handles):
                if p == self.rank:
                    bases.append(self._base)
                    continue
                ptr = C.c_void_p()
                buf = (C.c_ubyte * 64).from_buffer_copy(h)
                check(self.lib.r2x_ipc_open(buf, C.byref(ptr)), "r2x_ipc_open")
                self._opened.append(ptr.value)
                bases.append(ptr.value)
            self._bufs = [(C.c_void_p * self.world)(*[b + k * self.stride for b in bases]) for k in (0, 1)]
            self._flags = (C.c_void_p * self.world)(*[b + 2 * self.stride for b in bases])
            self._partials = [torch.as_tensor(_DevArray(self._base + k * self.stride, self.numel), device=self.device)
                              for k in (0, 1)]
            self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
            torch.cuda.synchronize(self.device)
        dist.barrier(group=group)
        self.epoch = 0

    def partial(self) -> torch.Tensor:
        """The buffer this rank fills for the NEXT reduce() (flat float32[numel]; reshape as needed)."""
        return self._partials[(self.epoch + 1) & 1]

    def reduce(self, out: torch.Tensor) -> torch.Tensor:
        """out[...] = sum over ranks of their partial(); enqueued on the current stream, no host sync."""
        if out.numel() != self.numel or out.dtype != torch.float32 or not out.is_contiguous():
            raise RuntimeError("PeerReducer.reduce: out must be a contiguous float32 tensor of the reducer's size")
        self.epoch += 1
        rc = self.lib.r2x_peer_allreduce_sum_t(torch.cuda.current_stream(self.device).cuda_stream, self.world, self.rank,
                                               self._bufs[self.epoch & 1], self._flags, self.epoch & 0xFFFFFFFF,
                                               out.data_ptr(), self.numel, self.status.data_ptr(), self.timeout_cycles)
        check(rc, "r2x_peer_allreduce_sum")
        return out

    def ok(self) -> bool:
        """Synchronises; False if a peer failed to arrive within the kernel's time-out."""
        return int(self.status.item()) == 0

    def close(self):
        if getattr(self, "_base", None) is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        self._partials = []
        for p in self._opened:
            self.lib.r2x_ipc_close(p)
        self.lib.r2x_peer_free(self._base)
        self._base = None
