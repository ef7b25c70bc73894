"""Row sharding of one image across ranks (one process per GPU) and the small all-reduce adapter.

The N-scaled work of the filter is a map over pixels followed by small reductions, so rank g owns a
contiguous slab of image rows and only p-vectors (Sinkhorn sums), one p x p Gram matrix and the
k-vector V^T z ever cross NVLink.  The C ABI takes the reduction as a callback
(nle_b200_allreduce_fn); this module builds it from torch.distributed."""
from __future__ import annotations

import ctypes as C

import numpy as np


def row_slab(rows: int, rank: int, world: int):
    """Contiguous, balanced row range [row0, row1) of rank `rank` out of `world`."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    if world > rows:
        raise ValueError("more ranks than image rows")
    base, rem = divmod(rows, world)
    row0 = rank * base + min(rank, rem)
    return row0, row0 + base + (1 if rank < rem else 0)


class _DevArray:
    """Zero-copy view of `count` doubles at a raw CUDA device pointer (for torch.as_tensor)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (int(ptr), False), "version": 3}


def torch_allreduce(device=None, group=None):
    """Returns f(ptr, count, stream) that sums `count` doubles in place across the process group.

    device=None -> the buffer is HOST memory (gloo, used by the CPU tests of the sharding logic);
    otherwise a torch.device for NCCL over NVLink."""
    import torch
    import torch.distributed as dist

    def _fn(ptr, count, stream=None):
        if device is None:
            arr = np.ctypeslib.as_array(C.cast(C.c_void_p(int(ptr)), C.POINTER(C.c_double)), shape=(count,))
            t = torch.from_numpy(arr)
        else:
            t = torch.as_tensor(_DevArray(ptr, count), device=device)
        dist.all_reduce(t, group=group)
    return _fn


class LibraryComm:
    """The library's own communicator (csrc/nccl_comm.cu): the all-reduces of a sharded training call are enqueued by
    libnle_b200.so itself -- no Python between a kernel and its collective.  The p x p Gram is an ncclAllReduce; the
    latency-sized p- and k-vectors are single launches of the library's peer_allreduce_kernel over NVLink peer memory
    (CUDA IPC inboxes), with ncclAllReduce as the fallback where peers cannot map each other's memory (`info()`).

    torch.distributed is only the bootstrap: rank 0 draws the 128-byte NCCL unique id and it is broadcast once over the
    already initialised process group.  `callback` / `user` are what the sharded entry points of the C ABI take
    (nle_b200_train_bgr_u8, nle_b200_train_u8_dev, ...)."""

    def __init__(self, device):
        import torch
        import torch.distributed as dist
        from . import _lib
        self._lib = _lib.load()
        rank, world = dist.get_rank(), dist.get_world_size()
        ident = (C.c_ubyte * 128)()
        if rank == 0:
            _lib.check(self._lib.nle_b200_comm_unique_id(ident))
        t = torch.tensor(list(ident), dtype=torch.uint8, device=device)
        dist.broadcast(t, src=0)
        ident = (C.c_ubyte * 128)(*t.cpu().tolist())
        h = C.c_void_p()
        _lib.check(self._lib.nle_b200_comm_create(ident, rank, world, C.byref(h)))
        self.handle = h
        self.callback = C.cast(self._lib.nle_b200_comm_allreduce, _lib.ALLREDUCE_FN)
        self.user = h

    def info(self):
        """Which path the latency-sized all-reduces take: {'peer_path': bool, 'why': str, 'peer_calls': n, 'nccl_calls': n}."""
        peer = C.c_int(0)
        calls = (C.c_ulonglong * 2)()
        why = self._lib.nle_b200_comm_info(self.handle, C.byref(peer), calls)
        return {"peer_path": bool(peer.value), "why": why.decode() if why else "", "peer_calls": int(calls[0]),
                "nccl_calls": int(calls[1])}

    def close(self):
        if self.handle:
            self._lib.nle_b200_comm_destroy(self.handle)
            self.handle = None
