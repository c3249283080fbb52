"""Data-parallel context of a model replica: who am I, how many are we, and the hand-written NVLink
exchange (csrc/comm.cu through sn_comm_* / sn_allreduce_sum_*).

Two ways to get one:

  * one process per GPU (``torchrun``): ``Parallel.from_torch_distributed()`` — torch.distributed is
    used for the rendezvous only (exchange of the 64-byte IPC handles, one barrier, the initial
    weight broadcast); every per-step exchange is our own kernel over peer-mapped windows;
  * one process driving G GPUs (``fit(n_gpus=G)``): ``Parallel.init_all(G)`` — device r is rank r.

``SN_COMM=nccl`` keeps the torch.distributed (NCCL) all-reduce for A/B measurements.
"""
import ctypes
import os

from ._lib import lib
from .device import torch


class _CudaView(object):
    """Zero-copy torch view of raw device memory (CUDA array interface)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = dict(shape=(int(n),), typestr=typestr, data=(int(ptr), False), version=2)


class Comm(object):
    """One communicator = one peer-mapped window + its flags (include/shinnosuke_b200.h)."""

    def __init__(self, handle, rank, world, device):
        self.handle, self.rank, self.world, self.device = handle, rank, world, device
        ptr, nbytes = ctypes.c_void_p(), ctypes.c_size_t()
        lib.sn_comm_window(handle, ctypes.byref(ptr), ctypes.byref(nbytes))
        self.window_ptr, self.window_bytes = int(ptr.value or 0), int(nbytes.value)
        self._used = 0

    def carve_f32(self, n):
        """n floats of the window's user region as a torch tensor (128-byte aligned, zero-filled)."""
        off = (self._used + 127) // 128 * 128
        if off + 4 * n > self.window_bytes:
            raise RuntimeError('communicator window exhausted (%d + %d > %d bytes)' % (off, 4 * n, self.window_bytes))
        self._used = off + 4 * n
        t = torch()
        with t.cuda.device(self.device):
            view = t.as_tensor(_CudaView(self.window_ptr + off, n, '<f4'), device=t.device('cuda', self.device))
        view._sn_keepalive = self           # the window must outlive the view
        return view

    def allreduce_f32(self, ptr, n, stream):
        lib.sn_allreduce_sum_f32(self.handle, ptr, int(n), stream)

    def allreduce_f64(self, ptr, n, stream):
        lib.sn_allreduce_sum_f64(self.handle, ptr, int(n), stream)

    def close(self):
        if self.handle is not None:
            lib.sn_comm_destroy(self.handle)
            self.handle = None


def _create_ipc(rank, world, user_bytes, dist):
    t = torch()
    h = ctypes.c_void_p()
    lib.sn_comm_create(rank, world, int(user_bytes), ctypes.byref(h))
    mine = ctypes.create_string_buffer(64)
    lib.sn_comm_ipc_handle(h, mine)
    handles = [None] * world
    dist.all_gather_object(handles, mine.raw)
    lib.sn_comm_connect(h, b''.join(handles))
    dist.barrier()                          # nobody signals a window that a peer has not mapped yet
    return Comm(h, rank, world, t.cuda.current_device())


class Parallel(object):
    """rank / world of this replica plus its communicators.  `grad` carries the gradient arena (in
    place, in its window), `stats` the small float64 BatchNorm sums: two windows, because the arena
    exchange runs on a side stream WHILE the remaining layers (and their BatchNorm exchanges) run
    their backward pass, and one window's flags serve one exchange at a time."""

    def __init__(self, rank, world, grad=None, stats=None, dist=None):
        self.rank, self.world = rank, world
        self.grad, self.stats = grad, stats
        self.dist = dist                    # torch.distributed in torchrun mode (rendezvous / NCCL fallback)

    @property
    def own_kernels(self):
        return self.grad is not None

    @classmethod
    def single(cls):
        return cls(0, 1)

    @classmethod
    def from_torch_distributed(cls, arena_floats):
        t = torch()
        d = t.distributed
        if not (d.is_available() and d.is_initialized()) or d.get_world_size() == 1:
            return cls.single()
        rank, world = d.get_rank(), d.get_world_size()
        if os.environ.get('SN_COMM', '') == 'nccl' or world > 8:
            return cls(rank, world, dist=d)
        grad = _create_ipc(rank, world, 4 * int(arena_floats) + 1024, d)
        stats = _create_ipc(rank, world, 0, d)
        return cls(rank, world, grad, stats, d)

    @classmethod
    def init_all(cls, ndev, arena_floats):
        """One process, ndev devices: a list of contexts, one per device."""
        if ndev == 1:
            return [cls.single()]
        t = torch()
        out = []
        grads = (ctypes.c_void_p * ndev)()
        stats = (ctypes.c_void_p * ndev)()
        lib.sn_comm_init_all(ndev, 4 * int(arena_floats) + 1024, grads)
        lib.sn_comm_init_all(ndev, 0, stats)
        for r in range(ndev):
            out.append(cls(r, ndev, Comm(ctypes.c_void_p(grads[r]), r, ndev, r), Comm(ctypes.c_void_p(stats[r]), r, ndev, r)))
        return out

    # ---- exchanges --------------------------------------------------------
    def allreduce_grads(self, tensor, stream):
        """In-place sum of a float32 torch tensor over the replicas (asynchronous on `stream`)."""
        if self.world == 1:
            return
        if self.grad is not None:
            self.grad.allreduce_f32(tensor.data_ptr(), tensor.numel(), stream)
        else:
            self.dist.all_reduce(tensor)

    def allreduce_stats(self, tensor, stream):
        """In-place sum of a small float64 torch tensor (BatchNorm sums) over the replicas."""
        if self.world == 1:
            return
        if self.stats is not None:
            self.stats.allreduce_f64(tensor.data_ptr(), tensor.numel(), stream)
        else:
            self.dist.all_reduce(tensor)
