"""Several GPUs in ONE process: z-slabs over the C ABI's s2m_multi (csrc/multi.cpp).

The reference drives one device from one thread (main.rs:180-196, :298-356).  MultiContext holds one engine
context and one host thread per GPU and a NCCL communicator per GPU (ncclCommInitAll inside the library);
mesh_run cuts the grid into cost-balanced z-slabs, meshes them concurrently, exchanges the per-slab vertex
counts with one ncclAllGather and returns the slab results in z order.  (One process per GPU over
torch.distributed is the other multi-GPU form: sdf2mesh_b200/distributed.py.)
"""
import ctypes
from typing import List, Sequence

import numpy as np

from . import _capi
from ._capi import MeshParams, MultiTimings, check, lib
from .engine import MeshResult, Module


class _BorrowedContext:
    """a device context owned by the MultiContext (never destroyed through this handle)"""

    def __init__(self, handle, device, owner):
        self._h, self.device, self._owner = handle, device, owner

    def close(self):
        pass


class MultiContext:
    def __init__(self, devices: Sequence[int], flags: int = 0):
        self._h = ctypes.c_void_p()
        arr = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
        check(lib().s2m_multi_create(arr, len(devices), flags, ctypes.byref(self._h)))
        self.devices = [int(d) for d in devices]

    def close(self):
        if getattr(self, "_h", None):
            lib().s2m_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return lib().s2m_multi_size(self._h)

    def ctx(self, k: int) -> _BorrowedContext:
        return _BorrowedContext(ctypes.c_void_p(lib().s2m_multi_ctx(self._h, k)), self.devices[k], self)

    @property
    def nccl_version(self) -> int:
        """NCCL version code if the counts travel through ncclAllGather, 0 if through host memory"""
        v = ctypes.c_int()
        return v.value if lib().s2m_multi_uses_nccl(self._h, ctypes.byref(v)) else 0

    def mesh_run(self, module: Module, params: MeshParams) -> List[MeshResult]:
        """module: compiled with ctx=None (Sdf3DShader.create_shader_module(None)) or for any context"""
        n = len(self)
        out = (ctypes.c_void_p * n)()
        check(lib().s2m_multi_mesh_run(self._h, module._h, ctypes.byref(params), out))
        return [MeshResult(ctypes.c_void_p(out[k]), self) for k in range(n)]

    def partition(self) -> List[int]:
        b = (ctypes.c_uint32 * (len(self) + 1))()
        check(lib().s2m_multi_get_partition(self._h, b))
        return [int(x) for x in b]

    def timings(self) -> dict:
        t = MultiTimings()
        check(lib().s2m_multi_last_timings(self._h, ctypes.byref(t)))
        n = t.n
        return {"wall_ms": t.wall_ms, "begin_ms": list(t.begin_ms[:n]), "exchange_ms": list(t.exchange_ms[:n]), "finish_ms": list(t.finish_ms[:n])}


def partition_slices(n_slices: int, world: int, cost=None) -> List[int]:
    """s2m_partition_slices: the C ABI's slab boundaries (same contract as distributed.partition_slices)"""
    out = (ctypes.c_uint32 * (world + 1))()
    c = None if cost is None else np.ascontiguousarray(cost, np.float64)
    check(lib().s2m_partition_slices(n_slices, world, None if c is None else c.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                     0 if c is None else len(c), out))
    return [int(x) for x in out]


def rebalance_slices(bounds, seconds, cost=None) -> List[int]:
    world = len(bounds) - 1
    b = (ctypes.c_uint32 * (world + 1))(*[int(x) for x in bounds])
    out = (ctypes.c_uint32 * (world + 1))()
    s = np.ascontiguousarray(seconds, np.float64)
    c = None if cost is None else np.ascontiguousarray(cost, np.float64)
    check(lib().s2m_rebalance_slices(b, world, s.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                     None if c is None else c.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 0 if c is None else len(c), out))
    return [int(x) for x in out]
