"""Page-locked host staging tensors backed by cp360_host_alloc (include/cp360.h).

    frames = pinned_empty((B, 960, 1920, 3), torch.uint8)            # cudaHostAlloc, like .pin_memory()
    frames = pinned_empty(shape, torch.uint8, mode="hugepage")       # THP-backed + cudaHostRegister

The allocation is freed when the last tensor viewing it dies. ``tensor.is_pinned()`` is True, so
``dev.copy_(tensor, non_blocking=True)`` is a true asynchronous DMA.
"""
import ctypes
import weakref

import numpy as np
import torch

from . import _lib

MODES = {"pinned": 0, "write_combined": 1, "hugepage": 2}


def _release(ptr):
    try:
        _lib.lib().cp360_host_free(ptr)
    except Exception:                                       # noqa: BLE001 - interpreter shutdown
        pass


def pinned_empty(shape, dtype=torch.uint8, mode="pinned"):
    """Uninitialised page-locked host tensor of `shape` / `dtype`; mode: 'pinned' | 'write_combined' | 'hugepage'."""
    if mode not in MODES:
        raise ValueError("mode must be one of %s" % sorted(MODES))
    shape = tuple(int(v) for v in shape)
    itemsize = torch.empty((), dtype=dtype).element_size()
    n = int(np.prod(shape, dtype=np.int64)) if shape else 1
    if n == 0:
        return torch.empty(shape, dtype=dtype)
    nbytes = n * itemsize
    ptr = ctypes.c_void_p()
    _lib.check(_lib.lib().cp360_host_alloc(nbytes, MODES[mode], ctypes.byref(ptr)))
    buf = (ctypes.c_uint8 * nbytes).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=np.uint8, count=nbytes)
    # torch.from_numpy keeps `arr` alive for as long as the storage (and so any view of it) exists;
    # the allocation is released when the array is collected
    weakref.finalize(arr, _release, ptr.value)
    t = torch.from_numpy(arr)
    return t.view(dtype)[:n].reshape(shape)
