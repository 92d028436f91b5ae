"""halide_buffer_t construction for Python callers — the role Halide::Buffer<T>::raw_buffer()
plays in the reference's wrapper_*.cpp harnesses.

A C-contiguous array of shape (d_{n-1}, ..., d_0) becomes a buffer whose Halide dimensions are
innermost-first (dim[0] = d_0 with stride 1), which is the "indices are flipped" convention of
benchmarks/linear_algebra/blas/level3/sgemm/gpu/wrapper.cpp:59-62.  numpy arrays and CPU torch
tensors are host buffers (->host); CUDA torch tensors are device-resident (->device, ->host NULL).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import HalideBuffer, HalideDimension

_NP_CODES = {"i": 0, "u": 1, "f": 2, "b": 1}


class Buffer:
    """Owns a halide_buffer_t plus whatever keeps its memory alive."""

    def __init__(self, obj):
        self.obj = obj
        shape, itemsize, code, ptr, on_device = _describe(obj)
        n = len(shape)
        self._dims = (HalideDimension * max(n, 1))()
        stride = 1
        for i, ext in enumerate(reversed(shape)):
            self._dims[i].min = 0
            self._dims[i].extent = int(ext)
            self._dims[i].stride = int(stride)
            self._dims[i].flags = 0
            stride *= int(ext)
        self.raw = HalideBuffer()
        self.raw.device = ptr if on_device else 0
        self.raw.device_interface = None
        self.raw.host = None if on_device else ptr
        self.raw.flags = 0
        self.raw.type.code = code
        self.raw.type.bits = itemsize * 8
        self.raw.type.lanes = 1
        self.raw.dimensions = n
        self.raw.dim = C.cast(self._dims, C.POINTER(HalideDimension))
        self.raw.padding = None
        self.on_device = on_device

    def ptr(self):
        return C.byref(self.raw)


def _describe(obj):
    if isinstance(obj, np.ndarray):
        if not obj.flags["C_CONTIGUOUS"]:
            raise ValueError("halide buffers must be C-contiguous")
        return obj.shape, obj.dtype.itemsize, _NP_CODES[obj.dtype.kind], obj.ctypes.data, False
    try:
        import torch
    except ImportError:  # pragma: no cover
        torch = None
    if torch is not None and isinstance(obj, torch.Tensor):
        if not obj.is_contiguous():
            raise ValueError("halide buffers must be contiguous")
        if obj.dtype.is_floating_point:
            code = 2
        elif obj.dtype in (torch.uint8, torch.bool) or "uint" in str(obj.dtype):
            code = 1
        else:
            code = 0
        return tuple(obj.shape), obj.element_size(), code, obj.data_ptr(), obj.is_cuda
    raise TypeError(f"cannot make a halide_buffer_t from {type(obj)!r}")


def raw_host_only(ptr: int) -> HalideBuffer:
    """A halide_buffer_t with only .host set and zeroed metadata, as
    benchmarks/halide/wrapper_heat2d_dist.cpp:52-66 passes."""
    b = HalideBuffer()
    b.host = ptr
    return b
