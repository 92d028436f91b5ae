"""ctypes binding of tiramisu_b200/lib/libtiramisu_b200.so (the C ABI in include/tiramisu_b200.h).

There is no CPU or library fallback: if the CUDA extension is missing, importing the product API
fails here, loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtiramisu_b200.so")


class HalideDimension(C.Structure):
    _fields_ = [("min", C.c_int32), ("extent", C.c_int32), ("stride", C.c_int32), ("flags", C.c_uint32)]


class HalideType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class HalideBuffer(C.Structure):
    """halide_buffer_t (include/tiramisu_b200/halide_buffer.h; 56 bytes)."""
    _fields_ = [("device", C.c_uint64), ("device_interface", C.c_void_p), ("host", C.c_void_p),
                ("flags", C.c_uint64), ("type", HalideType), ("dimensions", C.c_int32),
                ("dim", C.POINTER(HalideDimension)), ("padding", C.c_void_p)]


assert C.sizeof(HalideBuffer) == 56 and C.sizeof(HalideDimension) == 16

_lib = None


class Halo(C.Structure):
    """tiramisu_b200_halo_t (include/tiramisu_b200.h): where a row-sharded stencil finds the rows its
    neighbours own, and the optional sweep handshake."""
    _fields_ = [("up", C.c_void_p), ("dn", C.c_void_p), ("lo", C.c_int32), ("hi", C.c_int32),
                ("flags", C.c_void_p), ("sig_up", C.c_void_p), ("sig_dn", C.c_void_p),
                ("wait_value", C.c_int32), ("signal_value", C.c_int32)]


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m tiramisu_b200.build` "
            "(there is no CPU fallback for the sm_100a backend)")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    P = C.POINTER(HalideBuffer)
    u64, i32, f32, f64, vp, b = C.c_uint64, C.c_int32, C.c_float, C.c_double, C.c_void_p, C.c_bool

    def sig(name, restype, *argtypes):
        fn = getattr(L, name)
        fn.restype = restype
        fn.argtypes = list(argtypes)

    sig("tiramisu_b200_version", C.c_char_p)
    sig("tiramisu_b200_last_error", C.c_char_p)
    sig("tiramisu_b200_set_error_mode", C.c_int, C.c_int)
    sig("tiramisu_b200_set_stream", None, vp)
    sig("tiramisu_b200_get_stream", vp)
    sig("tiramisu_b200_kernel_launches", u64)
    sig("tiramisu_b200_pool_trim", C.c_int)
    sig("tiramisu_b200_null_launch", C.c_int, i32, i32, i32, i32)
    sig("tiramisu_b200_set_gemm_mode", C.c_int, C.c_int)
    sig("tiramisu_b200_set_gemm_kernel", C.c_int, C.c_int)
    sig("tiramisu_cuda_memcpy_to_device", C.c_int, vp, vp, u64)
    sig("tiramisu_cuda_memcpy_to_host", C.c_int, vp, vp, u64)
    sig("tiramisu_cuda_memcpy_to_symbol", C.c_int, vp, vp, u64)
    sig("tiramisu_cuda_memcpy_to_device_async", C.c_int, vp, vp, u64)
    sig("tiramisu_cuda_memcpy_to_host_async", C.c_int, vp, vp, u64)
    sig("tiramisu_cuda_malloc", vp, u64)
    sig("tiramisu_cuda_free", C.c_int, vp)
    sig("tiramisu_cuda_stream_synchronize", i32, i32)
    sig("tiramisu_cublas_sgemm", C.c_int, vp, vp, vp, u64, u64, u64, f32, f32, u64, u64, u64, u64, u64, u64, b, b)
    sig("tiramisu_cublas_dgemm", C.c_int, vp, vp, vp, u64, u64, u64, f64, f64, u64, u64, u64, u64, u64, u64, b, b)
    for t in ("int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64", "float32", "float64"):
        sig(f"tiramisu_address_of_{t}", vp, P, C.c_ulong)
    sig("tiramisu_b200_heat2d", C.c_int, vp, vp, i32, i32, i32, i32, i32)
    sig("tiramisu_b200_blur_f32", C.c_int, vp, vp, i32, i32, i32, i32)
    sig("tiramisu_b200_blurxy_u8", C.c_int, vp, vp, i32, i32, i32)
    sig("tiramisu_b200_boxblur_u32", C.c_int, vp, vp, i32, i32)
    sig("tiramisu_b200_matmul_u8", C.c_int, vp, vp, vp, i32)
    sig("tiramisu_b200_fill_i32", C.c_int, vp, u64, i32)
    sig("tiramisu_b200_transpose_f32", C.c_int, vp, vp, i32, i32)
    sig("tiramisu_b200_conv_relu_maxpool", C.c_int, vp, vp, vp, vp, i32, i32, i32, i32)
    sig("tiramisu_b200_set_conv_mode", C.c_int, C.c_int)
    sig("tiramisu_b200_set_conv_kernel", C.c_int, C.c_int)
    sig("tiramisu_b200_lstm", C.c_int, vp, vp, vp, vp, i32, i32, i32, i32)
    sig("tiramisu_b200_set_lstm_mode", C.c_int, C.c_int)
    sig("tiramisu_b200_set_lstm_kernel", C.c_int, C.c_int)
    sig("tiramisu_b200_lstm_last_engine", C.c_char_p)
    sig("tiramisu_b200_lstm_ex", C.c_int, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32)
    sig("tiramisu_b200_comm_unique_id", C.c_int, vp)
    sig("tiramisu_b200_comm_init", C.c_int, i32, i32, vp)
    sig("tiramisu_b200_comm_destroy", C.c_int)
    sig("tiramisu_b200_comm_rank", C.c_int)
    sig("tiramisu_b200_comm_size", C.c_int)
    sig("tiramisu_b200_comm_broadcast", C.c_int, vp, u64, i32)
    sig("tiramisu_b200_comm_broadcast_large", C.c_int, vp, u64, i32)
    sig("tiramisu_b200_comm_exchange", C.c_int, i32, C.POINTER(i32), C.POINTER(vp), C.POINTER(u64),
        C.POINTER(i32), C.POINTER(vp), C.POINTER(u64))
    sig("tiramisu_b200_comm_alloc_symmetric", C.c_int, u64, C.POINTER(vp), C.POINTER(vp))
    sig("tiramisu_b200_comm_free_symmetric", C.c_int, vp)
    sig("tiramisu_b200_comm_broadcast_sym", C.c_int, C.POINTER(vp), u64, u64, u64, i32, C.c_uint32, i32)
    sig("tiramisu_b200_comm_broadcast_trace_dump", C.c_int)
    sig("tiramisu_b200_sgemm_bcast_b", C.c_int, C.POINTER(vp), u64, u64, i32, C.c_uint32, i32, vp, u64, vp, vp, u64, u64, u64,
        C.c_float, C.c_float, u64, u64)
    sig("tiramisu_b200_heat2d_peer", C.c_int, vp, vp, i32, i32, i32, i32, i32, C.POINTER(Halo))
    sig("tiramisu_b200_blur_f32_peer", C.c_int, vp, vp, i32, i32, i32, i32, C.POINTER(Halo))
    sig("tiramisu_b200_post_launch", C.c_int, C.c_char_p)
    sig("tiramisu_MPI_init", C.c_int)
    sig("tiramisu_MPI_cleanup", None)
    sig("tiramisu_MPI_global_barrier", None)
    sig("tiramisu_MPI_Comm_rank", C.c_int, C.c_int)
    for t in ("int8", "int16", "int32", "int64", "uint8", "uint16", "uint32", "uint64", "f32", "f64"):
        for op in ("Send", "Ssend", "Recv"):
            sig(f"tiramisu_MPI_{op}_{t}", None, C.c_int, C.c_int, C.c_int, vp)
        for op in ("Isend", "Issend", "Irecv"):
            sig(f"tiramisu_MPI_{op}_{t}", None, C.c_int, C.c_int, C.c_int, vp, vp)
    sig("tiramisu_b200_set_heat2d_dist", C.c_int, i32, i32, i32, i32)
    for name, n in (("tiramisu_b200_lstm_gpu_lib", 6), ("tiramisu_b200_lstm_cpu_lib", 6), ("heat2d_dist_tiramisu", 2), ("conv_relu_maxpool_block", 4), ("lstm", 5), ("heat2d_tiramisu", 2), ("heat2dgpu_tiramisu", 3), ("blurxy_tiramisu", 2),
                    ("blurxy_f32_tiramisu", 2), ("boxblur", 2), ("blurautodist_tiramisu", 2),
                    ("sgemm_tiramisu", 6), ("matmul", 6), ("test_162", 5), ("test_164", 6), ("test_167", 3),
                    ("test_168", 2), ("test_197", 2)):
        sig(name, C.c_int, *([P] * n))
    # Python callers want exceptions, not exit(1) (see tiramisu_b200_set_error_mode).
    L.tiramisu_b200_set_error_mode(1)
    _lib = L
    return L


class TiramisuError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().tiramisu_b200_last_error().decode() or "unknown error"
        raise TiramisuError(f"{what}: {msg} (rc={rc})" if what else f"{msg} (rc={rc})")
