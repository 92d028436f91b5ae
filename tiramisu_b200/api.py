"""Host-side mirror of the reference's generated-function interface (B1 in include/tiramisu_b200.h).

Every function has the name and argument order of the `extern "C"` symbol the reference's wrapper
programs call, takes numpy arrays / torch tensors where the C++ wrappers take Halide::Buffer, and
calls straight through the C ABI.  Host arrays are copied to the GPU and back inside the call (as
the reference's generated code does); CUDA tensors are used in place.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .halide import Buffer


def _call(name, *objs):
    bufs = [o if isinstance(o, Buffer) else Buffer(o) for o in objs]
    rc = getattr(_lib.lib(), name)(*[b.ptr() for b in bufs])
    _lib.check(rc, name)


# ---- stencils -------------------------------------------------------------------------------------
def heat2d_tiramisu(inp, out):
    """benchmarks/halide/wrapper_heat2d.h:10 — in,out f32[H][W]."""
    _call("heat2d_tiramisu", inp, out)


def heat2dgpu_tiramisu(sizes, inp, out):
    """benchmarks/halide/wrapper_heat2dgpu.cpp:44 — sizes i32[2] = {W, H}."""
    _call("heat2dgpu_tiramisu", sizes, inp, out)


def heat2d_dist_tiramisu(buff_out_even, buff_out_odd, rows_per_rank=None, iters=None, protocol=None):
    """benchmarks/halide/wrapper_heat2d_dist.cpp:69 — one call per rank under a communicator
    (tiramisu_MPI_init / dist.comm_init_from_torch): host arrays f32[rows+2][M], rows 1..rows owned.
    Like the reference's wrapper, only the host pointers are passed (no extents in the halide_buffer_t);
    rows_per_rank / iters / protocol go through tiramisu_b200_set_heat2d_dist (cols = the arrays' width)."""
    from .halide import raw_host_only
    if rows_per_rank is not None or iters is not None or protocol is not None:
        _lib.check(_lib.lib().tiramisu_b200_set_heat2d_dist(rows_per_rank or 0, buff_out_even.shape[-1], iters or 0,
                                                            protocol or 0), "tiramisu_b200_set_heat2d_dist")
    a, b = raw_host_only(buff_out_even.ctypes.data), raw_host_only(buff_out_odd.ctypes.data)
    _lib.check(_lib.lib().heat2d_dist_tiramisu(C.byref(a), C.byref(b)), "heat2d_dist_tiramisu")


def blurxy_tiramisu(inp, out):
    """benchmarks/halide/wrapper_blurxy.cpp:21 — in u8[3][H][W], out u8[3][H-8][W-8]."""
    _call("blurxy_tiramisu", inp, out)


def blurxy_f32_tiramisu(inp, out):
    """fp32 blur of BASELINE config 2 — in f32[H][W], out f32[H-2][W-2]."""
    _call("blurxy_f32_tiramisu", inp, out)


def boxblur(inp, out):
    """tests/wrapper_test_160.h — in u32[rows+2][cols+2], out u32[rows][cols]."""
    _call("boxblur", inp, out)


def blurautodist_tiramisu(inp, out):
    _call("blurautodist_tiramisu", inp, out)


# ---- GEMMs ----------------------------------------------------------------------------------------
def sgemm_tiramisu(SIZES, alpha, beta, A, B, Cm):
    """sgemm/cpu/sgemm_wrapper.h:13 — SIZES i32[3] = {N, M, K}; C[NxM] = a*A[NxK]*B[KxM] + b*C."""
    _call("sgemm_tiramisu", SIZES, alpha, beta, A, B, Cm)


def matmul(Consts, A, B, Cm, time_start=None, time_end=None):
    """sgemm/gpu/wrapper.h:6-11 — Consts f32[2] = {alpha, beta}."""
    t0 = time_start if time_start is not None else np.zeros(1, np.float32)
    t1 = time_end if time_end is not None else np.zeros(1, np.float32)
    _call("matmul", Consts, A, B, Cm, t0, t1)


def test_162(sizes, params, A, B, Cm):
    _call("test_162", sizes, params, A, B, Cm)


def test_164(sizes, params, transposes, A, B, Cm):
    _call("test_164", sizes, params, transposes, A, B, Cm)


def test_167(A, B, Cm):
    _call("test_167", A, B, Cm)


def test_168(A, B):
    _call("test_168", A, B)


def test_197(A, B):
    _call("test_197", A, B)


# ---- DNN blocks -------------------------------------------------------------------------------------
def conv_relu_maxpool_block(inp, filt, bias, out):
    """Conv-ReLU-MaxPool/cpu/dense/wrapper_nn_block.cpp:43-48 — input f32[B][N+2][N+2][FIn],
    filter f32[FOut/8][3][3][FIn][8], bias f32[FOut], output f32[B][FOut/8][N/2][N/2][8]."""
    _call("conv_relu_maxpool_block", inp, filt, bias, out)


def lstm(params, Weights, biases, x, y):
    """LSTM/cpu/wrapper.h:6-10 — params i32[4] = {F, B, L, T}; Weights f32[L][2][4F][F]."""
    _call("lstm", params, Weights, biases, x, y)


def lstm_gpu_lib(Weights, biases, x, y, time_start=None, time_end=None):
    """LSTM/gpu_lib/wrapper.cpp:74-79 — Weights f32[L][2][4F][F] with index 0 = W, 1 = R, gate rows i, f, z, o."""
    t0 = time_start if time_start is not None else np.zeros(1, np.float32)
    t1 = time_end if time_end is not None else np.zeros(1, np.float32)
    _call("tiramisu_b200_lstm_gpu_lib", Weights, biases, x, y, t0, t1)


def lstm_cpu_lib(weights, biases, x, h, c, out):
    """LSTM/cpu_lib/wrapper.cpp:65-72 — weights f32[L][2][F][4F] (transposed), h f32[T+1][B][F], c f32[B][F]."""
    _call("tiramisu_b200_lstm_cpu_lib", weights, biases, x, h, c, out)


def set_conv_mode(mode: int) -> int:
    return int(_lib.lib().tiramisu_b200_set_conv_mode(mode))


def set_conv_kernel(kernel: int) -> int:
    """0 auto, 1 = 8x16 tile (ky taps share a TMA tile), 2 = 16x8 tile (all nine taps share one),
    3 = as 2 on CTA pairs (cta_group::2)."""
    return int(_lib.lib().tiramisu_b200_set_conv_kernel(kernel))


def set_lstm_mode(mode: int) -> int:
    return int(_lib.lib().tiramisu_b200_set_lstm_mode(mode))


def set_lstm_kernel(kernel: int) -> int:
    """0 auto, 1 one launch per cell, 2 persistent layer kernel."""
    return int(_lib.lib().tiramisu_b200_set_lstm_kernel(kernel))


# ---- B2 runtime, for callers that manage device memory themselves ------------------------------------
def set_stream(cuda_stream: int | None) -> None:
    """Launch on this cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); None = per-thread default."""
    _lib.lib().tiramisu_b200_set_stream(C.c_void_p(cuda_stream or 0))


def synchronize() -> None:
    _lib.check(_lib.lib().tiramisu_cuda_stream_synchronize(0), "tiramisu_cuda_stream_synchronize")


def kernel_launches() -> int:
    return int(_lib.lib().tiramisu_b200_kernel_launches())


def set_gemm_mode(mode: int) -> int:
    """0 auto, 1 exact fp32 SIMT, 2 force tcgen05 TF32."""
    return int(_lib.lib().tiramisu_b200_set_gemm_mode(mode))


def set_gemm_kernel(kernel: int) -> int:
    """0 auto, 1 one-tile-per-CTA, 2 persistent 128x256, 3 persistent CTA pairs (cta_group::2), 4 split-K 2-CTA clusters."""
    return int(_lib.lib().tiramisu_b200_set_gemm_kernel(kernel))


def cublas_sgemm(A_ptr, B_ptr, C_ptr, M, N, K, alpha, beta, ldA=0, ldB=0, ldC=0, offA=0, offB=0, offC=0,
                 tA=False, tB=False):
    """tiramisu_cublas_sgemm on raw device pointers (src/tiramisu_cuda_wrappers.cpp:74)."""
    rc = _lib.lib().tiramisu_cublas_sgemm(A_ptr, B_ptr, C_ptr, M, N, K, alpha, beta, ldA, ldB, ldC, offA, offB,
                                          offC, tA, tB)
    _lib.check(rc, "tiramisu_cublas_sgemm")


def cublas_dgemm(A_ptr, B_ptr, C_ptr, M, N, K, alpha, beta, ldA=0, ldB=0, ldC=0, offA=0, offB=0, offC=0,
                 tA=False, tB=False):
    rc = _lib.lib().tiramisu_cublas_dgemm(A_ptr, B_ptr, C_ptr, M, N, K, alpha, beta, ldA, ldB, ldC, offA, offB,
                                          offC, tA, tB)
    _lib.check(rc, "tiramisu_cublas_dgemm")
