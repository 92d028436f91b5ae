"""ctypes/numpy front for oracle/libtiramisu_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module (see oracle/tiramisu_oracle.c header).  The product
package tiramisu_b200/ never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtiramisu_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "tiramisu_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"], "oracle wants C-contiguous arrays"
    return a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def host_cores() -> int:
    """Cores this process may run on (the GPU box's host cores as the scheduler gives them to us)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def set_num_threads(n: int | None = None) -> int:
    """OpenMP threads for every oracle entry point; None = all host cores.  torchrun exports
    OMP_NUM_THREADS=1, which must not reach a CPU baseline that claims 'all cores'."""
    n = host_cores() if n is None else int(n)
    lib().oracle_set_num_threads(C.c_int(n))
    return num_threads()


def heat2d(inp: np.ndarray, form: str = "cpu", out: np.ndarray | None = None) -> np.ndarray:
    """form: 'cpu' = heat2d_tiramisu.cpp:68, 'gpu' = heat2dgpu_tiramisu.cpp:54.  `out`: a preallocated
    (already touched) result array, so that a timed call does not pay page faults for a fresh one."""
    inp = np.ascontiguousarray(inp, dtype=np.float32)
    H, W = inp.shape
    if out is None:
        out = np.empty_like(inp)
    assert out.shape == inp.shape and out.dtype == np.float32 and out.flags["C_CONTIGUOUS"]
    fn = {"cpu": lib().oracle_heat2d, "gpu": lib().oracle_heat2dgpu}[form]
    fn(_p(inp), _p(out), C.c_int(H), C.c_int(W))
    return out


def heat2d_step(inp: np.ndarray, out: np.ndarray) -> np.ndarray:
    """One Jacobi half-step of heat2d_dist (interior only; borders of `out` untouched)."""
    H, W = inp.shape
    lib().oracle_heat2d_step(_p(inp), _p(out), C.c_int(H), C.c_int(W))
    return out


def blurxy_u8(inp: np.ndarray) -> np.ndarray:
    inp = np.ascontiguousarray(inp, dtype=np.uint8)
    Cc, H, W = inp.shape
    out = np.empty((Cc, H - 8, W - 8), dtype=np.uint8)
    lib().oracle_blurxy_u8(_p(inp), _p(out), C.c_int(Cc), C.c_int(H), C.c_int(W))
    return out


def boxblur_u32(inp: np.ndarray) -> np.ndarray:
    """inp: u32[rows+2][cols+2] -> u32[rows][cols] (vertical pass first; test_160)."""
    inp = np.ascontiguousarray(inp, dtype=np.uint32)
    rows, cols = inp.shape[0] - 2, inp.shape[1] - 2
    out = np.empty((rows, cols), dtype=np.uint32)
    lib().oracle_boxblur_u32(_p(inp), _p(out), C.c_int(rows), C.c_int(cols))
    return out


def blur_f32(inp: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
    inp = np.ascontiguousarray(inp, dtype=np.float32)
    H, W = inp.shape
    if out is None:
        out = np.empty((H - 2, W - 2), dtype=np.float32)
    assert out.shape == (H - 2, W - 2) and out.dtype == np.float32 and out.flags["C_CONTIGUOUS"]
    lib().oracle_blur_f32(_p(inp), _p(out), C.c_int(H), C.c_int(W))
    return out


def sgemm(a: float, b: float, A: np.ndarray, B: np.ndarray, Cm: np.ndarray) -> np.ndarray:
    """sgemm_tiramisu semantics: returns a*(A@B) + b*C with sequential-k fp32, no FMA."""
    A = np.ascontiguousarray(A, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    out = np.array(Cm, dtype=np.float32, order="C", copy=True)
    N, K = A.shape
    K2, M = B.shape
    assert K == K2 and out.shape == (N, M)
    lib().oracle_sgemm(C.c_int(N), C.c_int(M), C.c_int(K), C.c_float(a), C.c_float(b),
                       _p(A), _p(B), _p(out))
    return out


def sgemm_sched(a: float, b: float, A: np.ndarray, B: np.ndarray, Cm: np.ndarray) -> np.ndarray:
    """oracle.sgemm under the reference's CPU schedule (tiled, 8-wide lanes over j): same bits, timed by bench.py."""
    A = np.ascontiguousarray(A, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    out = np.array(Cm, dtype=np.float32, order="C", copy=True)
    N, K = A.shape
    K2, M = B.shape
    assert K == K2 and out.shape == (N, M)
    lib().oracle_sgemm_sched(C.c_int(N), C.c_int(M), C.c_int(K), C.c_float(a), C.c_float(b),
                             _p(A), _p(B), _p(out))
    return out


def gemm_ld(A, B, Cm, M, N, K, alpha, beta, ldA=0, ldB=0, ldC=0, offA=0, offB=0, offC=0,
            tA=False, tB=False, dtype=np.float32):
    """tiramisu_cublas_{s,d}gemm contract on flat arrays; returns the updated flat C."""
    A = np.ascontiguousarray(A, dtype=dtype).ravel()
    B = np.ascontiguousarray(B, dtype=dtype).ravel()
    out = np.array(Cm, dtype=dtype, copy=True).ravel()
    u = C.c_uint64
    if dtype == np.float32:
        lib().oracle_gemm_ld(_p(A), _p(B), _p(out), u(M), u(N), u(K), C.c_float(alpha),
                             C.c_float(beta), u(ldA), u(ldB), u(ldC), u(offA), u(offB), u(offC),
                             C.c_int(int(tA)), C.c_int(int(tB)))
    else:
        lib().oracle_dgemm_ld(_p(A), _p(B), _p(out), u(M), u(N), u(K), C.c_double(alpha),
                              C.c_double(beta), u(ldA), u(ldB), u(ldC), u(offA), u(offB), u(offC),
                              C.c_int(int(tA)), C.c_int(int(tB)))
    return out


def matmul_u8(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    A = np.ascontiguousarray(A, dtype=np.uint8)
    B = np.ascontiguousarray(B, dtype=np.uint8)
    N = A.shape[0]
    out = np.empty((N, N), dtype=np.uint8)
    lib().oracle_matmul_u8(_p(A), _p(B), _p(out), C.c_int(N))
    return out


def conv_relu_maxpool(inp: np.ndarray, filt: np.ndarray, bias: np.ndarray) -> np.ndarray:
    """inp f32[B][N+2][N+2][FIn], filt f32[FOut/8][3][3][FIn][8], bias f32[FOut]
    -> f32[B][FOut/8][N/2][N/2][8]."""
    inp = np.ascontiguousarray(inp, dtype=np.float32)
    filt = np.ascontiguousarray(filt, dtype=np.float32)
    bias = np.ascontiguousarray(bias, dtype=np.float32).ravel()
    Bn, NP, _, FIn = inp.shape
    FB = filt.shape[0]
    N = NP - 2
    out = np.empty((Bn, FB, N // 2, N // 2, 8), dtype=np.float32)
    lib().oracle_conv_relu_maxpool(_p(inp), _p(filt), _p(bias), _p(out), C.c_int(Bn), C.c_int(N),
                                   C.c_int(FIn), C.c_int(FB * 8))
    return out


def lstm(Weights: np.ndarray, biases: np.ndarray, x: np.ndarray) -> np.ndarray:
    """Weights f32[L][2][4F][F] (0=R,1=W), biases f32[L][4F], x f32[T][B][F] -> y f32[T][B][F]."""
    Weights = np.ascontiguousarray(Weights, dtype=np.float32)
    biases = np.ascontiguousarray(biases, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    L, two, F4, F = Weights.shape
    assert two == 2 and F4 == 4 * F
    T, Bn, F2 = x.shape
    assert F2 == F
    y = np.empty_like(x)
    lib().oracle_lstm(C.c_int(F), C.c_int(Bn), C.c_int(L), C.c_int(T), _p(Weights), _p(biases),
                      _p(x), _p(y))
    return y


def lstm_with_state(Weights: np.ndarray, biases: np.ndarray, x: np.ndarray):
    """oracle.lstm plus the last layer's final cell state c f32[B][F] (LSTM/cpu_lib's `c` buffer)."""
    Weights = np.ascontiguousarray(Weights, dtype=np.float32)
    biases = np.ascontiguousarray(biases, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    L, two, F4, F = Weights.shape
    T, Bn, _ = x.shape
    y = np.empty_like(x)
    c = np.empty((Bn, F), np.float32)
    lib().oracle_lstm_c(C.c_int(F), C.c_int(Bn), C.c_int(L), C.c_int(T), _p(Weights), _p(biases), _p(x), _p(y), _p(c))
    return y, c


def lstm_relayout(Weights: np.ndarray, biases: np.ndarray, layout: int):
    """LSTM/cpu layout (f32[L][2][4F][F], 0 = R, 1 = W, gate row blocks i, z, o, f) -> the parameter layout of
    LSTM/gpu_lib (layout 1: 0 = W, 1 = R, blocks i, f, z, o; gpu_lib/generator.cpp:66-88) or LSTM/cpu_lib (layout 2: the
    same, each matrix transposed to [F][4F]; cpu_lib/generator.cpp:27).  Returns (weights, biases)."""
    L, _, F4, F = Weights.shape
    blocks = Weights.reshape(L, 2, 4, F, F)                  # [l][m][gate i,z,o,f][u][j]
    theirs = blocks[:, ::-1][:, :, [0, 3, 1, 2]]             # matrices swapped; gates i, f, z, o
    w = np.ascontiguousarray(theirs.reshape(L, 2, 4 * F, F))
    b = np.ascontiguousarray(biases.reshape(L, 4, F)[:, [0, 3, 1, 2]].reshape(L, 4 * F))
    if layout == 2:
        w = np.ascontiguousarray(w.transpose(0, 1, 3, 2))
    return w, b


def transpose_f32(A: np.ndarray) -> np.ndarray:
    A = np.ascontiguousarray(A, dtype=np.float32)
    r, c = A.shape
    out = np.empty((c, r), dtype=np.float32)
    lib().oracle_transpose_f32(_p(A), _p(out), C.c_int(r), C.c_int(c))
    return out

def conv_layer_blocked(inp: np.ndarray, filt: np.ndarray, bias: np.ndarray) -> np.ndarray:
    """The direct convolution layer with blocked layouts (reference
    benchmarks/DNN/layers/convolution/direct/cpu/conv_layer_generator_tiramisu.cpp:21-45, Layer I only):
        out[n][fout_b][y][x][ffout] = bias[fout_b][ffout]
                                      + sum over fin_b, k_y, k_x, ffin of filt[fout_b][fin_b][k_y][k_x][ffin][ffout]
                                                                          * inp[n][fin_b][y + k_y][x + k_x][ffin]
    in float64 (numpy), rounded to float32 at the end — a checker for a float kernel, compared with a tolerance.
    Parity unpinned: the reference holds no golden output for this layer (its wrapper compares with MKL-DNN, absent here)."""
    inp = np.asarray(inp, np.float64)
    filt = np.asarray(filt, np.float64)
    n, finb, yp, xp, ffin = inp.shape
    foutb, finb2, ky, kx, ffin2, ffout = filt.shape
    assert finb == finb2 and ffin == ffin2 and ky == kx
    y, x = yp - ky + 1, xp - kx + 1
    out = np.zeros((n, foutb, y, x, ffout), np.float64)
    for dy in range(ky):
        for dx in range(kx):
            win = inp[:, :, dy:dy + y, dx:dx + x, :]                          # n b y x i
            out += np.einsum("nbyxi,obif->noyxf", win, filt[:, :, dy, dx, :, :], optimize=True)
    out += np.asarray(bias, np.float64).reshape(1, foutb, 1, 1, ffout)
    return out.astype(np.float32)


def vgg_block(inp: np.ndarray, f1: np.ndarray, b1: np.ndarray, f2: np.ndarray, b2: np.ndarray):
    """The dense VGG block (reference benchmarks/DNN/blocks/vggBlock/cpu/dense/vgg_block_generator_tiramisu.cpp:15-96,
    Layer I only), float64, rounded to float32 at the end:
        conv1[n][b][y][x][f]  = relu(bias1[b][f] + sum over k_y, k_x, fin1 of f1[b][k_y][k_x][fin1][f] * inp[n][y+k_y][x+k_x][fin1])
        padded                = conv1 with a one-element zero border                               (conv1_buf, second result)
        conv2[n][y][x][o][g]  = bias2[o][g] + sum over b, k_y, k_x, f of f2[o][b][k_y][k_x][f][g] * padded[n][b][y+k_y][x+k_x][f]
        out[n][o][y/2][x/2][g] = max(0, conv2 over the 2 x 2 window)                               (max-pool initialised with 0)
    Returns (padded, out).  Parity unpinned: the reference's wrapper compares with MKL-DNN, absent here."""
    inp, f1, f2 = np.asarray(inp, np.float64), np.asarray(f1, np.float64), np.asarray(f2, np.float64)
    n, yp, xp, fin1 = inp.shape
    nb, ky, kx, _, nf = f1.shape
    y, x = yp - ky + 1, xp - kx + 1
    c1 = np.zeros((n, nb, y, x, nf), np.float64)
    for dy in range(ky):
        for dx in range(kx):
            c1 += np.einsum("nyxi,bif->nbyxf", inp[:, dy:dy + y, dx:dx + x, :], f1[:, dy, dx, :, :], optimize=True)
    c1 = np.maximum(c1 + np.asarray(b1, np.float64).reshape(1, nb, 1, 1, nf), 0.0)
    padded = np.zeros((n, nb, y + 2, x + 2, nf), np.float64)
    padded[:, :, 1:-1, 1:-1, :] = c1
    no, _, _, _, _, ng = f2.shape
    c2 = np.zeros((n, y, x, no, ng), np.float64)
    for dy in range(ky):
        for dx in range(kx):
            c2 += np.einsum("nbyxf,obfg->nyxog", padded[:, :, dy:dy + y, dx:dx + x, :], f2[:, :, dy, dx, :, :], optimize=True)
    c2 += np.asarray(b2, np.float64).reshape(1, 1, 1, no, ng)
    out = np.zeros((n, no, y // 2, x // 2, ng), np.float64)
    for dy in range(2):
        for dx in range(2):
            out = np.maximum(out, c2[:, dy:2 * (y // 2):2, dx:2 * (x // 2):2, :, :].transpose(0, 3, 1, 2, 4))
    return padded.astype(np.float32), out.astype(np.float32)
