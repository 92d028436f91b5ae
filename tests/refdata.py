"""Input generators that replay the reference wrappers' own initialisation code.

The reference seeds glibc's rand() (std::srand/std::rand) and fills buffers in a fixed loop
order; we call the same libc through ctypes so the synthetic inputs are the reference's.
"""
import ctypes

import numpy as np

from oracle import oracle as _oracle


def srand(seed: int) -> None:
    _oracle.lib().oracle_srand(ctypes.c_uint(seed))


def rand_array(n: int) -> np.ndarray:
    """n successive std::rand() values (glibc, same libc the reference wrappers use) as int64."""
    out = np.empty(n, dtype=np.int32)
    _oracle.lib().oracle_rand_fill(out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n))
    return out.astype(np.int64)


def rand_small_ints(shape) -> np.ndarray:
    """`std::rand() % 10 - 5` (tests/wrapper_test_162.cpp:19-27, sgemm/gpu/wrapper.cpp:53-55)."""
    n = int(np.prod(shape))
    return (rand_array(n) % 10 - 5).astype(np.float32).reshape(shape)


def conv_inputs(B, N, FIn, FOut, seed=1):
    """benchmarks/DNN/blocks/Conv-ReLU-MaxPool/cpu/dense/wrapper_nn_block.cpp:15-39:
    srand(1); filter (loops fout, fin, ky, kx), bias (fout), input (n, fin, y, x), each
    ((float)(rand()%256 - 128)) / 127.f; returned in the generated function's layouts."""
    srand(seed)
    val = lambda r: ((r % 256 - 128).astype(np.float32)) / np.float32(127.0)
    f = val(rand_array(FOut * FIn * 9)).reshape(FOut, FIn, 3, 3)            # [fout][fin][ky][kx]
    filt = np.ascontiguousarray(
        f.reshape(FOut // 8, 8, FIn, 3, 3).transpose(0, 3, 4, 2, 1))        # [fb][ky][kx][fin][ff]
    bias = val(rand_array(FOut))
    i = val(rand_array(B * FIn * (N + 2) * (N + 2))).reshape(B, FIn, N + 2, N + 2)
    inp = np.ascontiguousarray(i.transpose(0, 2, 3, 1))                     # [n][y][x][fin]
    return inp, filt, bias


def lstm_inputs(F, B, L, T, seed=0, scale=1.0):
    """benchmarks/DNN/blocks/LSTM/cpu/wrapper.cpp:31-52: srand(0); Weights, biases, x filled in
    C order with (rand()%200 - 100)/100. (double, then stored as float). `scale` divides the
    weights (at F=1024 the raw values saturate every gate)."""
    srand(seed)
    val = lambda r: ((r % 200 - 100).astype(np.float64) / 100.0)
    W = (val(rand_array(L * 2 * 4 * F * F)) / scale).astype(np.float32).reshape(L, 2, 4 * F, F)
    b = val(rand_array(L * 4 * F)).astype(np.float32).reshape(L, 4 * F)
    x = val(rand_array(T * B * F)).astype(np.float32).reshape(T, B, F)
    return W, b, x


def regblock_case(seed=3):
    """Inputs and fp32 results (statement order: bias first, then k = 0, 1, ... with separate multiply and add) of
    tests/frontend_cases/regblock_cpu_gen.cpp: (in, w, bias), conv[N][Y][X][8], pooled[N][Y/2][X/2][8]."""
    N, Y, X, K = 3, 6, 12, 5
    rng = np.random.default_rng(seed)
    inp = rng.uniform(-1, 1, (N, Y, X + K - 1)).astype(np.float32)
    w = rng.uniform(-1, 1, (K, 8)).astype(np.float32)
    bias = rng.uniform(-1, 1, (8,)).astype(np.float32)
    conv = np.broadcast_to(bias, (N, Y, X, 8)).astype(np.float32).copy()
    for k in range(K):
        prod = (w[k][None, None, None, :] * inp[:, :, k:k + X, None]).astype(np.float32)
        conv = (conv + prod).astype(np.float32)
    pooled = np.zeros((N, Y // 2, X // 2, 8), np.float32)
    for dy in range(2):
        for dx in range(2):
            pooled = np.maximum(pooled, conv[:, dy::2, dx::2, :])
    return (inp, w, bias), conv, pooled
