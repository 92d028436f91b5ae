#!/usr/bin/env python
"""Generate tests/golden/{conv,lstm}_*.npy: INDEPENDENT fp64 results for the two blocks whose reference
checks rest on MKL / MKL-DNN dumps that cannot be produced here (wrapper_nn_block.cpp:77-90,
LSTM/cpu/wrapper.cpp:96-111).

Inputs are the reference wrappers' own initialisation (glibc srand/rand replay, tests/refdata.py:
wrapper_nn_block.cpp:15-39 and LSTM/cpu/wrapper.cpp:31-52); the results come from PyTorch on the CPU in
float64 — torch.nn.functional.conv2d + relu + max_pool2d, and a hand-rolled LSTM cell with the reference's
gate order i, z, o, f — i.e. from code that shares nothing with oracle/tiramisu_oracle.c or the CUDA path.
tests/test_oracle.py and tests/test_gpu_dnn.py then apply the reference wrappers' own protocol
("percentage of correctness": |a - b| <= 1e-4 for the conv block, < 0.02 for the LSTM) against them.

    python tools/make_golden_dnn.py        (writes float32-rounded fp64 results, ~400 KB in total)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import refdata  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CONV_CASES = {"conv_default": (2, 32, 3, 32),      # configure.h defaults FIn=3, FOut=32 (small N, first 2 images)
              "conv_config3": (2, 32, 64, 64)}     # BASELINE configs[2] channels 64 -> 64
LSTM_CASES = {"lstm_readme": (32, 8, 4, 100, 1.0),     # LSTM/cpu/README.md:3 sizes, the wrapper's raw (saturating) data
              "lstm_scaled": (128, 16, 2, 10, 8.0)}    # weights / 8: gates not saturated


def conv_fp64(inp, filt, bias):
    """inp f32[B][N+2][N+2][FIn] (pre-padded NHWC), filt f32[FOut/8][3][3][FIn][8], bias f32[FOut]
    -> fp64 [B][FOut/8][N/2][N/2][8]."""
    x = torch.from_numpy(inp).double().permute(0, 3, 1, 2)                                   # NCHW
    FB, _, _, FIn, _ = filt.shape
    w = torch.from_numpy(filt).double().permute(0, 4, 3, 1, 2).reshape(FB * 8, FIn, 3, 3)    # [fout][fin][ky][kx]
    y = torch.nn.functional.conv2d(x, w, torch.from_numpy(bias).double())
    y = torch.nn.functional.max_pool2d(torch.relu(y), 2)
    B, FOut, NO, _ = y.shape
    return y.reshape(B, FB, 8, NO, NO).permute(0, 1, 3, 4, 2).contiguous().numpy()


def lstm_fp64(W, b, x):
    """Weights f32[L][2][4F][F] (0 = R recurrent, 1 = W input), gate rows i, z, o, f (LSTM/cpu/generator.cpp:43-46)."""
    L, _, F4, F = W.shape
    T, B, _ = x.shape
    Wd, bd = torch.from_numpy(W).double(), torch.from_numpy(b).double()
    below = torch.from_numpy(x).double()
    for l in range(L):
        h = torch.zeros(B, F, dtype=torch.float64)
        c = torch.zeros(B, F, dtype=torch.float64)
        outs = []
        for t in range(T):
            s = bd[l] + h @ Wd[l, 0].T + below[t] @ Wd[l, 1].T
            i, z, o, f = torch.sigmoid(s[:, :F]), torch.tanh(s[:, F:2 * F]), torch.sigmoid(s[:, 2 * F:3 * F]), torch.sigmoid(s[:, 3 * F:])
            c = i * z + f * c
            h = torch.tanh(c) * o
            outs.append(h)
        below = torch.stack(outs)
    return below.numpy()


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (B, N, FIn, FOut) in CONV_CASES.items():
        inp, filt, bias = refdata.conv_inputs(B, N, FIn, FOut)
        out = conv_fp64(inp, filt, bias).astype(np.float32)
        np.save(os.path.join(OUT, f"{name}_out.npy"), out)
        print(name, out.shape, float(np.abs(out).max()))
    for name, (F, B, L, T, scale) in LSTM_CASES.items():
        W, b, x = refdata.lstm_inputs(F, B, L, T, scale=scale)
        y = lstm_fp64(W, b, x).astype(np.float32)
        np.save(os.path.join(OUT, f"{name}_y.npy"), y)
        print(name, y.shape, float(np.abs(y).max()))


if __name__ == "__main__":
    main()
