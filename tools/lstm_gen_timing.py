"""Bring-up: time the reference's LSTM/cpu/generator.cpp as lowered by the front-end (pattern = warp-shuffle
reduction for the gate sums, generic = one output per thread) at a mid-size shape, through its ABI."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiramisu_b200 import _lib
from tiramisu_b200.halide import Buffer
_lib.lib()
F, B, L, T = 512, 32, 2, 10
rng = np.random.default_rng(0)
Wt = (rng.uniform(-1, 1, (L, 2, 4 * F, F)) / 16).astype(np.float32)
bias = rng.uniform(-1, 1, (L, 4 * F)).astype(np.float32)
x = rng.uniform(-1, 1, (T, B, F)).astype(np.float32)
for variant in ("lstm_cpu_so_generic", "lstm_cpu_so"):
    lib = C.CDLL(os.path.join(os.path.dirname(__file__), "..", "tests", "compat", "_bin", f"libgen_{variant}.so"), mode=C.RTLD_GLOBAL)
    y = np.zeros((T, B, F), np.float32)
    bufs = [Buffer(a) for a in (np.array([F, B, L, T], np.int32), Wt, bias, x, y)]
    lib.lstm.restype = C.c_int
    lib.lstm(*[b.ptr() for b in bufs])
    t0 = time.perf_counter()
    for _ in range(3):
        lib.lstm(*[b.ptr() for b in bufs])
    print(variant, f"F={F} B={B} L={L} T={T}", "ms per call (host buffers, whole call):", round((time.perf_counter() - t0) / 3 * 1e3, 2), flush=True)
