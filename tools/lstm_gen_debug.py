"""Bring-up: error of the lowered LSTM/cpu/generator.cpp (both lowerings) against the oracle at a few shapes."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle
from tiramisu_b200 import _lib
from tiramisu_b200.halide import Buffer
_lib.lib()
for variant in ("lstm_cpu_so", "lstm_cpu_so_generic"):
    lib = C.CDLL(os.path.join(os.path.dirname(__file__), "..", "tests", "compat", "_bin", f"libgen_{variant}.so"), mode=C.RTLD_GLOBAL)
    lib.lstm.restype = C.c_int
    for (F, B, L, T) in ((32, 8, 2, 3), (96, 5, 3, 7), (256, 16, 2, 4)):
        rng = np.random.default_rng(20)
        Wt = (rng.uniform(-1, 1, (L, 2, 4 * F, F)) / 16).astype(np.float32)
        bias = rng.uniform(-1, 1, (L, 4 * F)).astype(np.float32)
        x = rng.uniform(-1, 1, (T, B, F)).astype(np.float32)
        y = np.zeros((T, B, F), np.float32)
        bufs = [Buffer(a) for a in (np.array([F, B, L, T], np.int32), Wt, bias, x, y)]
        lib.lstm(*[b.ptr() for b in bufs])
        want = oracle.lstm(Wt, bias, x)
        d = np.abs(y - want)
        print(variant, (F, B, L, T), "max err", float(d.max()), flush=True)
