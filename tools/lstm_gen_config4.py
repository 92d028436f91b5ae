"""BASELINE configs[3] (F=1024, B=64, L=4, T=100) through the reference's own LSTM/cpu/generator.cpp as lowered by the
front-end (pattern lowering: gate sums -> tcgen05 GEMM calls), device-resident arguments, next to the runtime's lstm engine."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiramisu_b200 import _lib, api
from tiramisu_b200.halide import Buffer
L_ = _lib.lib()
F, B, L, T = 1024, 64, 4, 100
s = torch.cuda.Stream(); api.set_stream(s.cuda_stream)
with torch.cuda.stream(s):
    g = torch.Generator(device="cuda").manual_seed(4)
    W = (torch.rand((L, 2, 4 * F, F), device="cuda", generator=g) * 2 - 1) / 32
    b = torch.rand((L, 4 * F), device="cuda", generator=g) * 2 - 1
    x = torch.rand((T, B, F), device="cuda", generator=g) * 2 - 1
    y_gen = torch.zeros_like(x); y_lib = torch.zeros_like(x)
s.synchronize()
lib = C.CDLL(os.path.join(os.path.dirname(__file__), "..", "tests", "compat", "_bin", "libgen_lstm_cpu_so.so"), mode=C.RTLD_GLOBAL)
lib.lstm.restype = C.c_int
params = np.array([F, B, L, T], np.int32)
bufs = [Buffer(a) for a in (params, W, b, x, y_gen)]
def gen(): assert lib.lstm(*[q.ptr() for q in bufs]) == 0
def eng(): _lib.check(L_.tiramisu_b200_lstm(W.data_ptr(), b.data_ptr(), x.data_ptr(), y_lib.data_ptr(), F, B, L, T))
for fn, name in ((gen, "generated from LSTM/cpu/generator.cpp (pattern lowering)"), (eng, "runtime lstm engine")):
    fn(); fn(); s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(3): fn()
    e1.record(s); torch.cuda.synchronize()
    print(name, "ms per call:", round(e0.elapsed_time(e1) / 3, 3), flush=True)
print("max |generated - engine| =", float((y_gen - y_lib).abs().max()), flush=True)
