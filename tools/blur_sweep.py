"""Bring-up: time the fp32 blur kernel on 16384^2 for one TB200_BLUR_RY value (set in the environment)."""
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiramisu_b200 import _lib, api
L = _lib.lib()
N = 16384
s = torch.cuda.Stream(); api.set_stream(s.cuda_stream)
with torch.cuda.stream(s):
    a = torch.rand((N, N), device="cuda"); o = torch.empty((N - 2, N - 2), device="cuda"); h = torch.empty((N, N), device="cuda")
def blur(): _lib.check(L.tiramisu_b200_blur_f32(a.data_ptr(), o.data_ptr(), N, N, 0, N - 2))
def heat(): _lib.check(L.tiramisu_b200_heat2d(a.data_ptr(), h.data_ptr(), N, N, 0, 0, N))
for fn, name in ((blur, "blur"), (heat, "heat")):
    for _ in range(5): fn()
    s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(20): fn()
    e1.record(s); torch.cuda.synchronize()
    print(name, "RY", os.environ.get("TB200_BLUR_RY", "default"), "ms", round(e0.elapsed_time(e1) / 20, 4), flush=True)
