"""Device-resident timing of the reference's DNN generators as lowered by the front-end: the direct convolution
layer (conv_layer_generator_tiramisu.cpp, batch 8, 112 x 112, 128 -> 256) and the dense VGG block (batch 32, 112 x 112,
3 -> 32 -> 32), plus the reference's own sgemm CPU generator (one tcgen05 GEMM call) at 1024^3 / 4096^3."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiramisu_b200 import _lib, api
from tiramisu_b200.halide import Buffer
_lib.lib()
BIN = os.path.join(os.path.dirname(__file__), "..", "tests", "compat", "_bin")
s = torch.cuda.Stream(); api.set_stream(s.cuda_stream)


def timed(fn, reps=5):
    fn(); fn(); s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(reps): fn()
    e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def dev(shape, seed):
    with torch.cuda.stream(s):
        g = torch.Generator(device="cuda").manual_seed(seed)
        return torch.rand(shape, device="cuda", generator=g) * 2 - 1


def run(libname, fn, arrays, flop, what):
    path = os.path.join(BIN, libname)
    if not os.path.exists(path):
        print(what, ": not built"); return
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    f = getattr(lib, fn); f.restype = C.c_int
    bufs = [Buffer(a) for a in arrays]
    def call(): assert f(*[b.ptr() for b in bufs]) == 0
    ms = timed(call)
    print(f"{what}: {ms:.3f} ms per call, {flop / ms / 1e9:.1f} TFLOP/s", flush=True)


run("libgen_conv_layer.so", "conv_tiramisu", [dev((8, 8, 114, 114, 16), 1), dev((32, 8, 3, 3, 16, 8), 2), dev((32, 8), 3), dev((8, 32, 112, 112, 8), 4)],
    2.0 * 8 * 256 * 112 * 112 * 128 * 9, "conv layer generator (3 SIMT kernels)")
run("libgen_vgg_block.so", "vgg_block", [dev((32, 114, 114, 3), 1), dev((4, 3, 3, 3, 8), 2), dev((4, 8), 3), dev((4, 4, 3, 3, 8, 8), 4), dev((4, 8), 5),
                                          dev((32, 4, 114, 114, 8), 6), dev((32, 4, 56, 56, 8), 7)],
    2.0 * 32 * 112 * 112 * 9 * (3 * 32 + 32 * 32), "vgg block generator (11 SIMT kernels)")
