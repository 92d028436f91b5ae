"""the reference's other `*gpu` Halide benchmarks (benchmarks/halide/{convolution,gaussian,edge,fusion,
cvtcolor,rgbyuv420,heat2d,warp_affine}gpu_tiramisu.cpp) compiled UNMODIFIED against the front-end (tests/compat/build_generators.py,
targets halide_*), lowered to sm_100a and checked against numpy restatements of their Layer-I expressions
(oracle/halide_gpu.py).  Their own wrappers compare with Halide-generated pipelines and load PNG files, neither of
which exists here, so the generated functions are called through ctypes on synthetic images.

  -m gpu      : every generator on the device, bit-exact (uint8 / exact-in-fp32 data)
  not gpu     : the generators without barriers / shared memory through the host emulation of generated sources
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import halide_gpu as hg
from oracle import oracle
from tests.compat import build_generators as bg


def _lib_of(name):
    path = os.path.join(bg.BIN, f"libgen_halide_{name}.so")
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (tests/compat/build_generators.py needs /root/reference)")
    from tiramisu_b200 import _lib
    _lib.lib()                      # runtime symbols first (RTLD_GLOBAL)
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


def _emulated(name):
    from tests.emulation import emulate
    src = os.path.join(bg.GEN, f"halide_{name}", "build", f"generated_fct_{name}.o.b200.cu")
    if not os.path.exists(src):
        pytest.skip(f"{src} not generated")
    return C.CDLL(emulate.build_shared([src], f"halide_{name}"))


def _call(lib, fn, *arrays):
    from tiramisu_b200.halide import Buffer
    bufs = [Buffer(a) for a in arrays]
    f = getattr(lib, fn)
    f.restype = C.c_int
    assert f(*[b.ptr() for b in bufs]) == 0


def _image(seed, C_, M, N):
    return np.random.default_rng(seed).integers(0, 256, (C_, M, N), dtype=np.uint8)


# ---------------------------------------------------------------------------------------------------- the checks
def check_cvtcolor(lib, M=75, N=130):
    img = _image(1, 3, M, N)
    out = np.full((M, N), 7, np.uint8)
    _call(lib, "cvtcolorgpu_tiramisu", np.array([N, M], np.int32), img, out)
    assert np.array_equal(out, hg.cvtcolorgpu(img))


def check_fusion(lib, M=50, N=70):
    img = _image(2, 3, M, N)
    outs = [np.full((3, M, N), 9, np.uint8) for _ in range(4)]
    _call(lib, "fusiongpu_tiramisu", np.array([N, M], np.int32), img, *outs)
    for got, want in zip(outs, hg.fusiongpu(img)):
        assert np.array_equal(got, want)


def check_rgbyuv420(lib, M=64, N=96):
    img = _image(3, 3, M, N)
    y, u, v = np.zeros((M, N), np.uint8), np.zeros((M // 2, N // 2), np.uint8), np.zeros((M // 2, N // 2), np.uint8)
    _call(lib, "rgbyuv420gpu_tiramisu", np.array([N, M], np.int32), img, y, u, v)
    wy, wu, wv = hg.rgbyuv420gpu(img)
    assert np.array_equal(y, wy) and np.array_equal(u, wu) and np.array_equal(v, wv)


def check_gaussian(lib, M=70, N=101, C_=3):
    img = _image(4, C_, M, N)
    # dyadic taps that sum to 1: every product and partial sum is exact in fp32 and stays inside [0, 255]
    kx = np.array([1, 4, 6, 4, 1], np.float32) / 16
    ky = np.array([1, 2, 2, 2, 1], np.float32) / 8
    out = np.full((C_, M - 5, N - 5), 3, np.uint8)
    _call(lib, "gaussiangpu_tiramisu", np.array([N, M, C_], np.int32), img, kx, ky, out)
    assert np.array_equal(out, hg.gaussiangpu(img, kx, ky))


def check_convolution(lib, M=83, N=150):
    img = _image(5, 3, M, N)
    kernel = np.array([[0, 2, 0], [2, 4, 2], [1, 4, 1]], np.float32) / 16        # kernel_host[_j][_i]; asymmetric on purpose
    out = np.full((3, M - 2, N - 2), 5, np.uint8)
    _call(lib, "convolutiongpu_tiramisu", np.array([N, M], np.int32), img, kernel, out)
    assert np.array_equal(out, hg.convolutiongpu(img, kernel))


def check_warp_affine(lib, M=70, N=101):
    """warp_affinegpu_tiramisu.cpp: a hand-tiled iteration domain (`0 <= i0 * 16 + i1 < N and 0 <= i1 < 16`), constants declared at
    a loop level (lets), data-dependent reads, floor / select; sizes that are not multiples of the 16 x 16 tile."""
    img = np.random.default_rng(7).integers(0, 256, (M, N), dtype=np.uint8)
    out = np.full((M, N), -3.0, np.float32)
    _call(lib, "warp_affinegpu_tiramisu", np.array([N, M], np.int32), img, out)
    assert np.array_equal(out, hg.warp_affinegpu(img))


# ---------------------------------------------------------------------------------------------------- on the device
@pytest.mark.gpu
@pytest.mark.parametrize("name,check", [("cvtcolorgpu", check_cvtcolor), ("fusiongpu", check_fusion), ("rgbyuv420gpu", check_rgbyuv420),
                                        ("gaussiangpu", check_gaussian), ("convolutiongpu", check_convolution), ("warp_affinegpu", check_warp_affine)])
def test_halide_gpu_benchmark_bit_exact(name, check):
    check(_lib_of(name))


@pytest.mark.gpu
@pytest.mark.parametrize("mn", [(2112, 3520), (48, 48)])
def test_halide_gpu_benchmarks_at_image_size(mn):
    """The size of the reference's utils/images/rgb.png (2112 x 3520 x 3) and a size smaller than one tile row."""
    N, M = mn
    check_cvtcolor(_lib_of("cvtcolorgpu"), M, N)
    check_gaussian(_lib_of("gaussiangpu"), M, N)
    check_convolution(_lib_of("convolutiongpu"), M, N)
    check_fusion(_lib_of("fusiongpu"), M, N)
    check_rgbyuv420(_lib_of("rgbyuv420gpu"), M, N)


@pytest.mark.gpu
def test_halide_edgegpu_in_place():
    """edgegpu_tiramisu.cpp (function `edge_tiramisu`, fixed 8192 x 8192 x 3, result stored back into the image).
    Checked where the reference's result is defined (its last two columns / last row read a temporary that was
    never written, oracle/halide_gpu.py: edgegpu) and where the image must be untouched."""
    lib = _lib_of("edgegpu")
    img = _image(6, 3, 8192, 8192)
    want, (ys, xs) = hg.edgegpu(img)
    got = img.copy()
    _call(lib, "edge_tiramisu", got)
    assert np.array_equal(got[:, ys, xs], want[:, ys, xs])
    assert np.array_equal(got[:, 8190:, :], img[:, 8190:, :]) and np.array_equal(got[:, :, 8190:], img[:, :, 8190:])


@pytest.mark.gpu
@pytest.mark.parametrize("n", [256, 1000, 30])
def test_halide_heat2dgpu_generator(n):
    """heat2dgpu_tiramisu.cpp:26-138 through codegen(): two-point border domains `(i = 0 or i = M - 1)`, a 32 x 32
    shared tile with 30 x 30 useful points, allocate + sync inside the kernel.  Same association as the
    hand-registered heat2dgpu_tiramisu (form 1): bit-exact against the oracle.  Square grids: the generator
    takes sizes(0) as the row count while its wrapper passes the width."""
    lib = _lib_of("heat2dgpu")
    a = np.random.default_rng(8).random((n, n), dtype=np.float32)
    out = np.full_like(a, -5.0)
    _call(lib, "heat2dgpu_tiramisu", np.array([n, n], np.int32), a, out)
    assert np.array_equal(out, oracle.heat2d(a, "gpu"))


def check_blurxy(lib):
    """benchmarks/halide/blurxy_tiramisu.cpp (row a7's own generator): sized at generator time from the reference's
    utils/images/rgb.png (2112 x 3520 x 3); uint8 wrap-around sums, the x pass materialised once (compute_at)."""
    img = _image(9, 3, 3520, 2112)
    out = np.full((3, 3512, 2104), 11, np.uint8)
    _call(lib, "blurxy_tiramisu", img, out)
    assert np.array_equal(out, oracle.blurxy_u8(img))


@pytest.mark.gpu
def test_halide_blurxy_generator():
    check_blurxy(_lib_of("blurxy"))


# ---------------------------------------------------------------------------------------------------- host emulation
def test_emulated_halide_blurxy_generator():
    check_blurxy(_emulated("blurxy"))



@pytest.mark.parametrize("name,check", [("cvtcolorgpu", check_cvtcolor), ("fusiongpu", check_fusion), ("rgbyuv420gpu", check_rgbyuv420),
                                        ("gaussiangpu", check_gaussian), ("warp_affinegpu", check_warp_affine)])
def test_emulated_halide_gpu_benchmark(name, check):
    """The generators whose kernels need no barrier / shared memory, run thread by thread on the host."""
    check(_emulated(name))
