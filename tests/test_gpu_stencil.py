"""GPU parity of the stencil path (heat2d, blurs) against the CPU oracle — through the C ABI."""
import os

import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def api():
    from tiramisu_b200 import api as _api
    return _api


# ------------------------------------------------------------------------------------------ heat2d
@pytest.mark.parametrize("shape", [(64, 128), (70, 132), (33, 516), (9, 8), (3, 4), (257, 1024)])
def test_heat2d_tiramisu_bit_exact_vector_path(api, shape):
    rng = np.random.default_rng(1)
    a = rng.random(shape, dtype=np.float32)
    out = np.full(shape, np.nan, np.float32)
    api.heat2d_tiramisu(a, out)
    assert np.array_equal(out, oracle.heat2d(a, "cpu"))


@pytest.mark.parametrize("shape", [(17, 19), (5, 3), (3, 3), (40, 1001), (1, 7), (7, 1), (2, 2)])
def test_heat2d_tiramisu_bit_exact_ragged(api, shape):
    """W % 4 != 0 -> scalar kernel; degenerate grids have only borders (all zero)."""
    rng = np.random.default_rng(11)
    a = rng.random(shape, dtype=np.float32)
    out = np.full(shape, np.nan, np.float32)
    api.heat2d_tiramisu(a, out)
    assert np.array_equal(out, oracle.heat2d(a, "cpu"))


def test_heat2d_all_ones_known_answer(api):
    """benchmarks/halide/wrapper_heat2d.cpp:17-22: input all 1 => interior 1.9f, border 0."""
    a = np.ones((128, 256), np.float32)
    out = np.empty_like(a)
    api.heat2d_tiramisu(a, out)
    want = np.float32(np.float32(0.3) * np.float32(1)) + np.float32(np.float32(0.4) * np.float32(4))
    assert (out[1:-1, 1:-1] == want).all()
    assert (out[0] == 0).all() and (out[-1] == 0).all() and (out[:, 0] == 0).all() and (out[:, -1] == 0).all()


@pytest.mark.parametrize("shape", [(64, 128), (37, 41)])
def test_heat2dgpu_tiramisu_bit_exact(api, shape):
    rng = np.random.default_rng(2)
    a = rng.random(shape, dtype=np.float32)
    out = np.full(shape, np.nan, np.float32)
    sizes = np.array([shape[1], shape[0]], np.int32)          # wrapper_heat2dgpu.cpp:32-34
    api.heat2dgpu_tiramisu(sizes, a, out)
    assert np.array_equal(out, oracle.heat2d(a, "gpu"))


def test_heat2d_device_resident_and_shards(api):
    """Device-resident buffers (->device) skip the copies; row shards reproduce the full sweep."""
    import torch
    from tiramisu_b200 import _lib
    rng = np.random.default_rng(3)
    a = rng.random((96, 256), dtype=np.float32)
    want = oracle.heat2d(a, "cpu")
    d_in = torch.from_numpy(a).cuda()
    d_out = torch.full_like(d_in, float("nan"))
    api.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        api.heat2d_tiramisu(d_in, d_out)
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy(), want)
        d_out.fill_(float("nan"))
        L = _lib.lib()
        for rb, re_ in ((0, 31), (31, 64), (64, 96)):
            _lib.check(L.tiramisu_b200_heat2d(d_in.data_ptr(), d_out.data_ptr(), 96, 256, 0, rb, re_))
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy(), want)
    finally:
        api.set_stream(None)


def test_heat2d_dist_step_form(api):
    """form 2 = one Jacobi half-step of heat2d_dist (borders of out untouched)."""
    import torch
    from tiramisu_b200 import _lib
    rng = np.random.default_rng(4)
    for shape in ((40, 64), (21, 23)):
        a = rng.random(shape, dtype=np.float32)
        prev = rng.random(shape, dtype=np.float32)
        want = oracle.heat2d_step(a, prev.copy())
        d_in = torch.from_numpy(a).cuda()
        d_out = torch.from_numpy(prev).cuda()
        api.set_stream(torch.cuda.current_stream().cuda_stream)
        try:
            _lib.check(_lib.lib().tiramisu_b200_heat2d(d_in.data_ptr(), d_out.data_ptr(), shape[0], shape[1], 2,
                                                       0, shape[0]))
            torch.cuda.synchronize()
        finally:
            api.set_stream(None)
        assert np.array_equal(d_out.cpu().numpy(), want)


def test_heat2d_full_size_properties(api):
    """BASELINE config 2 size (16384^2): closed form on all-ones, linearity-free checks that do not
    need the oracle: interior == 1.9f, borders == 0, and a checksum of a random strip vs the oracle."""
    import torch
    N = 16384
    d_in = torch.ones((N, N), dtype=torch.float32, device="cuda")
    d_out = torch.empty_like(d_in)
    api.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        api.heat2d_tiramisu(d_in, d_out)
        torch.cuda.synchronize()
        want = np.float32(np.float32(0.3)) + np.float32(np.float32(0.4) * np.float32(4))
        assert bool((d_out[1:-1, 1:-1] == float(want)).all())
        assert float(d_out[0].abs().sum()) == 0 and float(d_out[:, -1].abs().sum()) == 0
        # random data: compare a 64-row strip against the oracle run on strip+halo
        g = torch.Generator(device="cuda").manual_seed(5)
        d_in = torch.rand((N, N), dtype=torch.float32, device="cuda", generator=g)
        api.heat2d_tiramisu(d_in, d_out)
        torch.cuda.synchronize()
        y0 = 8191
        strip = d_in[y0 - 1:y0 + 65].cpu().numpy()
        want = oracle.heat2d(strip, "cpu")[1:-1]
        assert np.array_equal(d_out[y0:y0 + 64].cpu().numpy(), want)
    finally:
        api.set_stream(None)


# ------------------------------------------------------------------------------------------ blur fp32
@pytest.mark.parametrize("shape", [(64, 128), (35, 260), (3, 4), (8, 8), (19, 23), (3, 3), (50, 1002)])
def test_blurxy_f32_bit_exact(api, shape):
    rng = np.random.default_rng(6)
    a = rng.random(shape, dtype=np.float32)
    out = np.full((shape[0] - 2, shape[1] - 2), np.nan, np.float32)
    api.blurxy_f32_tiramisu(a, out)
    assert np.array_equal(out, oracle.blur_f32(a))


def test_blur_f32_full_size_strip(api):
    import torch
    N = 16384
    g = torch.Generator(device="cuda").manual_seed(7)
    d_in = torch.rand((N, N), dtype=torch.float32, device="cuda", generator=g)
    d_out = torch.empty((N - 2, N - 2), dtype=torch.float32, device="cuda")
    api.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        api.blurxy_f32_tiramisu(d_in, d_out)
        torch.cuda.synchronize()
    finally:
        api.set_stream(None)
    for y0 in (0, 7777, N - 2 - 48):
        strip = d_in[y0:y0 + 50].cpu().numpy()
        assert np.array_equal(d_out[y0:y0 + 48].cpu().numpy(), oracle.blur_f32(strip))


# ------------------------------------------------------------------------------------------ blurxy uint8
@pytest.mark.parametrize("name", ["a", "b"])
def test_blurxy_u8_reference_golden(api, name):
    """Against the reference's own rgb.png -> reference_blurxy.png crops (tests/golden)."""
    i = np.load(os.path.join(GOLD, f"blurxy_{name}_in.npy"))
    o = np.load(os.path.join(GOLD, f"blurxy_{name}_out.npy"))
    out = np.zeros_like(o)
    api.blurxy_tiramisu(i, out)
    assert np.array_equal(out, o)


@pytest.mark.parametrize("shape", [(3, 9, 9), (1, 40, 77), (3, 64, 128), (2, 10, 12)])
def test_blurxy_u8_vs_oracle(api, shape):
    rng = np.random.default_rng(8)
    a = rng.integers(0, 256, shape, dtype=np.uint8)
    out = np.zeros((shape[0], shape[1] - 8, shape[2] - 8), np.uint8)
    api.blurxy_tiramisu(a, out)
    assert np.array_equal(out, oracle.blurxy_u8(a))


def test_blurxy_u8_full_image_size(api):
    """The reference's image size (3 x 3520 x 2112) on synthetic bytes, vs the oracle."""
    rng = np.random.default_rng(9)
    a = rng.integers(0, 256, (3, 3520, 2112), dtype=np.uint8)
    out = np.zeros((3, 3512, 2104), np.uint8)
    api.blurxy_tiramisu(a, out)
    assert np.array_equal(out, oracle.blurxy_u8(a))


# ------------------------------------------------------------------------------------------ boxblur uint32
def test_boxblur_test160_single_rank(api):
    """tests/wrapper_test_160.cpp on the last rank (which fills its own halo rows): (sum9)/9 ramp."""
    rows, cols = 10, 10
    img = np.zeros((rows + 2, cols + 2), np.uint32)
    r = np.arange(rows + 2)[:, None]
    c = np.arange(cols + 2)[None, :]
    img[:] = r + c
    out = np.zeros((rows, cols), np.uint32)
    api.boxblur(img, out)
    i64 = img.astype(np.int64)
    want = sum(i64[dy:dy + rows, dx:dx + cols] for dy in range(3) for dx in range(3)) // 9
    assert np.array_equal(out, want.astype(np.uint32))
    assert np.array_equal(out, oracle.boxblur_u32(img))


@pytest.mark.parametrize("rows,cols", [(100, 1000), (1, 1), (7, 300)])
def test_boxblur_vs_oracle(api, rows, cols):
    rng = np.random.default_rng(10)
    img = rng.integers(0, 2 ** 32, (rows + 2, cols + 2), dtype=np.uint32)    # sums wrap mod 2^32
    out = np.zeros((rows, cols), np.uint32)
    api.blurautodist_tiramisu(img, out)
    assert np.array_equal(out, oracle.boxblur_u32(img))
