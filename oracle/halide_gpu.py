"""numpy restatements of the reference's `*gpu` Halide benchmarks — TEST INFRASTRUCTURE ONLY.

Each function restates the Layer-I expression of one generator under /root/reference/benchmarks/halide/, in the
generator's own operand order, element types and wrap-around rules (uint8 arithmetic wraps mod 256, float32
products are rounded before they are added: the lowering compiles with -fmad=false), on arrays in the generated
function's layouts (images are planar `u8[C][H][W]`: `{3, M, N}` with N = width the unit-stride extent).
Only tests/ may import this module; the product never does.  Parity is pinned by construction only (the
reference checks these benchmarks against Halide pipelines that cannot be built here) — the restatements are
small enough to be compared with the generator text line by line.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def convolutiongpu(img: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    """convolutiongpu_tiramisu.cpp:46-56: out[c][y][x] = u8(sum over _j (rows) then _i (columns) of
    float(in[c][y+_j][x+_i]) * kernel(_i, _j)), e = 0 + t00 + t10 + t20 + t01 + ... (left to right), kernel(_i, _j) =
    kernel_gpu[_j][_i] (:72) = the caller's kernel_host[_j][_i].  img u8[3][M][N] -> u8[3][M-2][N-2]."""
    C, M, N = img.shape
    e = np.zeros((C, M - 2, N - 2), f32)
    x = img.astype(f32)
    for _j in range(3):
        for _i in range(3):
            e = e + x[:, _j:_j + M - 2, _i:_i + N - 2] * f32(kernel[_j][_i])
    return e.astype(np.uint8)                    # values stay inside [0, 255] in the tests: float -> u8 truncates


def gaussiangpu(img: np.ndarray, kx: np.ndarray, ky: np.ndarray) -> np.ndarray:
    """gaussiangpu_tiramisu.cpp:44-55: gaussian_x (inlined) = sum_k float(in[c][y][x+k]) * kx[k]; gaussian_y =
    u8(sum_k gaussian_x(x, y+k) * ky[k]); both sums start from 0.f and add left to right; output extents
    (M-5) x (N-5) (:36-37).  img u8[C][M][N] -> u8[C][M-5][N-5]."""
    C, M, N = img.shape
    x = img.astype(f32)
    NB, MB = N - 5, M - 5
    gx = np.zeros((C, M, NB), f32)
    for k in range(5):
        gx = gx + x[:, :, k:k + NB] * f32(kx[k])
    e = np.zeros((C, MB, NB), f32)
    for k in range(5):
        e = e + gx[:, k:k + MB, :] * f32(ky[k])
    return e.astype(np.uint8)


def edgegpu(img: np.ndarray) -> tuple[np.ndarray, tuple[slice, slice]]:
    """edgegpu_tiramisu.cpp:23-29 (function `edge_tiramisu`, fixed 8192 x 8192 x 3, in place):
    R = (ring of 8 neighbours, uint8 sums wrap) / 8 over x < N-2, y < M-2;
    Out[c][y][x] = (R(x+1,y+1) - R(x+2,y)) + (R(x+2,y+1) - R(x+1,y)), uint8, stored back into the image.
    R is a temporary that is only written for x < N-2, y < M-2 while Out reads R(x+2, y+1): the last two columns and
    the last row of Out depend on uninitialised memory in the reference too — the second return value is the region
    of the result that is defined.  img u8[3][M][N] -> (u8[3][M][N] with the defined region filled, region)."""
    C, M, N = img.shape
    a = img                                             # uint8: numpy wraps like the generated code
    Img = lambda dx, dy: a[:, dy:dy + M - 2, dx:dx + N - 2]   # noqa: E731   Img(i+dx, j+dy, c)
    R = (Img(0, 0) + Img(0, 1) + Img(0, 2) + Img(1, 0) + Img(1, 2) + Img(2, 0) + Img(2, 1) + Img(2, 2)) // np.uint8(8)
    h, w = M - 2 - 1, N - 2 - 2                          # Out(y, x) defined for y+1 < M-2, x+2 < N-2
    r = lambda dx, dy: R[:, dy:dy + h, dx:dx + w]        # noqa: E731
    out = img.copy()
    out[:, :h, :w] = (r(1, 1) - r(2, 0)) + (r(2, 1) - r(1, 0))
    return out, (slice(0, h), slice(0, w))


def fusiongpu(img: np.ndarray):
    """fusiongpu_tiramisu.cpp:36-41: f = 255 - in, g = 2 * in, h = f + g, k = f - g, all uint8 (wrapping)."""
    f = np.uint8(255) - img
    g = np.uint8(2) * img
    return f, g, f + g, f - g


def cvtcolorgpu(img: np.ndarray) -> np.ndarray:
    """cvtcolorgpu_tiramisu.cpp:41-55: u8((R*4899 + G*9617 + B*1868 + (1 << 13)) >> 14) in uint32.  u8[3][M][N] -> u8[M][N]."""
    x = img.astype(np.uint32)
    v = x[0] * np.uint32(4899) + x[1] * np.uint32(9617) + x[2] * np.uint32(1868)
    return ((v + (np.uint32(1) << np.uint32(13))) >> np.uint32(14)).astype(np.uint8)


def rgbyuv420gpu(img: np.ndarray):
    """rgbyuv420gpu_tiramisu.cpp:38-52 as written (`% 256`, Euclidean like Halide's Mod, where the usual formula has
    `>> 8`): y over the full image, u and v on the even pixels; `74 * copy_g` and `94 * copy_g` are uint8 products
    (they wrap) before the cast to int32.  u8[3][M][N] -> y u8[M][N], u, v u8[M/2][N/2] (even M, N)."""
    r, g, b = (img[c].astype(np.int32) for c in range(3))
    y = (((66 * r + 129 * g + 25 * b + 128) % 256) + 16).astype(np.uint8)
    sr, sb = r[::2, ::2], b[::2, ::2]
    sg8 = img[1][::2, ::2]
    g74 = (np.uint8(74) * sg8).astype(np.int32)
    g94 = (np.uint8(94) * sg8).astype(np.int32)
    u = (((-38 * sr - g74 + 112 * sb + 128) % 256) + 128).astype(np.uint8)
    v = (((112 * sr - g94 - 18 * sb + 128) % 256) + 128).astype(np.uint8)
    return y, u, v


def warp_affinegpu(img: np.ndarray) -> np.ndarray:
    """benchmarks/halide/warp_affinegpu_tiramisu.cpp:43-86, Layer I as written (fp32, separate multiplies and adds, every
    coefficient 0.1f).  img: in_host u8[M][N] (in(i, j) is img[j][i]); result out_host f32[M][N].  Two quirks of the reference
    are kept: coord_c is floor(o_r) (:66, not o_c), and clamp() lets an index reach N / M (:18-20) — callers keep
    0.1 (N + M) + 2 below min(N, M) so that no read leaves the image."""
    M, N = img.shape
    f = np.float32
    a = f(0.1)
    i = np.arange(N, dtype=np.int32)[None, :].astype(f)
    j = np.arange(M, dtype=np.int32)[:, None].astype(f)
    o_r = ((a * i).astype(f) + (a * j).astype(f)).astype(f) + a
    o_c = ((a * i).astype(f) + (a * j).astype(f)).astype(f) + a
    r = (o_r - np.floor(o_r)).astype(f)
    c = (o_c - np.floor(o_c)).astype(f)
    coord_r = np.floor(o_r).astype(np.int32)
    coord_c = np.floor(o_r).astype(np.int32)

    def clamp(n, hi):
        return np.where(n < 0, 0, np.where(n > hi, hi, n))

    def at(rr, cc):                                      # in(rr, cc) -> in_gpu[cc][rr]
        return img[cc, rr].astype(f)

    A00 = at(clamp(coord_r, N), clamp(coord_c, M))
    A01 = at(clamp(coord_r, N), clamp(coord_c + 1, M))
    A10 = at(clamp(coord_r + 1, N), clamp(coord_c, M))
    A11 = at(clamp(coord_r + 1, N), clamp(coord_c + 1, M))

    def mixf(x, y, t):
        return ((x * (f(1.0) - t).astype(f)).astype(f) + (y * t).astype(f)).astype(f)

    return mixf(mixf(A00, A01, r), mixf(A10, A11, r), c)
