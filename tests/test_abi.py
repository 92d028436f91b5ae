"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol that
include/*.h declares, and the halide_buffer_t mirror has the reference layout."""
import ctypes as C
import glob
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from tiramisu_b200 import build
    return build.build()


def _declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        text = re.sub(r"//[^\n]*", "", text)
        for m in re.finditer(r"^[A-Za-z_][\w \*]*?[\s\*]+([A-Za-z_]\w*)\s*\(", text, flags=re.M):
            names.append(m.group(1))
    return sorted(set(names))


def test_header_declares_the_reference_symbols():
    names = _declared_symbols()
    for must in ("tiramisu_cuda_malloc", "tiramisu_cuda_free", "tiramisu_cuda_memcpy_to_device",
                 "tiramisu_cuda_memcpy_to_host", "tiramisu_cublas_sgemm", "tiramisu_cublas_dgemm",
                 "tiramisu_cuda_stream_synchronize", "tiramisu_address_of_float32", "heat2d_tiramisu",
                 "heat2dgpu_tiramisu", "blurxy_tiramisu", "boxblur", "sgemm_tiramisu", "matmul",
                 "conv_relu_maxpool_block", "lstm"):
        assert must in names, must


def test_library_exports_every_declared_symbol(libpath):
    out = subprocess.check_output(["nm", "-D", "--defined-only", libpath], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [n for n in _declared_symbols() if n not in exported]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_lstm_abi_shims_export_lstm(libpath):
    """The reference's LSTM/gpu_lib and LSTM/cpu_lib wrappers link against a symbol `lstm` with 6 arguments (the main
    library's `lstm` is LSTM/cpu's 5-argument one): each ABI has a one-function shared library of its own."""
    for name in ("lstm_gpu_lib", "lstm_cpu_lib"):
        so = os.path.join(os.path.dirname(libpath), f"libtiramisu_b200_{name}.so")
        assert os.path.exists(so), so
        out = subprocess.check_output(["nm", "-D", "--defined-only", so], text=True)
        assert any(line.split()[-1] == "lstm" and " T " in line for line in out.splitlines())
        undef = subprocess.check_output(["nm", "-D", "--undefined-only", so], text=True)
        assert f"tiramisu_b200_{name}" in undef


def test_library_is_sm100a_only(libpath):
    out = subprocess.check_output(["cuobjdump", "-lelf", libpath], text=True)
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_halide_buffer_layout():
    from tiramisu_b200._lib import HalideBuffer, HalideDimension
    assert C.sizeof(HalideBuffer) == 56                      # HalideRuntime.h:1434-1460 on x86-64
    assert HalideBuffer.host.offset == 16 and HalideBuffer.dim.offset == 40
    assert C.sizeof(HalideDimension) == 16


def test_buffer_dims_are_innermost_first():
    from tiramisu_b200.halide import Buffer
    a = np.zeros((3, 5, 7), np.float32)
    b = Buffer(a)
    assert b.raw.dimensions == 3
    assert [b.raw.dim[i].extent for i in range(3)] == [7, 5, 3]
    assert [b.raw.dim[i].stride for i in range(3)] == [1, 7, 35]
    assert b.raw.type.bits == 32 and b.raw.type.code == 2 and b.raw.host == a.ctypes.data


def test_address_of_matches_reference_semantics(libpath):
    """src/tiramisu_externs.cpp:8-54: &((T*)buf->host)[idx] — pure host arithmetic, no GPU."""
    from tiramisu_b200 import _lib
    from tiramisu_b200.halide import Buffer
    L = _lib.lib()
    a = np.arange(10, dtype=np.float32)
    b = Buffer(a)
    assert L.tiramisu_address_of_float32(b.ptr(), 3) == a.ctypes.data + 12
    d = np.arange(10, dtype=np.float64)
    assert L.tiramisu_address_of_float64(Buffer(d).ptr(), 2) == d.ctypes.data + 16
    u = np.arange(10, dtype=np.uint8)
    assert L.tiramisu_address_of_uint8(Buffer(u).ptr(), 7) == u.ctypes.data + 7


def test_validation_errors_return_codes_without_a_gpu(libpath):
    """Null / inconsistent arguments are rejected before any CUDA call (the reference would be UB)."""
    from tiramisu_b200 import _lib, api
    a = np.zeros((8, 8), np.float32)
    o = np.zeros((8, 9), np.float32)
    with pytest.raises(_lib.TiramisuError, match="extents differ"):
        api.heat2d_tiramisu(a, o)
    with pytest.raises(_lib.TiramisuError, match="32-bit float"):
        api.heat2d_tiramisu(a.astype(np.float64), a.astype(np.float64))
    with pytest.raises(_lib.TiramisuError, match="smaller than the stencil"):
        api.blurxy_tiramisu(np.zeros((3, 8, 8), np.uint8), np.zeros((3, 0, 0), np.uint8))
