"""Host emulation of generated sources (tests/emulation) — CPU-side check of the LOWERING's semantics.

The container that builds the repo has no GPU.  These tests take the `.b200.cu` sources the generators wrote,
run every kernel body for every (block, thread) index on the host (tests/emulation/include/cuda_runtime.h) and
compare with the oracle / the reference's own wrappers, so that an index expression, a guard, a host loop or a
device-mirror copy that the lowering gets wrong fails HERE and not only at the next GPU run.  This is test
infrastructure: nothing of it is part of the product, the library kernels are not involved (the runtime entry
points are host stand-ins), races between threads cannot be seen, and kernels that need barriers / shared
memory / shuffles are not emulatable.  The -m gpu tests of test_frontend.py remain the parity gate.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from tests.compat import build_generators as bg
from tests.emulation import emulate
from tests.frontend_cases import build_cases as bc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cases():
    from tiramisu_b200 import build
    build.build()
    build.build_frontend()
    res = bc.build()
    failed = {k: v[1] for k, v in res.items() if v[0] != 0}
    assert not failed, failed
    return res


def _emulated(name):
    return C.CDLL(emulate.build_shared([bc.generated_source(name)], name))


def _call(lib, fn, *arrays):
    from tiramisu_b200.halide import Buffer
    bufs = [Buffer(a) for a in arrays]
    f = getattr(lib, fn)
    f.restype = C.c_int
    assert f(*[b.ptr() for b in bufs]) == 0


def test_rewriter_refuses_what_a_thread_loop_cannot_emulate(cases, tmp_path):
    with pytest.raises(emulate.NotEmulatable):                       # warp-shuffle reduction
        emulate.rewrite(bc.generated_source("matvec_cpu_gen"), str(tmp_path / "x.cpp"))
    n = emulate.rewrite(bc.generated_source("heat2d_gen_generic"), str(tmp_path / "y.cpp"))
    assert n == 2 and "<<<" not in open(tmp_path / "y.cpp").read()


@pytest.mark.parametrize("variant", ["heat2d_gen", "heat2d_gen_generic"])
@pytest.mark.parametrize("shape", [(64, 64), (130, 260), (1, 5), (3, 3)])
def test_emulated_heat2d(cases, shape, variant):
    from oracle import oracle
    lib = _emulated(variant)
    a = np.random.default_rng(11).random(shape, dtype=np.float32)
    out = np.full(shape, -7.0, np.float32)
    _call(lib, variant, np.array(shape, np.int32), a, out)
    assert np.array_equal(out, oracle.heat2d(a, "cpu"))


@pytest.mark.parametrize("hw", [(16, 24), (67, 131), (9, 9)])
def test_emulated_blurxy_u8(cases, hw):
    from oracle import oracle
    lib = _emulated("blurxy_u8_gen")
    h, w = hw
    a = np.random.default_rng(12).integers(0, 256, (3, h, w), dtype=np.uint8)
    out = np.zeros((3, h - 8, w - 8), np.uint8)
    _call(lib, "blurxy_u8_gen", np.array([h, w], np.int32), a, out)
    assert np.array_equal(out, oracle.blurxy_u8(a))


@pytest.mark.parametrize("hw", [(16, 24), (67, 131), (3, 3)])
def test_emulated_retargeted_blur_f32(cases, hw):
    from oracle import oracle
    lib = _emulated("blur_f32_cpu_gen")
    h, w = hw
    a = np.random.default_rng(16).random((h, w), dtype=np.float32)
    out = np.full((h - 2, w - 2), -1.0, np.float32)
    _call(lib, "blur_f32_cpu_gen", np.array([h, w], np.int32), a, out)
    assert np.array_equal(out, oracle.blur_f32(a))


@pytest.mark.parametrize("variant", ["sgemm_cpu_sched_gen", "sgemm_cpu_sched_tc"])
@pytest.mark.parametrize("nmk", [(64, 64, 64), (100, 36, 52), (33, 130, 17)])
def test_emulated_sgemm_cpu_schedule(cases, nmk, variant):
    """Both lowerings of the sgemm statement triple: the three SIMT kernels (generic) and the single runtime GEMM
    call (pattern; here the stand-in evaluates it in fp32, k sequential) — the latter checks the call's operands,
    leading dimensions, alpha / beta wiring."""
    from oracle import oracle
    lib = _emulated(variant)
    n, m, k = nmk
    rng = np.random.default_rng(17)
    A = rng.uniform(-1, 1, (n, k)).astype(np.float32)
    B = rng.uniform(-1, 1, (k, m)).astype(np.float32)
    Cm = rng.uniform(-1, 1, (n, m)).astype(np.float32)
    got = Cm.copy()
    _call(lib, variant, np.array([n, m, k], np.int32), np.array([1.5], np.float32), np.array([-0.75], np.float32), A, B, got)
    assert np.array_equal(got, oracle.sgemm(1.5, -0.75, A, B, Cm.copy()))


@pytest.mark.parametrize("variant", ["gemm_sym_gen", "gemm_sym_generic"])
@pytest.mark.parametrize("mnk", [(64, 48, 40), (100, 36, 52)])
def test_emulated_contraction(cases, mnk, variant):
    lib = _emulated(variant)
    m, n, k = mnk
    rng = np.random.default_rng(15)
    A = rng.integers(-5, 5, (m, k)).astype(np.float32)
    Bt = rng.integers(-5, 5, (n, k)).astype(np.float32)
    Cm = rng.integers(-5, 5, (m, n)).astype(np.float32)
    got = Cm.copy()
    _call(lib, variant, np.array(mnk, np.int32), A, Bt, got)
    assert np.array_equal(got, Cm + A @ Bt.T)                       # small integers: exact in fp32 in any order


@pytest.mark.parametrize("nkm", [(64, 64, 64), (33, 17, 300), (7, 0, 5)])
def test_emulated_fused_reduction_grid_mapping(cases, nkm):
    lib = _emulated("fused_reduce_cpu_gen")
    n, k, m = nkm
    rng = np.random.default_rng(21)
    A = rng.uniform(-1, 1, (n, max(k, 1))).astype(np.float32)[:, :k].copy()
    B = rng.uniform(-1, 1, (max(k, 1), m)).astype(np.float32)[:k].copy()
    got = np.full((n, m), -9.0, np.float32)
    _call(lib, "fused_reduce_cpu_gen", np.array([n, k, m], np.int32), A, B, got)
    want = np.zeros((n, m), np.float32)
    for kk in range(k):
        want = want + A[:, kk:kk + 1] * B[kk][None, :]
    assert np.array_equal(got, want)


@pytest.mark.parametrize("mn", [(64, 64), (33, 17)])
def test_emulated_matvec_generic(cases, mn):
    lib = _emulated("matvec_cpu_generic")
    m, n = mn
    rng = np.random.default_rng(18)
    A = rng.integers(-4, 4, (m, n)).astype(np.float32)
    x = rng.integers(-4, 4, (n,)).astype(np.float32)
    y0 = rng.integers(-4, 4, (m,)).astype(np.float32)
    y = np.full((m,), -3.0, np.float32)
    _call(lib, "matvec_cpu_generic", np.array([m, n], np.int32), A, x, y0, y)
    assert np.array_equal(y, y0 + A @ x)


def test_emulated_advisor_cases_polymul_and_unequal_fuse_depths(cases):
    """The two miscompiles ADVICE.md (round 1) reproduced.  (1) polymul_cpu_gen: S(i, j) stored at b_O[i + j] — the
    generated kernel must keep BOTH loops in one thread (an emulated run cannot see a race, so the source is
    inspected: no blockIdx / threadIdx feeds the statement) and the result is the full convolution.
    (2) prefix_fuse_cpu_gen: S2 fused with S1 inside k reads the running sum; un-fusing the run would give K * total."""
    src = open(bc.generated_source("polymul_cpu_gen")).read()
    kern = src[src.index("__global__"):src.index('extern "C" int polymul_cpu_gen')]
    assert "threadIdx" not in kern and "__launch_bounds__(1)" in kern and kern.count("for (") == 2
    rng = np.random.default_rng(23)
    n, m = 300, 77
    A = rng.integers(-4, 5, n).astype(np.float32)
    B = rng.integers(-4, 5, m).astype(np.float32)
    got = np.zeros(n + m - 1, np.float32)
    _call(_emulated("polymul_cpu_gen"), "polymul_cpu_gen", np.array([n, m], np.int32), A, B, got)
    assert np.array_equal(got, np.convolve(A.astype(np.float64), B.astype(np.float64)).astype(np.float32))

    rng = np.random.default_rng(24)
    n, k = 37, 21
    A2 = rng.integers(-4, 5, (n, k)).astype(np.float32)
    got = np.full(n, -7.0, np.float32)
    _call(_emulated("prefix_fuse_cpu_gen"), "prefix_fuse_cpu_gen", np.array([n, k], np.int32), A2, got)
    assert np.array_equal(got, np.cumsum(A2.astype(np.float64), axis=1).sum(axis=1).astype(np.float32))


def test_floor_division_and_modulo_follow_the_reference(tmp_path):
    """Integer / and % in generated code are Halide's (floor division, non-negative remainder:
    src/tiramisu_codegen_halide.cpp:3563-3569), not C's truncation: the helpers the lowering emits for operands
    whose sign it cannot prove, compiled and run on the host."""
    import subprocess
    src = open(bc.generated_source("polymul_cpu_gen")).read()
    a = src.index("template <typename T> __host__ __device__ __forceinline__ T tb_floordiv")
    b = src.index("// Halide's lerp")
    helpers = src[a:b].replace("__host__ __device__ __forceinline__", "inline")
    prog = tmp_path / "fd.cpp"
    prog.write_text("#include <cstdio>\n" + helpers + """
int main() {
    int bad = 0;
    for (int a = -40; a <= 40; a++) for (int b = -7; b <= 7; b++) {
        if (b == 0) { bad += tb_floordiv<int>(a, b) != 0 || tb_floormod<int>(a, b) != 0; continue; }
        const int q = tb_floordiv<int>(a, b), r = tb_floormod<int>(a, b);
        bad += !(r >= 0 && r < (b < 0 ? -b : b) && q * b + r == a);
    }
    printf("%d\\n", bad);
    return bad != 0;
}
""")
    exe = tmp_path / "fd"
    subprocess.check_call(["g++", "-std=c++17", "-O1", str(prog), "-o", str(exe)], env={k: v for k, v in os.environ.items() if k not in ("CC", "CXX")})
    assert subprocess.run([str(exe)], capture_output=True, text=True).stdout.strip() == "0"


@pytest.mark.parametrize("variant", ["lstm_cpu_so", "lstm_cpu_so_generic"])
def test_emulated_reference_lstm_generator(variant):
    """The reference's LSTM/cpu/generator.cpp, compiled unmodified by the front-end (tests/compat/build_generators.py),
    run on the host against the oracle: statement order, the bias copy, the offset GEMM calls of the pattern lowering
    (here evaluated in fp32 by the stand-in) and the fused pointwise cell kernel."""
    from oracle import oracle
    src = os.path.join(ROOT, "tests/compat/_gen", variant, "lstm.o.b200.cu")
    if not os.path.exists(src):
        pytest.skip(f"{src} not generated (needs /root/reference at build time)")
    lib = C.CDLL(emulate.build_shared([src], variant))
    F, B, L, T = 32, 3, 2, 4
    rng = np.random.default_rng(20)
    Wt = (rng.uniform(-1, 1, (L, 2, 4 * F, F)) / 16).astype(np.float32)
    bias = rng.uniform(-1, 1, (L, 4 * F)).astype(np.float32)
    x = rng.uniform(-1, 1, (T, B, F)).astype(np.float32)
    y = np.zeros((T, B, F), np.float32)
    _call(lib, "lstm", np.array([F, B, L, T], np.int32), Wt, bias, x, y)
    want = oracle.lstm(Wt, bias, x)
    assert np.abs(y - want).max() <= 1e-5 * max(1.0, np.abs(want).max()), np.abs(y - want).max()


@pytest.mark.skipif(not os.path.isdir(os.path.join(bg.REF, "tests")), reason="needs /root/reference (wrappers)")
def test_reference_feature_tests_pass_their_wrappers_in_emulation():
    """Every reference test pinned as passing on the B200 must also pass when its generated source is run on the
    host: the wrapper's own comparison decides.  (Needs the sweep's generator outputs: __graft_entry__.build() or
    tests/compat/build_reference_tests.py.)"""
    expected = json.load(open(os.path.join(ROOT, "tests/compat/reference_tests_expected.json")))
    pinned = expected["pass"] + expected["host_emulation_only"]["tests"]
    have = [n for n in pinned if os.path.isdir(os.path.join(ROOT, "tests/compat/_gen/reftests", n))]
    if len(have) < len(pinned):
        pytest.skip("sweep outputs not built")
    # + the gpu-scheduled tests build_generators.py covers one by one
    have += [n for n in ("162", "164", "165", "166", "167", "168", "169", "170", "171", "197")
             if os.path.isdir(os.path.join(ROOT, "tests/compat/_gen", f"test_{n}"))]
    res = emulate.sweep(have)
    bad = {k: v for k, v in res.items() if v != "pass"}
    assert not bad, bad


@pytest.mark.skipif(not os.path.isdir(os.path.join(bg.REF, "tests")), reason="needs /root/reference (wrappers)")
def test_distributed_reference_tests_pass_their_wrappers_as_ten_emulated_ranks():
    """tests/test_98, 99, 101, 102 (explicit transfers, blocking; fan-out / fan-in; collapse_many) and test_160
    (gen_communication: the halo derived from the accesses) of the reference, generators and wrappers unmodified: the
    lowered functions run as 10 host processes over the file-mailbox stand-in of tiramisu_MPI_* and every rank's
    wrapper compares its buffers itself."""
    expected = json.load(open(os.path.join(ROOT, "tests/compat/reference_tests_expected.json")))
    ids = expected["distributed"]["tests"]
    have = [n for n in ids if os.path.isdir(os.path.join(ROOT, "tests/compat/_gen/reftests", n))]
    if len(have) < len(ids):
        pytest.skip("sweep outputs not built")
    res = dict(emulate.reference_dist_test(n, 10) for n in ids)
    bad = {k: v for k, v in res.items() if v != "pass"}
    assert not bad, bad


@pytest.mark.parametrize("world", [2, 4, 8])
def test_generated_distributed_functions_as_emulated_ranks(cases, world, tmp_path):
    """tests/frontend_cases/dist_blur_gen.cpp (gen_communication halo) and dist_xfer_gen.cpp (explicit fan-out / ring with
    blocking synchronous sends / hand-collapsed fan-in) at 2, 4 and 8 ranks: tools/dist_gen_check.py compares every rank's
    buffers with numpy restatements; here the ranks are host processes over the file-mailbox tiramisu_MPI_* stand-in, on
    the GPU box the same script runs under torchrun over NCCL (tests/test_gpu_multi.py)."""
    import subprocess
    import sys
    mpi = (os.path.join(ROOT, "tests", "emulation", "mpi_emu.cpp"),)
    blur = emulate.build_shared([bc.generated_source(f"dist_blur_gen_{world}")], f"dist_blur_gen_{world}", mpi)
    xfer = emulate.build_shared([bc.generated_source(f"dist_xfer_gen_{world}")], f"dist_xfer_gen_{world}", mpi)
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tools", "dist_gen_check.py"), "--emulated", blur, xfer],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                              env=dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), TB200_EMU_MPI_DIR=str(tmp_path)))
             for r in range(world)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 and "DIST_GEN_OK" in o for p, o in zip(procs, outs)), "\n".join(o[-800:] for o in outs)


def test_emulated_register_blocking_idioms(cases):
    """tests/frontend_cases/regblock_cpu_gen.cpp: a load / accumulate / store triple through a staging buffer (-> in-place
    accumulation), an init / accumulate / consume group (-> staging buffer expanded), a max-pool over y / 2, x / 2 (-> iterators
    cut), fused nests distributed: bit-exact against the statement-order fp32 restatement."""
    from tests import refdata
    (inp, w, bias), conv, pooled = refdata.regblock_case()
    out = np.full(conv.shape, 3.0, np.float32)
    pool = np.full(pooled.shape, 5.0, np.float32)
    _call(_emulated("regblock_cpu_gen"), "regblock_cpu_gen", inp, w, bias, out, pool)
    assert np.array_equal(out, conv) and np.array_equal(pool, pooled)


@pytest.mark.parametrize("nm", [(16, 32), (7, 21), (5, 9)])
def test_emulated_statement_pieces_and_after_low_level(cases, nm):
    """tests/frontend_cases/pieces_cpu_gen.cpp: full-tile + partial-tile pieces ordered with after_low_level and scheduled by
    loop level become one statement (sizes that are and are not multiples of the tile)."""
    n, m = nm
    A = np.random.default_rng(n * m).uniform(-1, 1, (n, m)).astype(np.float32)
    B = np.full((n, m), 9.0, np.float32)
    _call(_emulated("pieces_cpu_gen"), "pieces_cpu_gen", np.array([n, m], np.int32), A, B)
    assert np.array_equal(B, (A * np.float32(2.0) + np.float32(1.0)).astype(np.float32))


@pytest.mark.skipif(not os.path.isdir("/root/reference/benchmarks/automatic_comm"), reason="needs /root/reference")
@pytest.mark.parametrize("name", ["blurautodist", "edgeautodist", "cvtcolorautodist", "convolutionautodist"])
def test_emulated_reference_autodist_wrappers_16_ranks(name):
    """benchmarks/automatic_comm/<name>_tiramisu.cpp (blurautodist is row a8's own generator: the halo exchange derived by
    gen_communication) and <name>_ref.cpp (the hand-written exchange), both unmodified, with the reference's own
    wrapper_<name>.cpp (-DWITH_MPI, 16 ranks hard-coded, compare_buffers of the two outputs on every rank) as 16 host
    processes over the message stand-in."""
    cus = [os.path.join(bg.GEN, name + v, "build", f"generated_fct_{name}{v}.o.b200.cu") for v in ("", "_ref")]
    if not all(os.path.exists(p) for p in cus):
        bg.build([name, name + "_ref"])
    res = emulate.dist_wrapper_test(name, cus, f"benchmarks/automatic_comm/wrapper_{name}.cpp", ["benchmarks/automatic_comm", "benchmarks"], 16)
    assert res == "pass", res
