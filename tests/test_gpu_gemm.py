"""GPU parity of the GEMM path (sgemm_tiramisu, matmul, tiramisu_cublas_*gemm and the reference's
cublas_gemm / gpu_tile feature tests) against the CPU oracle — through the C ABI."""
import numpy as np
import pytest

from oracle import oracle
from tests import refdata
from tests.test_oracle import CASES_164

pytestmark = pytest.mark.gpu

TF32_RTOL = 1e-2          # tolerance of the TF32 tensor-core paths


@pytest.fixture(scope="module")
def api():
    from tiramisu_b200 import api as _api
    return _api


@pytest.fixture()
def exact(api):
    prev = api.set_gemm_mode(1)
    yield
    api.set_gemm_mode(prev)


@pytest.fixture()
def tensor(api):
    prev = api.set_gemm_mode(2)
    yield
    api.set_gemm_mode(prev)


def _sgemm(api, N, M, K, a, b, A, B, Cm):
    out = Cm.copy()
    api.sgemm_tiramisu(np.array([N, M, K], np.int32), np.array([a], np.float32), np.array([b], np.float32),
                       A, B, out)
    return out


def _rel_err(got, want):
    return float(np.abs(got - want).max() / max(1e-30, np.abs(want).max()))


# ------------------------------------------------------------------------------ exact fp32 path
@pytest.mark.parametrize("N,M,K", [(64, 64, 64), (100, 50, 32), (1, 2, 3), (129, 67, 45), (256, 256, 256)])
def test_sgemm_exact_mode_bit_exact_vs_oracle(api, exact, N, M, K):
    """Sequential-k, no-FMA SIMT path == the reference x86 order (oracle), bit for bit."""
    rng = np.random.default_rng(20)
    A = rng.uniform(-1, 1, (N, K)).astype(np.float32)
    B = rng.uniform(-1, 1, (K, M)).astype(np.float32)
    Cm = rng.uniform(-1, 1, (N, M)).astype(np.float32)
    got = _sgemm(api, N, M, K, 3.0, 2.0, A, B, Cm)
    assert np.array_equal(got, oracle.sgemm(3.0, 2.0, A, B, Cm))


# ------------------------------------------------------------------------------ tensor-core path
@pytest.mark.parametrize("n", [128, 512, 1024])
def test_sgemm_all_ones_known_answer(api, tensor, n):
    """sgemm_wrapper.cpp:24,164-169: a=b=3, A=B=C=1 => 3K+3 exactly (integers are exact in TF32)."""
    A = np.ones((n, n), np.float32)
    got = _sgemm(api, n, n, n, 3.0, 3.0, A, A, A)
    assert (got == 3 * n + 3).all()


@pytest.mark.parametrize("N,M,K", [(128, 128, 128), (256, 512, 384), (1024, 1024, 1024), (136, 72, 100),
                                   (4, 8, 4), (300, 260, 36), (128, 64, 32)])
def test_sgemm_tf32_small_int_exact(api, tensor, N, M, K):
    """Integer-valued operands (the reference's own GPU check, sgemm/gpu/wrapper.cpp:51-55,134-150,
    compares with != against cuBLAS): TF32 products and fp32 sums are exact."""
    refdata.srand(0)
    A = refdata.rand_small_ints((N, K))
    B = refdata.rand_small_ints((K, M))
    Cm = refdata.rand_small_ints((N, M))
    got = _sgemm(api, N, M, K, 3.0, 2.0, A, B, Cm)
    assert np.array_equal(got, oracle.sgemm(3.0, 2.0, A, B, Cm))


@pytest.mark.parametrize("N,M,K", [(1024, 1024, 1024), (384, 640, 2048), (132, 260, 72)])
def test_sgemm_tf32_random_within_tolerance(api, tensor, N, M, K):
    rng = np.random.default_rng(5)
    A = rng.uniform(-1, 1, (N, K)).astype(np.float32)
    B = rng.uniform(-1, 1, (K, M)).astype(np.float32)
    Cm = rng.uniform(-1, 1, (N, M)).astype(np.float32)
    got = _sgemm(api, N, M, K, 3.0, 2.0, A, B, Cm)
    want = oracle.sgemm(3.0, 2.0, A, B, Cm)
    assert _rel_err(got, want) < TF32_RTOL


# ------------------------------------------------------------------------------ partly filled last wave: split tiles / stream-K
@pytest.mark.parametrize("mode,kernel,N,M,K,beta", [
    # mode 2 (default): k-aligned split of the remainder tiles between 2..4 clusters
    ("2", 3, 4096, 4096, 1024, 2.0), ("2", 3, 2304, 2560, 2048, 0.0), ("2", 3, 1000, 900, 4104, 2.0),
    ("2", 2, 1024, 2048, 2048, 2.0), ("2", 2, 1000, 1000, 2056, 0.0), ("2", 3, 4096, 4096, 512, 1.0),
    # mode 1: classic stream-K, equal contiguous runs (measured slower, kept for the record)
    ("1", 3, 2048, 2048, 2048, 2.0), ("1", 3, 2304, 2560, 1024, 0.0), ("1", 3, 4096, 4096, 512, 2.0),
    ("1", 3, 2000, 1900, 1000, 2.0), ("1", 2, 2048, 2048, 1024, 2.0), ("1", 2, 1500, 2300, 2056, 0.0),
    ("1", 0, 2048, 2048, 2048, 1.0), ("1", 3, 2048, 2048, 96, 2.0)])
def test_sgemm_split_last_wave_small_int_exact(api, tensor, monkeypatch, mode, kernel, N, M, K, beta):
    """Persistent kernels (2: 128x256 tiles, 3: CTA pairs) when the tiles do not fill the last wave: the k-slabs of the
    remainder tiles are shared between clusters, the tile's owner sums the others' raw accumulators in its epilogue
    through the workspace (gemm_tc.cu: GemmParams::sk_*).  Integer operands: every partial sum is exact, so the
    result must equal the plain product bit for bit whatever the split; twice in a row (the flags reset themselves),
    ragged tiles, K tails, beta = 0 and a K too short to split (falls back to whole tiles)."""
    monkeypatch.setenv("TB200_GEMM_STREAMK", mode)
    prev = api.set_gemm_kernel(kernel)
    try:
        rng = np.random.default_rng(N + K)
        A = rng.integers(-3, 4, (N, K)).astype(np.float32)
        B = rng.integers(-3, 4, (K, M)).astype(np.float32)
        Cm = rng.integers(-3, 4, (N, M)).astype(np.float32)
        want = (3.0 * (A.astype(np.float64) @ B.astype(np.float64)) + beta * Cm).astype(np.float32)
        for _ in range(2):
            got = _sgemm(api, N, M, K, 3.0, beta, A, B, Cm)
            assert np.array_equal(got, want)
    finally:
        api.set_gemm_kernel(prev)


# ------------------------------------------------------------------------------ every tcgen05 kernel variant
@pytest.mark.parametrize("kernel", [1, 2, 3, 4])
@pytest.mark.parametrize("N,M,K", [(256, 256, 32), (512, 768, 256), (1024, 1024, 1024), (300, 700, 96),
                                   (136, 264, 100), (2048, 512, 160), (1000, 900, 776), (128, 128, 64)])
def test_sgemm_each_kernel_small_int_exact(api, tensor, kernel, N, M, K):
    """One-tile-per-CTA (1), persistent 128x256 (2), persistent CTA-pair 256x256 (3) and the 2-CTA split-K cluster
    kernel (4: two K halves combined through distributed shared memory; a K of one slab falls back to 1) all
    give the oracle's result bit for bit on integer operands, including ragged tiles and K tails."""
    prev = api.set_gemm_kernel(kernel)
    try:
        refdata.srand(kernel)
        A = refdata.rand_small_ints((N, K))
        B = refdata.rand_small_ints((K, M))
        Cm = refdata.rand_small_ints((N, M))
        got = _sgemm(api, N, M, K, 3.0, 2.0, A, B, Cm)
        assert np.array_equal(got, oracle.sgemm(3.0, 2.0, A, B, Cm))
    finally:
        api.set_gemm_kernel(prev)


@pytest.mark.parametrize("kernel", [2, 3])
def test_sgemm_persistent_many_tiles_per_cta(api, tensor, kernel):
    """More tiles than SMs: every CTA loops (both TMEM accumulator buffers, ring wrap-around)."""
    import torch
    prev = api.set_gemm_kernel(kernel)
    try:
        M, N, K = 4096, 4608, 96
        g = torch.Generator(device="cuda").manual_seed(7)
        A = torch.randint(-4, 5, (M, K), device="cuda", generator=g).float()
        B = torch.randint(-4, 5, (K, N), device="cuda", generator=g).float()
        Cm = torch.randint(-4, 5, (M, N), device="cuda", generator=g).float()
        want = 3.0 * (A.double() @ B.double()) + 2.0 * Cm.double()
        api.set_stream(torch.cuda.current_stream().cuda_stream)
        try:
            api.cublas_sgemm(A.data_ptr(), B.data_ptr(), Cm.data_ptr(), M, N, K, 3.0, 2.0)
            torch.cuda.synchronize()
        finally:
            api.set_stream(None)
        assert bool((Cm.double() == want).all())
    finally:
        api.set_gemm_kernel(prev)


def test_matmul_gpu_wrapper_check(api):
    """sgemm/gpu/wrapper.cpp: Consts={3,2}, rand()%10-5 operands, exact comparison; times recorded."""
    refdata.srand(0)
    M = N = K = 768
    A = refdata.rand_small_ints((M, K))
    B = refdata.rand_small_ints((K, N))
    Cm = refdata.rand_small_ints((M, N))
    want = oracle.sgemm(3.0, 2.0, A, B, Cm)
    t0, t1 = np.zeros(1, np.float32), np.zeros(1, np.float32)
    got = Cm.copy()
    api.matmul(np.array([3, 2], np.float32), A, B, got, t0, t1)
    assert np.array_equal(got, want)
    assert t1[0] >= t0[0] > 0


# ------------------------------------------------------------------------------ tiramisu_cublas_sgemm contract
def _run_164(api, case):
    M, N, K, alpha, beta, rA, cA, rB, cB, rC, cC, oA, oB, oC, tA, tB = case
    A = refdata.rand_small_ints((rA, cA))
    B = refdata.rand_small_ints((rB, cB))
    Cm = refdata.rand_small_ints((rC, cC))
    want = oracle.gemm_ld(A, B, Cm, M, N, K, alpha, beta, cA, cB, cC, oA, oB, oC, tA, tB).reshape(rC, cC)
    got = Cm.copy()
    sizes = np.array([M, N, K, rA, cA, rB, cB, rC, cC, oA, oB, oC], np.int32)
    api.test_164(sizes, np.array([alpha, beta], np.float32), np.array([tA, tB], np.bool_), A, B, got)
    return got, want


@pytest.mark.parametrize("name", list(CASES_164))
def test_164_ld_offset_transpose(api, name):
    """tests/wrapper_test_164.cpp:70-125 (exact compare; integer-valued)."""
    refdata.srand(164)
    got, want = _run_164(api, CASES_164[name])
    assert np.array_equal(got, want)


@pytest.mark.parametrize("M,N,K,alpha,beta", [(100, 50, 32, 3, 4), (1, 2, 3, 4, 7), (128, 128, 128, 3, -1),
                                              (100, 100, 100, 1, 0), (100, 100, 100, 0, 1)])
def test_162_cublas_gemm(api, M, N, K, alpha, beta):
    """tests/wrapper_test_162.cpp:44-48."""
    refdata.srand(162)
    A = refdata.rand_small_ints((M, K))
    B = refdata.rand_small_ints((K, N))
    Cm = refdata.rand_small_ints((M, N))
    want = oracle.gemm_ld(A, B, Cm, M, N, K, alpha, beta).reshape(M, N)
    got = Cm.copy()
    api.test_162(np.array([M, N, K], np.int32), np.array([alpha, beta, 0], np.float32), A, B, got)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("M,N,K", [(100, 50, 32), (1, 1, 1), (162, 1, 113), (16, 100, 13), (2, 2, 1000),
                                   (256, 384, 128)])
def test_165_B_transposed(api, M, N, K):
    """tests/wrapper_test_165.cpp:27-35: C = A * B^T (B stored N x K), device pointers."""
    import torch
    refdata.srand(165)
    A = refdata.rand_small_ints((M, K))
    B = refdata.rand_small_ints((N, K))
    want = (A.astype(np.int64) @ B.astype(np.int64).T).astype(np.float32)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    dC = torch.full((M, N), float("nan"), device="cuda")
    api.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        api.cublas_sgemm(dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), M, N, K, 1.0, 0.0, tB=True)
        torch.cuda.synchronize()
    finally:
        api.set_stream(None)
    assert np.array_equal(dC.cpu().numpy(), want)


@pytest.mark.parametrize("tB", [False, True])
def test_sgemm_splitk_default_for_config0_shape(api, tensor, tB):
    """BASELINE configs[0] (1024^3) takes the split-K cluster kernel by default (kernel 0 = auto): random data
    within the TF32 tolerance, element-wise against |A| |B| (the scale a dot product's rounding is relative to),
    both B layouts, launched back to back (programmatic dependent launch) on rotating outputs."""
    import torch
    rng = np.random.default_rng(12)
    n = 1024
    A = rng.uniform(-1, 1, (n, n)).astype(np.float32)
    B = rng.uniform(-1, 1, (n, n)).astype(np.float32)
    C0 = rng.uniform(-1, 1, (n, n)).astype(np.float32)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    outs = [torch.from_numpy(C0).cuda() for _ in range(4)]
    launches0 = api.kernel_launches()
    api.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        for dC in outs:
            api.cublas_sgemm(dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), n, n, n, 3.0, 2.0, tB=tB)
        torch.cuda.synchronize()
    finally:
        api.set_stream(None)
    assert api.kernel_launches() - launches0 == 4
    Bm = B.T if tB else B
    want = 3.0 * (A.astype(np.float64) @ Bm.astype(np.float64)) + 2.0 * C0
    scale = 3.0 * (np.abs(A).astype(np.float64) @ np.abs(Bm).astype(np.float64)) + 2.0 * np.abs(C0)
    for dC in outs:
        got = dC.cpu().numpy()
        assert (np.abs(got - want) <= TF32_RTOL * scale).all()
        assert np.array_equal(got, outs[0].cpu().numpy())          # deterministic: fixed summation order


def test_165_B_transposed_tensor_path_random(api, tensor):
    """K-major B on the tcgen05 path (the LSTM weight layout), random data within TF32 tolerance."""
    import torch
    rng = np.random.default_rng(6)
    M, N, K = 64, 4096, 1024
    A = rng.uniform(-1, 1, (M, K)).astype(np.float32)
    B = rng.uniform(-1, 1, (N, K)).astype(np.float32)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    dC = torch.zeros((M, N), device="cuda")
    api.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        api.cublas_sgemm(dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), M, N, K, 1.0, 0.0, tB=True)
        torch.cuda.synchronize()
    finally:
        api.set_stream(None)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    assert _rel_err(dC.cpu().numpy(), want) < TF32_RTOL


def test_167_gpu_tile_gemm(api):
    """tests/wrapper_test_167.cpp:12-31: C += A*B, 128^3, exact."""
    refdata.srand(167)
    A = refdata.rand_small_ints((128, 128))
    B = refdata.rand_small_ints((128, 128))
    Cm = refdata.rand_small_ints((128, 128))
    want = (Cm.astype(np.int64) + A.astype(np.int64) @ B.astype(np.int64)).astype(np.float32)
    got = Cm.copy()
    api.test_167(A, B, got)
    assert np.array_equal(got, want)


def test_168_transpose(api):
    refdata.srand(168)
    A = refdata.rand_small_ints((128, 256))
    Bt = np.zeros((256, 128), np.float32)
    api.test_168(A, Bt)
    assert np.array_equal(Bt, oracle.transpose_f32(A))


def test_197_allocate_fill(api):
    """tests/wrapper_test_197.cpp: A := 1, B := 2 (int32, bit-exact)."""
    rng = np.random.default_rng(197)
    A = rng.integers(-20, -10, (2, 512, 512)).astype(np.int32)
    B = np.full((2, 512, 512), -20, np.int32)
    api.test_197(A, B)
    assert (A == 1).all() and (B == 2).all()


def test_dgemm_166(api):
    """tests/test_166: cublas_gemm in fp64 with transposes/offsets."""
    import torch
    rng = np.random.default_rng(166)
    M, N, K = 51, 17, 21
    A = rng.integers(-5, 5, (100, 200)).astype(np.float64)
    B = rng.integers(-5, 5, (222, 131)).astype(np.float64)
    Cm = rng.integers(-5, 5, (100, 123)).astype(np.float64)
    offA, offB, offC = 200 * 4 + 3, 131 * 5 + 6, 123 * 4 + 23
    want = oracle.gemm_ld(A, B, Cm, M, N, K, 75.0, 15.0, 200, 131, 123, offA, offB, offC, True, True,
                          dtype=np.float64).reshape(100, 123)
    dA, dB, dC = (torch.from_numpy(v).cuda() for v in (A, B, Cm))
    api.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        api.cublas_dgemm(dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), M, N, K, 75.0, 15.0, 200, 131, 123, offA,
                         offB, offC, True, True)
        torch.cuda.synchronize()
    finally:
        api.set_stream(None)
    assert np.array_equal(dC.cpu().numpy(), want)


def test_matmul_u8_tutorial_04gpu(api):
    import torch
    from tiramisu_b200 import _lib
    rng = np.random.default_rng(4)
    A = rng.integers(0, 256, (100, 100), dtype=np.uint8)
    B = rng.integers(0, 256, (100, 100), dtype=np.uint8)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    dC = torch.zeros((100, 100), dtype=torch.uint8, device="cuda")
    api.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        _lib.check(_lib.lib().tiramisu_b200_matmul_u8(dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), 100))
        torch.cuda.synchronize()
    finally:
        api.set_stream(None)
    assert np.array_equal(dC.cpu().numpy(), oracle.matmul_u8(A, B))
