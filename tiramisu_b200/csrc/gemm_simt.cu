// gemm_simt.cu — generic row-major GEMM with leading dimensions, element offsets and
// transposes: the full contract of tiramisu_cublas_sgemm / tiramisu_cublas_dgemm
// (/root/reference/src/tiramisu_cuda_wrappers.cpp:74-162).  This is the path for operands the
// tcgen05/TMA kernel cannot address (unaligned bases or leading dimensions, tiny sizes, fp64,
// uint8), and the "exact" mode of the fp32 GEMM: k is accumulated sequentially with separately
// rounded multiply and add, which is bit-identical to the reference's x86 code (no FMA,
// src/tiramisu_codegen_halide.cpp:3988-3995; k monotone per output,
// benchmarks/linear_algebra/blas/level3/sgemm/cpu/sgemm_generator.cpp:102).
//
// 64x64 output tile per 256-thread CTA, 4x4 register block per thread, BK=16 smem stages.
#include "common.cuh"
#include "tiramisu_b200.h"

using namespace tb200;

namespace {

template <class T> struct Arith;
template <> struct Arith<float> {
    static __device__ __forceinline__ float mac(float acc, float a, float b) { return __fadd_rn(acc, __fmul_rn(a, b)); }
    static __device__ __forceinline__ float epi(float acc, float alpha, float c, float beta, bool use_c)
    {
        float t0 = __fmul_rn(acc, alpha);
        return use_c ? __fadd_rn(t0, __fmul_rn(c, beta)) : t0;
    }
};
template <> struct Arith<double> {
    static __device__ __forceinline__ double mac(double acc, double a, double b) { return __dadd_rn(acc, __dmul_rn(a, b)); }
    static __device__ __forceinline__ double epi(double acc, double alpha, double c, double beta, bool use_c)
    {
        double t0 = __dmul_rn(acc, alpha);
        return use_c ? __dadd_rn(t0, __dmul_rn(c, beta)) : t0;
    }
};
template <> struct Arith<uint8_t> {   // tutorial_04gpu.cpp:54 — uint8 arithmetic wraps mod 256
    static __device__ __forceinline__ uint8_t mac(uint8_t acc, uint8_t a, uint8_t b) { return (uint8_t)(acc + (uint8_t)(a * b)); }
    static __device__ __forceinline__ uint8_t epi(uint8_t acc, uint8_t, uint8_t, uint8_t, bool) { return acc; }
};

constexpr int BM = 64, BN = 64, BK = 16;

template <class T>
__global__ void __launch_bounds__(256)
tiramisu_sm100_gemm_generic(const T *__restrict__ A, const T *__restrict__ B, T *__restrict__ C,
                            int64_t M, int64_t N, int64_t K, T alpha, T beta,
                            int64_t ldA, int64_t ldB, int64_t ldC, bool tA, bool tB, bool use_c)
{
    __shared__ T As[BK][BM + 4];
    __shared__ T Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t i0 = (int64_t)blockIdx.y * BM, j0 = (int64_t)blockIdx.x * BN;

    T acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = T(0);

    for (int64_t k0 = 0; k0 < K; k0 += BK) {
        // A tile: 64 x 16 elements, 4 per thread; walk the contiguous direction with the low tid bits
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int e = tid + r * 256;
            int ii, kk;
            if (tA) { ii = e & 63; kk = e >> 6; } else { kk = e & 15; ii = e >> 4; }
            const int64_t gi = i0 + ii, gk = k0 + kk;
            T v = T(0);
            if (gi < M && gk < K) v = tA ? A[gk * ldA + gi] : A[gi * ldA + gk];
            As[kk][ii] = v;
        }
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int e = tid + r * 256;
            int jj, kk;
            if (tB) { kk = e & 15; jj = e >> 4; } else { jj = e & 63; kk = e >> 6; }
            const int64_t gj = j0 + jj, gk = k0 + kk;
            T v = T(0);
            if (gj < N && gk < K) v = tB ? B[gj * ldB + gk] : B[gk * ldB + gj];
            Bs[kk][jj] = v;
        }
        __syncthreads();
        const int kmax = (K - k0 < (int64_t)BK) ? (int)(K - k0) : BK;
        for (int kk = 0; kk < kmax; kk++) {      // sequential k: bit-exact vs the x86 order
            T a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { a[u] = As[kk][ty * 4 + u]; b[u] = Bs[kk][tx * 4 + u]; }
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int w = 0; w < 4; w++) acc[u][w] = Arith<T>::mac(acc[u][w], a[u], b[w]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int64_t gi = i0 + ty * 4 + u;
        if (gi >= M) continue;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const int64_t gj = j0 + tx * 4 + w;
            if (gj >= N) continue;
            T *p = C + gi * ldC + gj;
            *p = Arith<T>::epi(acc[u][w], alpha, use_c ? *p : T(0), beta, use_c);
        }
    }
}

template <class T>
int launch_generic(const T *A, const T *B, T *C, uint64_t M, uint64_t N, uint64_t K, T alpha, T beta,
                   uint64_t ldA, uint64_t ldB, uint64_t ldC, bool tA, bool tB, bool use_c, const char *name)
{
    if (M == 0 || N == 0) return 0;
    dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM));
    if (grid.y > 65535) return fail(name, "M too large for the generic GEMM path");
    tiramisu_sm100_gemm_generic<T><<<grid, 256, 0, stream()>>>(A, B, C, (int64_t)M, (int64_t)N, (int64_t)K, alpha,
                                                                beta, (int64_t)ldA, (int64_t)ldB, (int64_t)ldC, tA,
                                                                tB, use_c);
    return post_launch(name);
}

}  // namespace

namespace tb200 {
// Exact fp32 path (used by tiramisu_cublas_sgemm when the tensor-core kernel cannot address the
// operands or exact mode is forced).  Pointers already include their offsets.
int gemm_f32_exact(const float *A, const float *B, float *C, uint64_t M, uint64_t N, uint64_t K,
                   float alpha, float beta, uint64_t ldA, uint64_t ldB, uint64_t ldC, bool tA, bool tB)
{
    // cuBLAS semantics (which the reference calls): C is not read when beta == 0.
    return launch_generic<float>(A, B, C, M, N, K, alpha, beta, ldA, ldB, ldC, tA, tB, beta != 0.0f,
                                 "tiramisu_sm100_gemm_generic<float>");
}
}  // namespace tb200

TB200_API int tiramisu_cublas_dgemm(double *A, double *B, double *C, uint64_t M, uint64_t N, uint64_t K,
                                    double alpha, double beta, uint64_t ldA, uint64_t ldB, uint64_t ldC,
                                    uint64_t offsetA, uint64_t offsetB, uint64_t offsetC,
                                    bool transposeA, bool transposeB)
{
    if (!A || !B || !C) return fail(__func__, "null operand");
    // src/tiramisu_cuda_wrappers.cpp:137-145: 0 means tightly packed
    if (ldA == 0) ldA = transposeA ? M : K;
    if (ldB == 0) ldB = transposeB ? K : N;
    if (ldC == 0) ldC = N;
    return launch_generic<double>(A + offsetA, B + offsetB, C + offsetC, M, N, K, alpha, beta, ldA, ldB, ldC,
                                  transposeA, transposeB, beta != 0.0, "tiramisu_sm100_gemm_generic<double>");
}

TB200_API int tiramisu_b200_matmul_u8(const uint8_t *A, const uint8_t *B, uint8_t *C, int32_t N)
{
    if (!A || !B || !C || N < 1) return fail(__func__, "bad argument");
    return launch_generic<uint8_t>(A, B, C, N, N, N, 1, 0, N, N, N, false, false, false,
                                   "tiramisu_sm100_gemm_generic<uint8>");
}
