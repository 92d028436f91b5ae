// compat_support.cpp — TEST INFRASTRUCTURE for tests/compat: the "reference implementation" halves
// that the reference's wrapper_*.cpp harnesses link besides the Tiramisu-generated function
// (Halide-compiled *_ref pipelines, MKL's cblas_sgemm).  They are provided here on top of the CPU
// oracle (oracle/libtiramisu_oracle.so) so that the UNMODIFIED harnesses can compare the B200
// backend against something on the box.  Never linked into the product library.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "HalideRuntime.h"
#include "mkl_cblas.h"

extern "C" {
void oracle_heat2d(const float *in, float *out, int H, int W);
void oracle_heat2dgpu(const float *in, float *out, int H, int W);

// benchmarks/halide/wrapper_heat2d.h:12 — the Halide reference pipeline of heat2d
int heat2d_ref(halide_buffer_t *in, halide_buffer_t *out)
{
    oracle_heat2d(reinterpret_cast<const float *>(in->host), reinterpret_cast<float *>(out->host), in->dim[1].extent,
                  in->dim[0].extent);
    return 0;
}

// benchmarks/halide/wrapper_heat2dgpu.cpp:14
int heat2dgpu_ref(halide_buffer_t *in, halide_buffer_t *out)
{
    oracle_heat2dgpu(reinterpret_cast<const float *>(in->host), reinterpret_cast<float *>(out->host), in->dim[1].extent,
                     in->dim[0].extent);
    return 0;
}

// wrapper_heat2dgpu.cpp brings the Halide-GPU reference's output back with the Halide runtime; our
// heat2dgpu_ref computes on the host, so these are no-ops.
int halide_copy_to_host(void *, halide_buffer_t *) { return 0; }
int halide_device_free(void *, halide_buffer_t *) { return 0; }

// Row-major, no-transpose sgemm (the only form sgemm_wrapper.cpp calls): C = alpha*A*B + beta*C.
// Cache-blocked, OpenMP, AVX via the compiler's vectoriser; fp32 accumulation in k order per (i,j)
// block, which is exact for the wrapper's all-ones operands.
void cblas_sgemm(CBLAS_LAYOUT layout, CBLAS_TRANSPOSE transA, CBLAS_TRANSPOSE transB, long long M, long long N,
                 long long K, float alpha, const float *A, long long lda, const float *B, long long ldb, float beta,
                 float *C, long long ldc)
{
    if (layout != CblasRowMajor || transA != CblasNoTrans || transB != CblasNoTrans) __builtin_trap();
    const long long BI = 64, BJ = 256, BKK = 256;
#pragma omp parallel for collapse(2) schedule(static)
    for (long long i0 = 0; i0 < M; i0 += BI)
        for (long long j0 = 0; j0 < N; j0 += BJ) {
            const long long i1 = std::min(i0 + BI, M), j1 = std::min(j0 + BJ, N);
            std::vector<float> acc((size_t)(i1 - i0) * (j1 - j0), 0.0f);
            for (long long k0 = 0; k0 < K; k0 += BKK) {
                const long long k1 = std::min(k0 + BKK, K);
                for (long long i = i0; i < i1; i++) {
                    float *arow = acc.data() + (size_t)(i - i0) * (j1 - j0);
                    for (long long k = k0; k < k1; k++) {
                        const float a = A[i * lda + k];
                        const float *brow = B + k * ldb + j0;
                        for (long long j = 0; j < j1 - j0; j++) arow[j] += a * brow[j];
                    }
                }
            }
            for (long long i = i0; i < i1; i++)
                for (long long j = j0; j < j1; j++)
                    C[i * ldc + j] = alpha * acc[(size_t)(i - i0) * (j1 - j0) + (j - j0)] + beta * C[i * ldc + j];
        }
}
}  // extern "C"
