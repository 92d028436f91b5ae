// mkl_cblas.h — TEST INFRASTRUCTURE: the slice of the CBLAS interface that
// benchmarks/linear_algebra/blas/level3/sgemm/cpu/sgemm_wrapper.cpp uses for its MKL reference run;
// implemented in compat_support.cpp on top of the CPU oracle (MKL is not in the image).
#pragma once
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_LAYOUT;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
#ifdef __cplusplus
extern "C" {
#endif
void cblas_sgemm(CBLAS_LAYOUT layout, CBLAS_TRANSPOSE transA, CBLAS_TRANSPOSE transB, long long M, long long N,
                 long long K, float alpha, const float *A, long long lda, const float *B, long long ldb, float beta,
                 float *C, long long ldc);
#ifdef __cplusplus
}
#endif
