// tests/emulation/runtime_emu.cpp — TEST INFRASTRUCTURE: host stand-ins for the runtime ABI a generated object
// calls (include/tiramisu_b200.h), for the emulated builds of tests/emulation only.  Plain loops in the order the
// statements are written; the pattern entry points (GEMM, heat2d, conv block) forward to straightforward loop
// nests / the oracle.  Nothing here is linked into libtiramisu_b200.so or any shipped binary.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

struct halide_dimension_t { int32_t min, extent, stride; uint32_t flags; };
struct halide_type_t { uint8_t code, bits; uint16_t lanes; };
struct halide_buffer_t { uint64_t device; const void *device_interface; uint8_t *host; uint64_t flags; halide_type_t type;
                         int32_t dimensions; halide_dimension_t *dim; void *padding; };

// row-major C[MxN] = alpha * op(A) * op(B) + beta * C with element offsets and leading dimensions (0 = tight),
// sequential k, separate multiply and add: the semantics of src/tiramisu_cuda_wrappers.cpp:74-115
template <typename T>
static int gemm(T *A, T *B, T *C, uint64_t M, uint64_t N, uint64_t K, T alpha, T beta, uint64_t ldA, uint64_t ldB, uint64_t ldC,
                uint64_t offA, uint64_t offB, uint64_t offC, bool tA, bool tB)
{
    if (!ldA) ldA = tA ? M : K;
    if (!ldB) ldB = tB ? K : N;
    if (!ldC) ldC = N;
    A += offA; B += offB; C += offC;
    for (uint64_t i = 0; i < M; i++)
        for (uint64_t j = 0; j < N; j++) {
            T acc = 0;
            for (uint64_t k = 0; k < K; k++) {
                const T a = tA ? A[k * ldA + i] : A[i * ldA + k];
                const T b = tB ? B[j * ldB + k] : B[k * ldB + j];
                acc = acc + a * b;
            }
            C[i * ldC + j] = (beta != 0) ? alpha * acc + beta * C[i * ldC + j] : alpha * acc;
        }
    return 0;
}

extern "C" {
void oracle_heat2d(const float *in, float *out, int H, int W);
void oracle_heat2dgpu(const float *in, float *out, int H, int W);
void oracle_blurxy_u8(const uint8_t *in, uint8_t *out, int C, int H, int W);
void oracle_conv_relu_maxpool(const float *in, const float *w, const float *b, float *out, int B, int N, int FIn, int FOut);

void *tiramisu_cuda_malloc(uint64_t n) { return std::calloc(n ? n : 1, 1); }
int tiramisu_cuda_free(void *p) { std::free(p); return 0; }
int tiramisu_cuda_memcpy_to_device(void *d, void *s, uint64_t n) { std::memcpy(d, s, n); return 0; }
int tiramisu_cuda_memcpy_to_host(void *d, void *s, uint64_t n) { std::memcpy(d, s, n); return 0; }
int tiramisu_cuda_memcpy_to_device_async(void *d, const void *s, uint64_t n) { std::memcpy(d, s, n); return 0; }
int32_t tiramisu_cuda_stream_synchronize(int32_t) { return 0; }
void *tiramisu_b200_get_stream(void) { return nullptr; }
int tiramisu_b200_post_launch(const char *) { return 0; }

int tiramisu_cublas_sgemm(float *A, float *B, float *C, uint64_t M, uint64_t N, uint64_t K, float alpha, float beta, uint64_t ldA,
                          uint64_t ldB, uint64_t ldC, uint64_t offA, uint64_t offB, uint64_t offC, bool tA, bool tB)
{
    return gemm<float>(A, B, C, M, N, K, alpha, beta, ldA, ldB, ldC, offA, offB, offC, tA, tB);
}
int tiramisu_cublas_dgemm(double *A, double *B, double *C, uint64_t M, uint64_t N, uint64_t K, double alpha, double beta, uint64_t ldA,
                          uint64_t ldB, uint64_t ldC, uint64_t offA, uint64_t offB, uint64_t offC, bool tA, bool tB)
{
    return gemm<double>(A, B, C, M, N, K, alpha, beta, ldA, ldB, ldC, offA, offB, offC, tA, tB);
}
int tiramisu_b200_heat2d(const float *in, float *out, int32_t H, int32_t W, int32_t form, int32_t rb, int32_t re)
{
    if (rb != 0 || re != H || form > 1) { std::fprintf(stderr, "runtime_emu: heat2d variant not emulated\n"); return 1; }
    if (form == 0) oracle_heat2d(in, out, H, W); else oracle_heat2dgpu(in, out, H, W);
    return 0;
}
int tiramisu_b200_blurxy_u8(const uint8_t *in, uint8_t *out, int32_t C, int32_t H, int32_t W)
{
    oracle_blurxy_u8(in, out, C, H, W);
    return 0;
}
int tiramisu_b200_heat2d_buffers(halide_buffer_t *in, halide_buffer_t *out, int32_t H, int32_t W, int32_t form)
{
    return tiramisu_b200_heat2d(reinterpret_cast<const float *>(in->host), reinterpret_cast<float *>(out->host), H, W, form, 0, H);
}
int tiramisu_b200_conv_relu_maxpool(const float *in, const float *w, const float *b, float *out, int32_t B, int32_t N, int32_t FIn, int32_t FOut)
{
    oracle_conv_relu_maxpool(in, w, b, out, B, N, FIn, FOut);
    return 0;
}
}
