/* tiramisu_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * CPU restatement of the loop nests that the reference's x86 codegen
 * (ISL -> Halide -> LLVM; /root/reference/src/tiramisu_codegen_halide.cpp)
 * executes for the hot path of BASELINE.json.  The reference itself cannot be
 * built here (needs ISL+Halide 14+LLVM 13+clang), so every
 * function below re-states the *expression text and association order* of the
 * generator it cites, compiled without FMA contraction
 * (-ffp-contract=off, matching the AVX/SSE4.1-only target of
 * src/tiramisu_codegen_halide.cpp:3988-3995).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (tiramisu_b200/) never links or calls it.
 *
 * Pinning (what ties this file to the reference's own outputs):
 *   - oracle_blurxy_u8      : bit-exact against utils/images/reference_blurxy.png
 *                             on utils/images/rgb.png (tests/golden/blurxy_*.npy)
 *   - oracle_boxblur_u32    : tests/wrapper_test_160.cpp:39-44 known answer
 *   - oracle_heat2d*        : benchmarks/halide/wrapper_heat2d.cpp (all-ones => 1.9f / 0)
 *   - oracle_sgemm*         : sgemm_wrapper.cpp:24,164-169 (all-ones => 3K+3) and the
 *                             integer-valued tests 162/164/165/167
 *   - oracle_conv_relu_maxpool, oracle_lstm : NOT pinned against a reference-produced
 *                             vector (the reference checks them only against MKL /
 *                             MKL-DNN dumps that cannot be produced here).  Pinned instead
 *                             against independent fp64 PyTorch results on the reference
 *                             wrappers' own srand/rand data under the wrappers' own
 *                             acceptance protocol (tests/golden/{conv,lstm}_*.npy,
 *                             tools/make_golden_dnn.py, tests/test_oracle.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define IDX2(y, x, W) ((size_t)(y) * (size_t)(W) + (size_t)(x))

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Launchers such as torchrun export OMP_NUM_THREADS=1 to every rank; the CPU baseline legs of
 * bench.py set the thread count explicitly (all host cores, or 1 for the single-thread number). */
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------ heat2d
 * benchmarks/halide/heat2d_tiramisu.cpp:52-69
 *   heat2d_s0: out[y][x] = 0 for all y,x                       (:52-56)
 *   heat2d_s1: 1<=y<=H-2, 1<=x<=W-2:
 *     out = alpha*in[y][x] + beta*(((in[y][x+1] + in[y][x-1]) + in[y+1][x]) + in[y-1][x])   (:68)
 *   alpha = 0.3f, beta = 0.4f                                   (:26-27)
 * schedule: parallel y, vectorize x by 8 (:83-86) — does not change per-point arithmetic.
 */
void oracle_heat2d(const float *in, float *out, int H, int W)
{
    const float alpha = 0.3f, beta = 0.4f;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            out[IDX2(y, x, W)] = 0.0f;
#pragma omp parallel for schedule(static)
    for (int y = 1; y <= H - 2; y++)
        for (int x = 1; x <= W - 2; x++) {
            float s = in[IDX2(y, x + 1, W)] + in[IDX2(y, x - 1, W)];
            s = s + in[IDX2(y + 1, x, W)];
            s = s + in[IDX2(y - 1, x, W)];
            float a = alpha * in[IDX2(y, x, W)];
            float b = beta * s;
            out[IDX2(y, x, W)] = a + b;
        }
}

/* benchmarks/halide/heat2dgpu_tiramisu.cpp:54
 *   out[i+1][j+1] = c*alpha + (in[i+1] + in[j+1] + in[i-1] + in[j-1]) * beta
 *   i.e. (((S + E) + N) + W) * beta; borders written 0 (out_init_i/out_init_j :52-53).
 * Same association is used by heat2d_dist (heat2d_dist_tiramisu.cpp:88: (N+S+W+E)); see
 * oracle_heat2d_step for that one.
 */
void oracle_heat2dgpu(const float *in, float *out, int H, int W)
{
    const float alpha = 0.3f, beta = 0.4f;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            if (y == 0 || y == H - 1 || x == 0 || x == W - 1) {
                out[IDX2(y, x, W)] = 0.0f;
                continue;
            }
            float s = in[IDX2(y + 1, x, W)] + in[IDX2(y, x + 1, W)];
            s = s + in[IDX2(y - 1, x, W)];
            s = s + in[IDX2(y, x - 1, W)];
            float a = in[IDX2(y, x, W)] * alpha;
            float b = s * beta;
            out[IDX2(y, x, W)] = a + b;
        }
}

/* One Jacobi half-step of benchmarks/halide/heat2d_dist_tiramisu.cpp:88,92:
 *   out[i][j] = in[i][j]*alpha + (((in[i-1][j] + in[i+1][j]) + in[i][j-1]) + in[i][j+1]) * beta
 * on rows 1..H-2, cols 1..W-2; border cells of `out` are left untouched
 * (init_even/init_odd only cover the interior, :61-63).
 */
void oracle_heat2d_step(const float *in, float *out, int H, int W)
{
    const float alpha = 0.3f, beta = 0.4f;
#pragma omp parallel for schedule(static)
    for (int y = 1; y <= H - 2; y++)
        for (int x = 1; x <= W - 2; x++) {
            float s = in[IDX2(y - 1, x, W)] + in[IDX2(y + 1, x, W)];
            s = s + in[IDX2(y, x - 1, W)];
            s = s + in[IDX2(y, x + 1, W)];
            float a = in[IDX2(y, x, W)] * alpha;
            float b = s * beta;
            out[IDX2(y, x, W)] = a + b;
        }
}

/* ------------------------------------------------------------------ blur
 * benchmarks/halide/blurxy_tiramisu.cpp:61,77 (uint8; sums wrap mod 256 before /3)
 *   bx[c][y][x]  = ((in[c][y][x] + in[c][y][x+1]) + in[c][y][x+2]) / 3     y < H-6, x < W-8
 *   out[c][y][x] = ((bx[c][y][x] + bx[c][y+1][x]) + bx[c][y+2][x]) / 3     y < H-8, x < W-8
 * in: u8[C][H][W], out: u8[C][H-8][W-8].
 */
void oracle_blurxy_u8(const uint8_t *in, uint8_t *out, int C, int H, int W)
{
    const int OH = H - 8, OW = W - 8;
#pragma omp parallel for collapse(2) schedule(static)
    for (int c = 0; c < C; c++)
        for (int y = 0; y < OH; y++) {
            const uint8_t *p = in + ((size_t)c * H + y) * W;
            uint8_t *o = out + ((size_t)c * OH + y) * OW;
            for (int x = 0; x < OW; x++) {
                uint8_t bx[3];
                for (int r = 0; r < 3; r++) {
                    const uint8_t *q = p + (size_t)r * W + x;
                    uint8_t s = (uint8_t)(q[0] + q[1]);
                    s = (uint8_t)(s + q[2]);
                    bx[r] = (uint8_t)(s / (uint8_t)3);
                }
                uint8_t t = (uint8_t)(bx[0] + bx[1]);
                t = (uint8_t)(t + bx[2]);
                o[x] = (uint8_t)(t / (uint8_t)3);
            }
        }
}

/* tests/test_160.cpp:23-29 and benchmarks/automatic_comm/blurautodist_tiramisu.cpp
 * (uint32, VERTICAL pass first):
 *   bx[i][j]  = (in[i][j] + in[i+1][j] + in[i+2][j]) / 3      0<=i<rows, 0<=j<cols+2
 *   out[i][j] = (bx[i][j] + bx[i][j+1] + bx[i][j+2]) / 3      0<=i<rows, 0<=j<cols
 * in: u32[rows+2][cols+2] (the 2 extra rows are the halo the reference receives
 * from rank+1), out: u32[rows][cols].
 */
void oracle_boxblur_u32(const uint32_t *in, uint32_t *out, int rows, int cols)
{
    const int IW = cols + 2;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < cols; j++) {
            uint32_t bx[3];
            for (int d = 0; d < 3; d++) {
                uint32_t s = in[IDX2(i, j + d, IW)] + in[IDX2(i + 1, j + d, IW)];
                s = s + in[IDX2(i + 2, j + d, IW)];
                bx[d] = s / 3u;
            }
            uint32_t t = bx[0] + bx[1];
            t = t + bx[2];
            out[IDX2(i, j, cols)] = t / 3u;
        }
}

/* fp32 3x3 separable blur for BASELINE config 2 (SURVEY.md §8 a8'): the reference has
 * no fp32 blur, so this restates blurxy (a7: horizontal pass first, same association,
 * true division by 3.0f) with the "valid" output extent (H-2)x(W-2).
 *   bx[y][x]  = ((in[y][x] + in[y][x+1]) + in[y][x+2]) / 3.0f
 *   out[y][x] = ((bx[y][x] + bx[y+1][x]) + bx[y+2][x]) / 3.0f
 */
void oracle_blur_f32(const float *in, float *out, int H, int W)
{
    const int OH = H - 2, OW = W - 2;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < OH; y++)
        for (int x = 0; x < OW; x++) {
            float bx[3];
            for (int r = 0; r < 3; r++) {
                const float *q = in + IDX2(y + r, x, W);
                float s = q[0] + q[1];
                s = s + q[2];
                bx[r] = s / 3.0f;
            }
            float t = bx[0] + bx[1];
            t = t + bx[2];
            out[IDX2(y, x, OW)] = t / 3.0f;
        }
}

/* ------------------------------------------------------------------ sgemm
 * benchmarks/linear_algebra/blas/level3/sgemm/cpu/sgemm_generator.cpp:90-110
 *   reduced_AB_0 : acc = 0
 *   reduced_AB_1 : acc = acc + A[i][k]*B[k][j]     (k increasing for a fixed (i,j): the
 *                  64x32x32 / 2x4x4 tilings keep k monotone per output element)
 *   result       : C[i][j] = a*acc + b*C[i][j]
 * A f32[N][K], B f32[K][M], C f32[N][M]  (SIZES = {N, M, K}, :296-304)
 * The i-k-j loop order below is the reference's own innermost order (j vectorised);
 * per (i,j) the k sequence is 0..K-1, products rounded before the add (no FMA).
 */
void oracle_sgemm(int N, int M, int K, float a, float b,
                  const float *A, const float *B, float *C)
{
#pragma omp parallel
    {
        float *acc = (float *)malloc(sizeof(float) * (size_t)M);
#pragma omp for schedule(static)
        for (int i = 0; i < N; i++) {
            for (int j = 0; j < M; j++) acc[j] = 0.0f;
            for (int k = 0; k < K; k++) {
                const float aik = A[IDX2(i, k, K)];
                const float *brow = B + (size_t)k * M;
                for (int j = 0; j < M; j++) {
                    float p = aik * brow[j];
                    acc[j] = acc[j] + p;
                }
            }
            for (int j = 0; j < M; j++) {
                float t0 = a * acc[j];
                float t1 = b * C[IDX2(i, j, M)];
                C[IDX2(i, j, M)] = t0 + t1;
            }
        }
        free(acc);
    }
}

/* The same statement triple under (a restatement of) the reference's CPU schedule — parallel outer i
 * tile, register-blocked rows, 8-wide vector lanes over j (sgemm_generator.cpp:73-79 L2 tiles 64x32x32,
 * :132-138 parallel, :273-289 unroll/vectorize) — used as the *timed* CPU baseline of bench.py.  Per
 * output element k still runs 0..K-1 with the product rounded before the add, so the result is
 * bit-identical to oracle_sgemm (checked in tests/test_oracle.py).
 */
#define SG_TI 4
#define SG_TJ 64
void oracle_sgemm_sched(int N, int M, int K, float a, float b,
                        const float *restrict A, const float *restrict B, float *restrict C)
{
    const int nti = (N + SG_TI - 1) / SG_TI, ntj = (M + SG_TJ - 1) / SG_TJ;
#pragma omp parallel for collapse(2) schedule(static)
    for (int ti = 0; ti < nti; ti++)
        for (int tj = 0; tj < ntj; tj++) {
            const int i0 = ti * SG_TI, j0 = tj * SG_TJ;
            const int ni = N - i0 < SG_TI ? N - i0 : SG_TI, nj = M - j0 < SG_TJ ? M - j0 : SG_TJ;
            float acc[SG_TI][SG_TJ];
            for (int ii = 0; ii < SG_TI; ii++)
                for (int jj = 0; jj < SG_TJ; jj++) acc[ii][jj] = 0.0f;
            if (ni == SG_TI && nj == SG_TJ) {
                for (int k = 0; k < K; k++) {
                    const float *restrict brow = B + (size_t)k * M + j0;
                    for (int ii = 0; ii < SG_TI; ii++) {
                        const float aik = A[IDX2(i0 + ii, k, K)];
#pragma omp simd
                        for (int jj = 0; jj < SG_TJ; jj++) {
                            float p = aik * brow[jj];
                            acc[ii][jj] = acc[ii][jj] + p;
                        }
                    }
                }
            } else {
                for (int k = 0; k < K; k++) {
                    const float *brow = B + (size_t)k * M + j0;
                    for (int ii = 0; ii < ni; ii++) {
                        const float aik = A[IDX2(i0 + ii, k, K)];
                        for (int jj = 0; jj < nj; jj++) {
                            float p = aik * brow[jj];
                            acc[ii][jj] = acc[ii][jj] + p;
                        }
                    }
                }
            }
            for (int ii = 0; ii < ni; ii++)
                for (int jj = 0; jj < nj; jj++) {
                    float t0 = a * acc[ii][jj];
                    float t1 = b * C[IDX2(i0 + ii, j0 + jj, M)];
                    C[IDX2(i0 + ii, j0 + jj, M)] = t0 + t1;
                }
        }
}

/* Row-major C[MxN] = alpha*op(A)*op(B) + beta*C with leading dimensions, element
 * offsets and transposes: the contract of tiramisu_cublas_sgemm
 * (src/tiramisu_cuda_wrappers.cpp:74-115; ld==0 => tight, :90-98), also the math of
 * benchmarks/.../sgemm/gpu/generator1.cpp:68-72 (c_acc(i,j,0)*alpha + c_C(i,j)*beta).
 * k sequential.
 */
void oracle_gemm_ld(const float *A, const float *B, float *C,
                    uint64_t M, uint64_t N, uint64_t K, float alpha, float beta,
                    uint64_t ldA, uint64_t ldB, uint64_t ldC,
                    uint64_t offA, uint64_t offB, uint64_t offC, int tA, int tB)
{
    if (ldA == 0) ldA = tA ? M : K;
    if (ldB == 0) ldB = tB ? K : N;
    if (ldC == 0) ldC = N;
    A += offA; B += offB; C += offC;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)M; i++)
        for (uint64_t j = 0; j < N; j++) {
            float acc = 0.0f;
            for (uint64_t k = 0; k < K; k++) {
                float av = tA ? A[k * ldA + i] : A[i * ldA + k];
                float bv = tB ? B[j * ldB + k] : B[k * ldB + j];
                float p = av * bv;
                acc = acc + p;
            }
            float t0 = acc * alpha;
            float t1 = C[i * ldC + j] * beta;
            C[i * ldC + j] = t0 + t1;
        }
}

void oracle_dgemm_ld(const double *A, const double *B, double *C,
                     uint64_t M, uint64_t N, uint64_t K, double alpha, double beta,
                     uint64_t ldA, uint64_t ldB, uint64_t ldC,
                     uint64_t offA, uint64_t offB, uint64_t offC, int tA, int tB)
{
    if (ldA == 0) ldA = tA ? M : K;
    if (ldB == 0) ldB = tB ? K : N;
    if (ldC == 0) ldC = N;
    A += offA; B += offB; C += offC;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)M; i++)
        for (uint64_t j = 0; j < N; j++) {
            double acc = 0.0;
            for (uint64_t k = 0; k < K; k++) {
                double av = tA ? A[k * ldA + i] : A[i * ldA + k];
                double bv = tB ? B[j * ldB + k] : B[k * ldB + j];
                double p = av * bv;
                acc = acc + p;
            }
            double t0 = acc * alpha;
            double t1 = C[i * ldC + j] * beta;
            C[i * ldC + j] = t0 + t1;
        }
}

/* tutorials/developers/tutorial_04gpu.cpp:54 — uint8 matmul, arithmetic wraps mod 256:
 *   C[i][j] = sum_k A[i][k]*B[k][j]   (all uint8)
 */
void oracle_matmul_u8(const uint8_t *A, const uint8_t *B, uint8_t *C, int N)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
            uint8_t acc = 0;
            for (int k = 0; k < N; k++)
                acc = (uint8_t)(acc + (uint8_t)(A[IDX2(i, k, N)] * B[IDX2(k, j, N)]));
            C[IDX2(i, j, N)] = acc;
        }
}

/* ------------------------------------------------------------------ conv+ReLU+maxpool
 * benchmarks/DNN/blocks/Conv-ReLU-MaxPool/cpu/dense/conv_relu_maxpool_generator_tiramisu.cpp
 *   init_output : out = 0                                                   (:28)
 *   conv_init   : acc = bias[fout_b][ffout]                                 (:31)
 *   conv        : acc = acc + in[n][y+ky][x+kx][fin] * w[fout_b][ky][kx][fin][ffout]
 *                 loop order ky, kx, fin (ffout, xx vector/register lanes)  (:33-37, :68-73)
 *   maxpool     : out[n][fout_b][y/2][x/2][ffout] = max(out[...], acc)      (:41-49, :104-105)
 * layouts (:21-23, :90): input f32[B][N+2][N+2][FIn] (pre-padded NHWC),
 * filter f32[FOut/8][3][3][FIn][8], bias f32[FOut], output f32[B][FOut/8][N/2][N/2][8].
 */
void oracle_conv_relu_maxpool(const float *input, const float *filter, const float *bias,
                              float *output, int B, int N, int FIn, int FOut)
{
    const int FB = FOut / 8, NP = N + 2, NO = N / 2;
    memset(output, 0, sizeof(float) * (size_t)B * FB * NO * NO * 8);
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < B; n++)
        for (int fb = 0; fb < FB; fb++)
            for (int y = 0; y < N; y++)
                for (int x = 0; x < N; x++) {
                    float acc[8];
                    for (int f = 0; f < 8; f++) acc[f] = bias[fb * 8 + f];
                    for (int ky = 0; ky < 3; ky++)
                        for (int kx = 0; kx < 3; kx++) {
                            const float *ip = input + (((size_t)n * NP + (y + ky)) * NP + (x + kx)) * FIn;
                            const float *wp = filter + ((((size_t)fb * 3 + ky) * 3 + kx) * FIn) * 8;
                            for (int fin = 0; fin < FIn; fin++) {
                                const float iv = ip[fin];
                                for (int f = 0; f < 8; f++) {
                                    float p = iv * wp[fin * 8 + f];
                                    acc[f] = acc[f] + p;
                                }
                            }
                        }
                    float *op = output + ((((size_t)n * FB + fb) * NO + y / 2) * NO + x / 2) * 8;
                    for (int f = 0; f < 8; f++)
                        op[f] = op[f] > acc[f] ? op[f] : acc[f];
                }
}

/* ------------------------------------------------------------------ LSTM
 * benchmarks/DNN/blocks/LSTM/cpu/generator.cpp:35-53, :97-117
 *   h[-1]=c[-1]=0; h[m][-1]=x[m]                                             (:35-37)
 *   s = b[l][g]; for j: s = (s + R[l][g][j]*h[m-1][l][k][j]) + W[l][g][j]*h[m][l-1][k][j]   (:40-41)
 *   rows g in [0,F)=i (sigmoid), [F,2F)=z (tanh), [2F,3F)=o (sigmoid), [3F,4F)=f (sigmoid)  (:43-46)
 *   sigmoid(x) = 1/(1+exp(-x))                                               (:42)
 *   c = i*z + f*c_prev ; h = tanh(c)*o                                       (:47-51)
 *   y[m] = h[m][L-1]                                                         (:53)
 * Weights f32[L][2][4F][F] (index 0 = R recurrent, 1 = W input), biases f32[L][4F],
 * x,y f32[T][B][F].  expf/tanhf = libm (Halide calls libm for f32 exp/tanh).
 */
static inline float sigmoidf_ref(float v) { return 1.0f / (1.0f + expf(-v)); }

void oracle_lstm_c(int F, int B, int L, int T, const float *Weights, const float *biases,
                   const float *x, float *y, float *c_last);

void oracle_lstm(int F, int B, int L, int T, const float *Weights, const float *biases,
                 const float *x, float *y)
{
    oracle_lstm_c(F, B, L, T, Weights, biases, x, y, NULL);
}

/* c_last (optional, f32[B][F]): the last layer's cell state after the last step — the `c` working buffer that
 * LSTM/cpu_lib/wrapper.cpp:36,70 passes and the generator leaves filled (cpu_lib/generator.cpp:41,78). */
void oracle_lstm_c(int F, int B, int L, int T, const float *Weights, const float *biases,
                   const float *x, float *y, float *c_last)
{
    const size_t BF = (size_t)B * F;
    /* buf_h[T+1][L+1][B][F], buf_c[T+1][L][B][F] as in the generator (:91-92) */
    float *h = (float *)calloc((size_t)(T + 1) * (L + 1) * BF, sizeof(float));
    float *c = (float *)calloc((size_t)(T + 1) * L * BF, sizeof(float));
    float *tmp = (float *)malloc(sizeof(float) * (size_t)B * 4 * F);
#define H_(m, l) (h + (((size_t)(m) + 1) * (L + 1) + ((size_t)(l) + 1)) * BF)
#define C_(m, l) (c + (((size_t)(m) + 1) * L + (size_t)(l)) * BF)
    for (int m = 0; m < T; m++)
        memcpy(H_(m, -1), x + (size_t)m * BF, sizeof(float) * BF);
    for (int m = 0; m < T; m++)
        for (int l = 0; l < L; l++) {
            const float *R = Weights + ((size_t)l * 2 + 0) * 4 * F * F;
            const float *W = Weights + ((size_t)l * 2 + 1) * 4 * F * F;
            const float *bl = biases + (size_t)l * 4 * F;
            const float *hprev = H_(m - 1, l), *hbelow = H_(m, l - 1);
#pragma omp parallel for collapse(2) schedule(static)
            for (int k = 0; k < B; k++)
                for (int g = 0; g < 4 * F; g++) {
                    float s = bl[g];
                    const float *r = R + (size_t)g * F, *w = W + (size_t)g * F;
                    const float *hp = hprev + (size_t)k * F, *hb = hbelow + (size_t)k * F;
                    for (int j = 0; j < F; j++) {
                        float p0 = r[j] * hp[j];
                        s = s + p0;
                        float p1 = w[j] * hb[j];
                        s = s + p1;
                    }
                    tmp[(size_t)k * 4 * F + g] = s;
                }
            const float *cprev = C_(m - 1, l);
            float *cn = C_(m, l), *hn = H_(m, l);
#pragma omp parallel for collapse(2) schedule(static)
            for (int k = 0; k < B; k++)
                for (int i = 0; i < F; i++) {
                    const float *t = tmp + (size_t)k * 4 * F;
                    float gi = sigmoidf_ref(t[i]);
                    float gz = tanhf(t[i + F]);
                    float go = sigmoidf_ref(t[i + 2 * F]);
                    float gf = sigmoidf_ref(t[i + 3 * F]);
                    float iz = gi * gz;
                    float fc = gf * cprev[(size_t)k * F + i];
                    float cv = iz + fc;
                    cn[(size_t)k * F + i] = cv;
                    hn[(size_t)k * F + i] = tanhf(cv) * go;
                }
        }
    for (int m = 0; m < T; m++)
        memcpy(y + (size_t)m * BF, H_(m, L - 1), sizeof(float) * BF);
    if (c_last) memcpy(c_last, C_(T - 1, L - 1), sizeof(float) * BF);
#undef H_
#undef C_
    free(h); free(c); free(tmp);
}

/* ------------------------------------------------------------------ small known-answer helpers
 * tests/test_197.cpp: every element of an int32 buffer := constant, per time-step.
 * tests/test_168.cpp: B[j][i] = A[i][j] (128x256 fp32 transpose).
 */
void oracle_fill_i32(int32_t *buf, size_t n, int32_t v)
{
    for (size_t i = 0; i < n; i++) buf[i] = v;
}

void oracle_transpose_f32(const float *A, float *Bt, int rows, int cols)
{
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < cols; j++)
            Bt[IDX2(j, i, rows)] = A[IDX2(i, j, cols)];
}

/* ------------------------------------------------------------------ libc rand() replay
 * The reference wrappers seed and draw from glibc's rand(); tests replay the exact
 * sequences (e.g. wrapper_nn_block.cpp:15-39, LSTM/cpu/wrapper.cpp:31-52).
 */
void oracle_srand(unsigned seed) { srand(seed); }
void oracle_rand_fill(int32_t *out, size_t n)
{
    for (size_t i = 0; i < n; i++) out[i] = rand();
}
