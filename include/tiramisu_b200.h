/* tiramisu_b200.h — C ABI of the B200-native (sm_100a) backend for Tiramisu's hot path.
 *
 * Two nested boundaries, both plain C (pointers and sizes only), exported by
 * tiramisu_b200/lib/libtiramisu_b200.so:
 *
 *  B1  generated-function ABI: `int <fn>(halide_buffer_t *...)`, the symbols the
 *      reference's wrapper_*.cpp / *_wrapper.cpp harnesses link against.
 *      Return 0 on success, non-zero on a validation error (message via
 *      tiramisu_b200_last_error()).  Buffers may live on the host (->host, copies are
 *      done inside, as the reference's generated code does) or on the device
 *      (->host == NULL and ->device = CUDA pointer: no copies).
 *
 *  B2  runtime ABI: the tiramisu_cuda_* / tiramisu_cublas_* / tiramisu_address_of_* /
 *      tiramisu_MPI_* shims that Tiramisu-generated host code calls
 *      (reference: src/tiramisu_cuda_wrappers.cpp, src/tiramisu_externs.cpp,
 *      src/tiramisu_mpi.cpp), plus per-kernel launchers on device pointers — the role of
 *      the reference's `_kernel_<N>_wrapper` functions (src/tiramisu_codegen_cuda.cpp:855-866),
 *      here named after what they compute.
 *
 * All launches go to the calling thread's stream (tiramisu_b200_set_stream; default
 * cudaStreamPerThread — the reference compiles with --default-stream per-thread,
 * src/tiramisu_codegen_cuda.cpp:1745).
 */
#ifndef TIRAMISU_B200_H
#define TIRAMISU_B200_H

#include <stdbool.h>
#include <stdint.h>

#include "tiramisu_b200/halide_buffer.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ===================================================================== library control */
const char *tiramisu_b200_version(void);
/* Last error message recorded on any thread ("" if none). */
const char *tiramisu_b200_last_error(void);
/* 0 (default): CUDA failures print and exit(1) like src/tiramisu_cuda_wrappers.cpp:14-31.
 * 1: they return non-zero / NULL instead. Returns the previous mode. */
int tiramisu_b200_set_error_mode(int return_instead_of_exit);
/* Launch stream of the calling thread (a cudaStream_t); NULL restores cudaStreamPerThread. */
void tiramisu_b200_set_stream(void *cuda_stream);
void *tiramisu_b200_get_stream(void);
/* Number of kernels this library has launched since load (all threads). */
uint64_t tiramisu_b200_kernel_launches(void);
/* Called by code the sm_100a lowering generates (tiramisu_b200/frontend/src/lowering_emit.cpp) right after
 * each `<<<>>>` launch: counts the launch and reports a launch error like the reference's kernel
 * wrappers do (src/tiramisu_codegen_cuda.cpp:855-866).  Returns 0 on success. */
int tiramisu_b200_post_launch(const char *kernel_name);
/* Measurement aid: launches an EMPTY kernel of the given shape (grid, block, dynamic shared memory, cluster width)
 * on the thread's stream — the launch floor that a few-microsecond kernel sits on (bench.py reports it next to
 * the 1024^3 GEMM). */
int tiramisu_b200_null_launch(int32_t grid, int32_t block, int32_t smem_bytes, int32_t cluster);
/* Release cached pool memory back to the driver. */
int tiramisu_b200_pool_trim(void);

/* ===================================================================== B2: runtime shims
 * Same names and argument meaning as src/tiramisu_cuda_wrappers.cpp. */
int tiramisu_cuda_memcpy_to_device(void *dst_dev, void *src_host, uint64_t bytes);   /* :34 blocking */
int tiramisu_cuda_memcpy_to_host(void *dst_host, void *src_dev, uint64_t bytes);     /* :41 blocking */
int tiramisu_cuda_memcpy_to_symbol(void *symbol, void *src_host, uint64_t bytes);    /* :48 */
void *tiramisu_cuda_malloc(uint64_t bytes);                                          /* :55 pooled, stream-ordered */
int tiramisu_cuda_free(void *ptr);                                                   /* :63 */
int32_t tiramisu_cuda_stream_synchronize(int32_t dummy);                             /* :164 syncs the thread's stream */
/* Stream-ordered variants used inside the generated functions (no host sync). */
int tiramisu_cuda_memcpy_to_device_async(void *dst_dev, const void *src_host, uint64_t bytes);
int tiramisu_cuda_memcpy_to_host_async(void *dst_host, const void *src_dev, uint64_t bytes);

/* Row-major C[MxN] = alpha*op(A)*op(B) + beta*C on DEVICE pointers, element offsets, leading
 * dimensions (0 => tight), transposes.  src/tiramisu_cuda_wrappers.cpp:74-115 / :121-162.
 * fp32 runs on tcgen05 TF32 tiles (fp32 accumulate) when the operands are TMA-addressable
 * (16-byte aligned bases and leading dimensions), else on an exact fp32 SIMT kernel. */
int tiramisu_cublas_sgemm(float *A, float *B, float *C, uint64_t M, uint64_t N, uint64_t K,
                          float alpha, float beta, uint64_t ldA, uint64_t ldB, uint64_t ldC,
                          uint64_t offsetA, uint64_t offsetB, uint64_t offsetC,
                          bool transposeA, bool transposeB);
int tiramisu_cublas_dgemm(double *A, double *B, double *C, uint64_t M, uint64_t N, uint64_t K,
                          double alpha, double beta, uint64_t ldA, uint64_t ldB, uint64_t ldC,
                          uint64_t offsetA, uint64_t offsetB, uint64_t offsetC,
                          bool transposeA, bool transposeB);
/* 0 = auto (tensor cores when addressable), 1 = force exact fp32 SIMT, 2 = force tcgen05
 * (fails if not addressable).  Returns the previous mode. */
int tiramisu_b200_set_gemm_mode(int mode);
/* Which tcgen05 GEMM kernel the tensor path uses: 0 = auto (by problem size), 1 = one output tile
 * per CTA, 2 = persistent single-CTA tiles (128x256), 3 = persistent CTA pairs (cta_group::2,
 * 256x256), 4 = split-K: a 2-CTA cluster per 128x128 tile, the two K halves combined through distributed
 * shared memory (small outputs with a long K, e.g. 1024^3).  Tuning/diagnostic knob; returns the previous value. */
int tiramisu_b200_set_gemm_kernel(int kernel);

/* src/tiramisu_externs.cpp:8-54 */
int8_t *tiramisu_address_of_int8(halide_buffer_t *buffer, unsigned long index);
uint8_t *tiramisu_address_of_uint8(halide_buffer_t *buffer, unsigned long index);
int16_t *tiramisu_address_of_int16(halide_buffer_t *buffer, unsigned long index);
uint16_t *tiramisu_address_of_uint16(halide_buffer_t *buffer, unsigned long index);
int32_t *tiramisu_address_of_int32(halide_buffer_t *buffer, unsigned long index);
uint32_t *tiramisu_address_of_uint32(halide_buffer_t *buffer, unsigned long index);
int64_t *tiramisu_address_of_int64(halide_buffer_t *buffer, unsigned long index);
uint64_t *tiramisu_address_of_uint64(halide_buffer_t *buffer, unsigned long index);
float *tiramisu_address_of_float32(halide_buffer_t *buffer, unsigned long index);
double *tiramisu_address_of_float64(halide_buffer_t *buffer, unsigned long index);

/* ===================================================================== B2: kernel launchers
 * Device pointers, launched on the thread's stream, no synchronisation.  These replace the
 * `_kernel_<N>_wrapper`s that src/tiramisu_codegen_cuda.cpp emits for each GPU loop nest. */

/* 5-point stencil, alpha=0.3f beta=0.4f, borders of `out` written 0.
 * form 0: heat2d_tiramisu.cpp:68      a*c + b*(((E+W)+S)+N)
 * form 1: heat2dgpu_tiramisu.cpp:54   c*a + (((S+E)+N)+W)*b
 * form 2: heat2d_dist_tiramisu.cpp:88 c*a + (((N+S)+W)+E)*b, borders of `out` NOT written
 * Rows [row_begin,row_end) of an HxW grid are produced (a rank's shard); in/out point at row 0. */
int tiramisu_b200_heat2d(const float *in, float *out, int32_t H, int32_t W, int32_t form,
                         int32_t row_begin, int32_t row_end);
/* 3x3 separable blurs.  f32: in[H][W] -> out[H-2][W-2] (horizontal first, /3.0f).
 * u8 : blurxy_tiramisu.cpp:61,77   in[C][H][W] -> out[C][H-8][W-8], wraps mod 256.
 * u32: tests/test_160.cpp:23-29    in[rows+2][cols+2] -> out[rows][cols], vertical first. */
/* heat2d on halide buffers (host: row-chunk pipelined copies; device-resident: in place).  Called by
 * generated code when the lowering recognises the statement pair of heat2d_tiramisu.cpp:52-69. */
int tiramisu_b200_heat2d_buffers(halide_buffer_t *in, halide_buffer_t *out, int32_t H, int32_t W, int32_t form);
int tiramisu_b200_blur_f32(const float *in, float *out, int32_t H, int32_t W,
                           int32_t row_begin, int32_t row_end);

/* Row-sharded grids: halo rows are read IN PLACE from the neighbours' HBM through
 * peer mappings (tiramisu_b200_comm_alloc_symmetric) instead of being exchanged before the sweep —
 * replaces the four MPI transfers per sweep of benchmarks/halide/heat2d_dist_tiramisu.cpp:133-153
 * and the two of benchmarks/automatic_comm/blurautodist_ref.cpp:42-61.  The rank owns global rows
 * [lo, hi); `up` points at row lo-1 as the upper neighbour stores it, `dn` at row hi (and hi+1) as
 * the lower neighbour stores them; either may be NULL at the ends of the grid.
 * Iterated sweeps: with `flags` set (4 ints of symmetric memory, zero-initialised), the first / last
 * row block of the launch waits until flags[0] / flags[1] >= wait_value and the launch publishes
 * signal_value to *sig_up / *sig_dn (the neighbours' flags[1] / flags[0]) when those blocks retire. */
typedef struct tiramisu_b200_halo {
    const float *up;
    const float *dn;
    int32_t lo, hi;
    int32_t *flags;
    int32_t *sig_up;
    int32_t *sig_dn;
    int32_t wait_value;
    int32_t signal_value;
} tiramisu_b200_halo_t;
int tiramisu_b200_heat2d_peer(const float *in, float *out, int32_t H, int32_t W, int32_t form,
                              int32_t row_begin, int32_t row_end, const tiramisu_b200_halo_t *halo);
int tiramisu_b200_blur_f32_peer(const float *in, float *out, int32_t H, int32_t W,
                                int32_t row_begin, int32_t row_end, const tiramisu_b200_halo_t *halo);
int tiramisu_b200_blurxy_u8(const uint8_t *in, uint8_t *out, int32_t C, int32_t H, int32_t W);
int tiramisu_b200_boxblur_u32(const uint32_t *in, uint32_t *out, int32_t rows, int32_t cols);
/* tutorials/developers/tutorial_04gpu.cpp: uint8 NxN matmul (wraps mod 256). */
int tiramisu_b200_matmul_u8(const uint8_t *A, const uint8_t *B, uint8_t *C, int32_t N);
/* tests/test_197.cpp (int32 fill) and tests/test_168.cpp (fp32 transpose) kernels. */
int tiramisu_b200_fill_i32(int32_t *buf, uint64_t n, int32_t value);
int tiramisu_b200_transpose_f32(const float *A, float *At, int32_t rows, int32_t cols);

/* conv(3x3)+ReLU+maxpool(2x2) block (conv_relu_maxpool_generator_tiramisu.cpp:28-49):
 * input f32[B][N+2][N+2][FIn] pre-padded NHWC, filter f32[FOut/8][3][3][FIn][8], bias f32[FOut],
 * output f32[B][FOut/8][N/2][N/2][8].  tcgen05 implicit GEMM (TF32) when FIn % 32 == 0, FOut == 64,
 * N % 32 == 0; else an fp32 SIMT kernel in the reference's accumulation order (bit-exact). */
int tiramisu_b200_conv_relu_maxpool(const float *input, const float *filter, const float *bias,
                                    float *output, int32_t B, int32_t N, int32_t FIn, int32_t FOut);
/* 0 auto, 1 exact fp32 SIMT, 2 force tcgen05.  Returns the previous mode. */
int tiramisu_b200_set_conv_mode(int mode);
/* Activation staging of the tcgen05 conv: 0 = auto, 1 = 8x16 tile, vertical taps share a TMA tile,
 * 2 = 16x8 tile, all nine taps share one TMA tile, 3 = as 2 on CTA pairs (cta_group::2, M = 256, the
 * filter split between the two CTAs).  Returns the previous value. */
int tiramisu_b200_set_conv_kernel(int kernel);
/* Multi-layer LSTM forward (LSTM/cpu/generator.cpp:35-53): Weights f32[L][2][4F][F] (0 = R, 1 = W),
 * biases f32[L][4F], x,y f32[T][B][F]; gate rows i,z,o,f. */
int tiramisu_b200_lstm(const float *Weights, const float *biases, const float *x, float *y,
                       int32_t F, int32_t B, int32_t L, int32_t T);
int tiramisu_b200_set_lstm_mode(int mode);
/* Recurrence engine of the tensor path: 0 = auto, 1 = one launch per cell, 2 = persistent layer
 * kernel, gate columns split across CTAs (recurrent weights resident in shared memory, device-wide
 * step barrier), 3 = persistent layer kernel, K split across 4-CTA clusters with a distributed-
 * shared-memory reduction (F % 128 == 0, B <= 64). */
int tiramisu_b200_set_lstm_kernel(int kernel);
/* The same block for the parameter layouts of the reference's other two LSTM benchmarks: layout 0 = LSTM/cpu and
 * LSTM/gpu (above); 1 = LSTM/gpu_lib (Weights [L][2][4F][F], index 0 = W, 1 = R, gate rows i, f, z, o:
 * gpu_lib/generator.cpp:66-88); 2 = LSTM/cpu_lib (weights [L][2][F][4F], transposed, same index and gate order:
 * cpu_lib/generator.cpp:27,50-76).  c_last: null, or f32[B][F] that receives the last layer's final cell state. */
int tiramisu_b200_lstm_ex(const float *Weights, const float *biases, const float *x, float *y, float *c_last,
                          int32_t F, int32_t B, int32_t L, int32_t T, int32_t layout);
/* Description of the engine (and operand type) the calling thread's last tiramisu_b200_lstm call ran on. */
const char *tiramisu_b200_lstm_last_engine(void);

/* ===================================================================== B2: distributed runtime
 * NCCL over NVLink, one process per GPU, replacing src/tiramisu_mpi.cpp (MPI point-to-point).
 * Same names and argument meaning as include/tiramisu/mpi_comm.h; data pointers may be device
 * memory (sent in place) or host memory (staged).  Messages are stream-ordered. */
int tiramisu_b200_comm_unique_id(void *out128);                 /* 128-byte NCCL id (rank 0) */
int tiramisu_b200_comm_init(int32_t rank, int32_t world, const void *id128);
int tiramisu_b200_comm_destroy(void);
int tiramisu_b200_comm_rank(void);
int tiramisu_b200_comm_size(void);
int tiramisu_b200_comm_broadcast(void *dev_buf, uint64_t bytes, int32_t root);
/* Same result for large buffers, as scatter from the root + in-place all-gather (bandwidth-optimal). */
int tiramisu_b200_comm_broadcast_large(void *dev_buf, uint64_t bytes, int32_t root);
/* n send/recv pairs as ONE NCCL group; a peer < 0 skips that entry. */
int tiramisu_b200_comm_exchange(int32_t n, const int32_t *send_peer, void *const *send_buf,
                                const uint64_t *send_bytes, const int32_t *recv_peer,
                                void *const *recv_buf, const uint64_t *recv_bytes);
/* Symmetric device memory: every rank allocates `bytes` (cudaMalloc), the CUDA IPC handles are
 * all-gathered over the NCCL communicator and every peer's allocation is mapped into this process.
 * peers[r] (world entries, caller-provided) is rank r's block as addressable from this GPU over
 * NVLink; peers[rank] == *local.  Collective: every rank must call it, in the same order. */
int tiramisu_b200_comm_alloc_symmetric(uint64_t bytes, void **local, void **peers);
int tiramisu_b200_comm_free_symmetric(void *local);
/* Broadcast of [data_offset, data_offset + bytes) of the root's symmetric block into every rank's block, in `panels`
 * (1..64) equal parts, as peer copies by the DMA engines on two library-owned streams forked from / joined to the
 * thread's stream (scatter: every non-root rank pulls its slice of a panel from the root; all-gather: the other
 * slices from the ranks that own them) — no SM kernels, so it overlaps compute.  peers: the table
 * comm_alloc_symmetric returned; flags_offset: 256 zero-initialised bytes inside the block (world <= 30); epoch: 1, 2,
 * 3, ... per call on this buffer, the same on every rank.  Replaces the per-rank MPI_Recv of B in the reference's
 * distributed GEMMs (src/tiramisu_core.cpp:8997-9044 send/recv semantics) for the row-sharded sgemm. */
int tiramisu_b200_comm_broadcast_sym(void *const *peers, uint64_t data_offset, uint64_t bytes, uint64_t flags_offset,
                                     int32_t root, uint32_t epoch, int32_t panels);
/* Row-sharded sgemm with B's arrival fused in: C[rows][N] = alpha A[rows][K] B[K][N] + beta C on
 * every rank; A, C: this rank's row shard; B: an ordinary row-major device matrix (pitch ldB) on `root` only.  The
 * symmetric block at stage_offset (2 halves of K * N * 4 bytes, each rounded up to 256, used alternately) is the staging
 * area B travels through, in `panels` column panels
 * (N / panels a multiple of 256, same value on every rank) stored panel-major: the root re-lays B out with local DMA
 * copies, the panels travel as above, and every other rank runs ONE persistent tcgen05 GEMM launch over the staged
 * layout (3-D tensor map) whose TMA producer warp polls a panel's arrival word before the B loads of a tile in a
 * panel it has not seen yet.  Stream-ordered; collective over the ranks; without a communicator a plain sgemm. */
int tiramisu_b200_sgemm_bcast_b(void *const *peers, uint64_t stage_offset, uint64_t flags_offset, int32_t root, uint32_t epoch,
                                int32_t panels, const float *B, uint64_t ldB, const float *A, float *C, uint64_t rows,
                                uint64_t N, uint64_t K, float alpha, float beta, uint64_t ldA, uint64_t ldC);
/* Diagnosis: with TB200_BCAST_TRACE=1 in the environment, prints the timeline of this thread's last broadcast. */
int tiramisu_b200_comm_broadcast_trace_dump(void);
int tiramisu_MPI_init(void);                 /* src/tiramisu_mpi.cpp:8  reads torchrun's env; returns rank */
void tiramisu_MPI_cleanup(void);             /* :19 */
void tiramisu_MPI_global_barrier(void);      /* :23 */
int tiramisu_MPI_Comm_rank(int offset);      /* :78 */
void tiramisu_MPI_Wait(void *request);       /* :85 */
#define TIRAMISU_B200_MPI_DECL(SUFFIX, T)                                                 \
    void tiramisu_MPI_Send_##SUFFIX(int count, int dest, int tag, T *data);               \
    void tiramisu_MPI_Ssend_##SUFFIX(int count, int dest, int tag, T *data);              \
    void tiramisu_MPI_Isend_##SUFFIX(int count, int dest, int tag, T *data, long *reqs);  \
    void tiramisu_MPI_Issend_##SUFFIX(int count, int dest, int tag, T *data, long *reqs); \
    void tiramisu_MPI_Recv_##SUFFIX(int count, int source, int tag, T *store_in);         \
    void tiramisu_MPI_Irecv_##SUFFIX(int count, int source, int tag, T *store_in, long *reqs);
TIRAMISU_B200_MPI_DECL(int8, char)           /* src/tiramisu_mpi.cpp:96-192 */
TIRAMISU_B200_MPI_DECL(int16, short)
TIRAMISU_B200_MPI_DECL(int32, int)
TIRAMISU_B200_MPI_DECL(int64, long)
TIRAMISU_B200_MPI_DECL(uint8, unsigned char)
TIRAMISU_B200_MPI_DECL(uint16, unsigned short)
TIRAMISU_B200_MPI_DECL(uint32, unsigned int)
TIRAMISU_B200_MPI_DECL(uint64, unsigned long)
TIRAMISU_B200_MPI_DECL(f32, float)
TIRAMISU_B200_MPI_DECL(f64, double)

/* ===================================================================== B1: generated functions */
/* benchmarks/halide/wrapper_heat2d.h:10 — in,out f32[H][W] (reference bakes 10000x10000). */
int heat2d_tiramisu(halide_buffer_t *in, halide_buffer_t *out);
/* benchmarks/halide/wrapper_heat2dgpu.cpp:44 — sizes i32[2] = {extent0 (W), extent1 (H)}. */
int heat2dgpu_tiramisu(halide_buffer_t *sizes, halide_buffer_t *in, halide_buffer_t *out);
/* benchmarks/halide/wrapper_heat2d_dist.h:23 / heat2d_dist_tiramisu.cpp:187-188,253 — one call per rank
 * (communicator up: tiramisu_MPI_init), buff_out_even f32[rows+2][M], buff_out_odd f32[rows+2][M] with local
 * rows 1..rows owned and rows 0 / rows+1 halos; ITERS even/odd Jacobi sweeps (form 2), result in buff_out_odd.
 * Only ->host is read (the reference's wrapper leaves every other field uninitialised).  Sizes default to the
 * reference's N = 3842, M = 10000 split over the communicator; see tiramisu_b200_set_heat2d_dist. */
int heat2d_dist_tiramisu(halide_buffer_t *buff_out_even, halide_buffer_t *buff_out_odd);
/* rows_per_rank (0: (3842-2)/world), cols (0: keep), iters (0: keep; default 2 = TOTAL_ITERATIONS), protocol:
 * 0 = the generator's (ITERS even, the first sweep reads buff_out_odd), 1 = the wrapper's (the first sweep reads
 * the first argument; ITERS = 1 gives out = sweep(in), what wrapper_heat2d_dist.cpp:101-111 checks).
 * Environment overrides: TB200_HEAT2D_DIST_ROWS / _COLS / _ITERS / _PROTOCOL. */
int tiramisu_b200_set_heat2d_dist(int32_t rows_per_rank, int32_t cols, int32_t iters, int32_t protocol);
/* benchmarks/halide/wrapper_blurxy.cpp:21 — in u8[3][H][W], out u8[3][H-8][W-8]. */
int blurxy_tiramisu(halide_buffer_t *in, halide_buffer_t *out);
/* fp32 blur of BASELINE config 2 (SURVEY.md §8 a8'): in f32[H][W], out f32[H-2][W-2]. */
int blurxy_f32_tiramisu(halide_buffer_t *in, halide_buffer_t *out);
/* tests/wrapper_test_160.h / benchmarks/automatic_comm: in u32[rows(+2)][cols+2], out u32[rows][cols].
 * With a communicator up (tiramisu_MPI_init) the 2 halo rows are received from rank+1. */
int boxblur(halide_buffer_t *in, halide_buffer_t *out);
int blurautodist_tiramisu(halide_buffer_t *in, halide_buffer_t *out);
/* benchmarks/linear_algebra/blas/level3/sgemm/cpu/sgemm_wrapper.h:13 — SIZES i32[3]={N,M,K},
 * C[NxM] = a*(A[NxK]*B[KxM]) + b*C. */
int sgemm_tiramisu(halide_buffer_t *SIZES, halide_buffer_t *alpha, halide_buffer_t *beta,
                   halide_buffer_t *A, halide_buffer_t *B, halide_buffer_t *C);
/* benchmarks/linear_algebra/blas/level3/sgemm/gpu/wrapper.h:6-11 — Consts f32[2]={alpha,beta};
 * time_start/time_end f32[1] receive get_time() around the kernel when the caller provides it. */
int matmul(halide_buffer_t *Consts, halide_buffer_t *A, halide_buffer_t *B, halide_buffer_t *C,
           halide_buffer_t *time_start, halide_buffer_t *time_end);
/* benchmarks/DNN/blocks/Conv-ReLU-MaxPool/cpu/dense/wrapper_nn_block.cpp:43-48 */
int conv_relu_maxpool_block(halide_buffer_t *input, halide_buffer_t *filter, halide_buffer_t *bias,
                            halide_buffer_t *output);
/* benchmarks/DNN/blocks/LSTM/cpu/wrapper.h:6-10 — params i32[4] = {F, B, L, T} */
int lstm(halide_buffer_t *params, halide_buffer_t *Weights, halide_buffer_t *biases, halide_buffer_t *x,
         halide_buffer_t *y);
/* The reference's other two LSTM benchmarks also export `lstm`, with other arities; here they carry
 * distinct names, and lib/libtiramisu_b200_lstm_{gpu,cpu}_lib.so re-export them as `lstm` (csrc/abi/).
 * LSTM/gpu_lib/wrapper.cpp:74-79: Weights {F, 4F, 2, L} (0 = W, 1 = R; gates i, f, z, o), time_* f32[1]. */
int tiramisu_b200_lstm_gpu_lib(halide_buffer_t *Weights, halide_buffer_t *biases, halide_buffer_t *x, halide_buffer_t *y,
                               halide_buffer_t *time_start, halide_buffer_t *time_end);
/* LSTM/cpu_lib/wrapper.cpp:65-72: weights {4F, F, 2, L} (transposed), h {F, B, T+1} and c {F, B} working buffers. */
int tiramisu_b200_lstm_cpu_lib(halide_buffer_t *weights, halide_buffer_t *biases, halide_buffer_t *x, halide_buffer_t *h,
                               halide_buffer_t *c, halide_buffer_t *out);
/* tests/wrapper_test_162.h..167.h — cublas_gemm / gpu_tile GEMM feature tests. */
int test_162(halide_buffer_t *sizes, halide_buffer_t *params, halide_buffer_t *A,
             halide_buffer_t *B, halide_buffer_t *C);
int test_164(halide_buffer_t *sizes, halide_buffer_t *params, halide_buffer_t *transposes,
             halide_buffer_t *A, halide_buffer_t *B, halide_buffer_t *C);
int test_167(halide_buffer_t *A, halide_buffer_t *B, halide_buffer_t *C);
int test_168(halide_buffer_t *A, halide_buffer_t *B);
int test_197(halide_buffer_t *A, halide_buffer_t *B);

#ifdef __cplusplus
}
#endif
#endif /* TIRAMISU_B200_H */
