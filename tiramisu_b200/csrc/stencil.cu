// stencil.cu — HBM-bound 2D stencils of the hot path: heat2d (5-point) and the 3x3 separable
// blurs (fp32 / uint8 / uint32).  Replaces the kernels the reference's cuda_ast emitter prints
// for benchmarks/halide/heat2dgpu_tiramisu.cpp (32x32 smem tiles, 30x30 useful, scalar loads)
// and restates the CPU-only blurxy / boxblur computations for the GPU.
//
// Design (sm_100a, HBM-bound): one thread owns a 4-wide column strip for RY consecutive output
// rows; it issues all RY+2 128-bit row loads up front (memory-level parallelism), gets the
// east/west neighbours of the centre row by warp shuffle (only the two edge lanes of a warp
// touch memory again), and streams 128-bit results out with evict-first stores.  Input is read
// ~(RY+2)/RY times from L2 and once from HBM; output is written once.  No shared memory: the
// reuse window (3 rows x 4 columns) fits in registers.
//
// Arithmetic is done with __fmul_rn/__fadd_rn/__fdiv_rn in the reference's association order so
// that fp32 results are bit-identical to the x86 path (which has no FMA,
// src/tiramisu_codegen_halide.cpp:3988-3995).
#include "common.cuh"
#include <cstdlib>
#include "tiramisu_b200.h"

using namespace tb200;

namespace {

constexpr float kAlpha = 0.3f;   // heat2d_tiramisu.cpp:26
constexpr float kBeta = 0.4f;    // heat2d_tiramisu.cpp:27

template <int FORM>
__device__ __forceinline__ float heat_point(float c, float n, float s, float w, float e)
{
    if (FORM == 0) {            // heat2d_tiramisu.cpp:68   a*c + b*(((E+W)+S)+N)
        float t = __fadd_rn(__fadd_rn(__fadd_rn(e, w), s), n);
        return __fadd_rn(__fmul_rn(kAlpha, c), __fmul_rn(kBeta, t));
    } else if (FORM == 1) {     // heat2dgpu_tiramisu.cpp:54 c*a + (((S+E)+N)+W)*b
        float t = __fadd_rn(__fadd_rn(__fadd_rn(s, e), n), w);
        return __fadd_rn(__fmul_rn(c, kAlpha), __fmul_rn(t, kBeta));
    } else {                    // heat2d_dist_tiramisu.cpp:88 c*a + (((N+S)+W)+E)*b
        float t = __fadd_rn(__fadd_rn(__fadd_rn(n, s), w), e);
        return __fadd_rn(__fmul_rn(c, kAlpha), __fmul_rn(t, kBeta));
    }
}

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// ---- row-sharded grids: halo rows read in place from the neighbours' memory ------------------------
// A rank owns rows [lo, hi) of the global grid.  Row lo-1 lives in the upper neighbour's HBM and rows
// hi, hi+1 in the lower neighbour's; both allocations are mapped into this process (CUDA IPC over
// NVLink, comm.cu: tiramisu_b200_comm_alloc_symmetric), so the stencil kernel fetches them with
// ordinary loads while it streams its own rows — the halo "exchange" of heat2d_dist / blurautodist
// (benchmarks/halide/heat2d_dist_tiramisu.cpp:133-153: four MPI transfers per sweep) is fused into
// the sweep and costs no launch and no staging buffer.
// For iterated sweeps (even/odd Jacobi) the same launch carries the neighbour handshake: the CTAs
// of the first / last row block wait until the neighbour has finished the previous sweep (its
// boundary rows are written, and it no longer reads ours), and the last of them to retire publishes
// this rank's sweep count into the neighbours' flag words with a system-scope release.
using Halo = tiramisu_b200_halo_t;   // include/tiramisu_b200.h

template <bool PEER>
__device__ __forceinline__ const float *halo_row(const float *in, const Halo &h, int y, size_t Ws, bool &remote)
{
    remote = false;
    if (PEER) {
        if (h.up && y < h.lo) { remote = true; return h.up + (ptrdiff_t)(y - (h.lo - 1)) * (ptrdiff_t)Ws; }
        if (h.dn && y >= h.hi) { remote = true; return h.dn + (size_t)(y - h.hi) * Ws; }
    }
    return in + (size_t)y * Ws;
}
// Rows owned by this rank go through the non-coherent path; a neighbour's rows were written by
// another GPU (possibly during this very stream of sweeps), so they are fetched with
// cache-volatile loads, ordered after the acquire in halo_wait.
__device__ __forceinline__ float4 ld_row4(const float *p, bool remote)
{
    return remote ? __ldcv(reinterpret_cast<const float4 *>(p)) : ldg4(p);
}
__device__ __forceinline__ float ld_row1(const float *p, bool remote) { return remote ? __ldcv(p) : __ldg(p); }

__device__ __forceinline__ int ld_acquire_sys(const int *p)
{
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int *p, int v)
{
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// first / last row block of the launch: wait for the neighbour's previous sweep (bounded: a protocol
// bug must trap, not hang the box)
__device__ __forceinline__ void halo_wait(const Halo &h, bool first, bool last)
{
    if (h.flags == nullptr) return;
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        if (first && h.up) while (ld_acquire_sys(h.flags + 0) < h.wait_value) if (clock64() - t0 > 4000000000LL) __trap();
        if (last && h.dn) while (ld_acquire_sys(h.flags + 1) < h.wait_value) if (clock64() - t0 > 4000000000LL) __trap();
    }
    __syncthreads();
}

__device__ __forceinline__ void halo_signal(const Halo &h, bool first, bool last)
{
    if (h.flags == nullptr) return;
    __syncthreads();                                   // every store of this CTA is issued
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (first && h.sig_up && atomicAdd(h.flags + 2, 1) == (int)gridDim.x - 1) {
            h.flags[2] = 0;
            st_release_sys(h.sig_up, h.signal_value);
        }
        if (last && h.sig_dn && atomicAdd(h.flags + 3, 1) == (int)gridDim.x - 1) {
            h.flags[3] = 0;
            st_release_sys(h.sig_dn, h.signal_value);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// heat2d, vector path: W % 4 == 0, 16-byte aligned rows.  grid = (ceil(W/4/128), ceil(rows/RY)).
template <int FORM, int RY, bool PEER>
__device__ __forceinline__ void heat2d_v4_body(const float *__restrict__ in, float *__restrict__ out, int H, int W,
                                               int row_begin, int row_end, const Halo &halo)
{
    const int lane = threadIdx.x & 31;
    const int x = (blockIdx.x * 128 + threadIdx.x) * 4;
    const int y0 = row_begin + blockIdx.y * RY;
    const bool col_ok = x < W;
    const size_t Ws = (size_t)W;

    // rows y0-1 .. y0+RY, all loads in flight before first use
    float4 v[RY + 2];
#pragma unroll
    for (int r = 0; r < RY + 2; r++) {
        const int y = y0 - 1 + r;
        bool remote;
        const float *row = halo_row<PEER>(in, halo, y, Ws, remote);
        v[r] = (col_ok && y >= 0 && y < H && y <= row_end) ? ld_row4(row + x, remote) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // west/east halo of the centre rows for the warp's edge lanes
    float wl[RY], er[RY];
#pragma unroll
    for (int r = 0; r < RY; r++) {
        const int y = y0 + r;
        const bool row_ok = y < row_end && y < H;
        wl[r] = (lane == 0 && col_ok && row_ok && x > 0) ? __ldg(in + (size_t)y * Ws + x - 1) : 0.f;
        er[r] = (lane == 31 && col_ok && row_ok && x + 4 < W) ? __ldg(in + (size_t)y * Ws + x + 4) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < RY; r++) {
        const int y = y0 + r;
        const float4 c = v[r + 1], n = v[r], s = v[r + 2];
        float wv = __shfl_up_sync(0xffffffffu, c.w, 1);
        float ev = __shfl_down_sync(0xffffffffu, c.x, 1);
        if (lane == 0) wv = wl[r];
        if (lane == 31) ev = er[r];
        if (!col_ok || y >= row_end || y >= H) continue;   // (warp-uniform in y; no barrier below depends on it)
        float4 o;
        o.x = heat_point<FORM>(c.x, n.x, s.x, wv, c.y);
        o.y = heat_point<FORM>(c.y, n.y, s.y, c.x, c.z);
        o.z = heat_point<FORM>(c.z, n.z, s.z, c.y, c.w);
        o.w = heat_point<FORM>(c.w, n.w, s.w, c.z, ev);
        float4 *dst = reinterpret_cast<float4 *>(out + (size_t)y * Ws + x);
        if (FORM == 2) {
            // Jacobi half-step: border cells of `out` keep their previous value
            if (y == 0 || y == H - 1) continue;
            if (x == 0 || x + 4 == W) {
                float4 old = *dst;
                if (x == 0) o.x = old.x;
                if (x + 4 == W) o.w = old.w;
            }
        } else {
            if (y == 0 || y == H - 1) o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (x == 0) o.x = 0.f;
            if (x + 4 == W) o.w = 0.f;
        }
        __stcs(dst, o);
    }
}

// Only the first and the last row block of a launch can touch a neighbour's rows: they take the
// halo-aware instantiation (per-row pointer select, volatile loads, handshake); every other CTA runs
// the instruction stream of the single-GPU kernel unchanged.
template <int FORM, int RY, bool PEER = false>
__global__ void __launch_bounds__(128)
tiramisu_sm100_heat2d_v4(const float *__restrict__ in, float *__restrict__ out, int H, int W,
                         int row_begin, int row_end, const Halo halo = Halo{})
{
    if (PEER) {
        const bool first = blockIdx.y == 0, last = blockIdx.y == gridDim.y - 1;
        if (first || last) {
            halo_wait(halo, first, last);
            heat2d_v4_body<FORM, RY, true>(in, out, H, W, row_begin, row_end, halo);
            halo_signal(halo, first, last);
            return;
        }
    }
    heat2d_v4_body<FORM, RY, false>(in, out, H, W, row_begin, row_end, halo);
}

// heat2d, generic path (any W, any alignment): one thread per output element.
template <int FORM>
__global__ void __launch_bounds__(256)
tiramisu_sm100_heat2d_scalar(const float *__restrict__ in, float *__restrict__ out, int H, int W,
                             int row_begin, int row_end)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = row_begin + blockIdx.y;
    if (x >= W || y >= row_end) return;
    const size_t i = (size_t)y * W + x;
    const bool border = (y == 0 || y == H - 1 || x == 0 || x == W - 1);
    if (border) {
        if (FORM != 2) out[i] = 0.f;
        return;
    }
    out[i] = heat_point<FORM>(__ldg(in + i), __ldg(in + i - W), __ldg(in + i + W), __ldg(in + i - 1),
                              __ldg(in + i + 1));
}

// ------------------------------------------------------------------------------------------------
// fp32 blur (SURVEY.md §8 a8', restating blurxy_tiramisu.cpp:61,77 in fp32):
//   bx  = ((in[y][x] + in[y][x+1]) + in[y][x+2]) / 3.0f ;  out = ((bx[y] + bx[y+1]) + bx[y+2]) / 3.0f
// Vector path: W % 4 == 0 and 16-byte aligned input rows; output rows are (W-2) wide, so stores are
// 8-byte aligned pairs.  A thread owns input columns [x, x+4) and produces output columns [x, x+4).
// x / 3.0f, correctly rounded, in 3 FMA-pipe instructions instead of the ~10-instruction IEEE
// division sequence (which made the blur kernel issue-bound: ncu r01 sm__throughput 80 %).
// q = RN(x*RN(1/3)); r = x - 3q (exact in an FMA); q' = RN(q + r*RN(1/3)).  Verified bit-identical
// to x / 3.0f for every finite fp32 including denormals (tools/verify_div3.c, exhaustive).
__device__ __forceinline__ float div3(float x)
{
    const float y = 0.3333333432674407958984375f;      // RN(1/3)
    const float q = __fmul_rn(x, y);
    const float r = __fmaf_rn(-3.0f, q, x);
    const float q2 = __fmaf_rn(r, y, q);
    return (fabsf(x) < __int_as_float(0x7f800000)) ? q2 : q;   // inf/nan: r is NaN, q is already right
}

__device__ __forceinline__ float blur3(float a, float b, float c)
{
    return div3(__fadd_rn(__fadd_rn(a, b), c));
}

template <int RY, bool PEER>
__device__ __forceinline__ void blur_f32_v4_body(const float *__restrict__ in, float *__restrict__ out, int H, int W,
                                                 int row_begin, int row_end, const Halo &halo)
{
    const int lane = threadIdx.x & 31;
    const int x = (blockIdx.x * 128 + threadIdx.x) * 4;
    const int y0 = row_begin + blockIdx.y * RY;
    const int OW = W - 2;
    const bool col_ok = x < W;
    const size_t Ws = (size_t)W;

    float4 v[RY + 2];
    float e0[RY + 2], e1[RY + 2];
#pragma unroll
    for (int r = 0; r < RY + 2; r++) {
        const int y = y0 + r;
        const bool ok = col_ok && y < H && y < row_end + 2;
        bool remote;
        const float *row = halo_row<PEER>(in, halo, y, Ws, remote);
        v[r] = ok ? ld_row4(row + x, remote) : make_float4(0.f, 0.f, 0.f, 0.f);
        const bool edge = ok && lane == 31 && x + 4 < W;      // W % 4 == 0 => x+5 < W too
        e0[r] = edge ? ld_row1(row + x + 4, remote) : 0.f;
        e1[r] = edge ? ld_row1(row + x + 5, remote) : 0.f;
    }
    float4 bx[RY + 2];
#pragma unroll
    for (int r = 0; r < RY + 2; r++) {
        float n0 = __shfl_down_sync(0xffffffffu, v[r].x, 1);
        float n1 = __shfl_down_sync(0xffffffffu, v[r].y, 1);
        if (lane == 31) { n0 = e0[r]; n1 = e1[r]; }
        bx[r].x = blur3(v[r].x, v[r].y, v[r].z);
        bx[r].y = blur3(v[r].y, v[r].z, v[r].w);
        bx[r].z = blur3(v[r].z, v[r].w, n0);
        bx[r].w = blur3(v[r].w, n0, n1);
    }
    // Output rows are W-2 wide: every other row starts 8 bytes off a 16-byte boundary.  Aligned rows take one
    // 128-bit store per thread; on the others the thread's window is shifted by two columns — it stores
    // {o.z, o.w, next lane's o.x, o.y}, lane 0 adds its own first pair and lane 31 keeps its last pair as 64-bit
    // stores — so all but two lanes of a warp still issue ONE 128-bit store (r01: two 64-bit stores per thread
    // on every row left this kernel at 0.94 of the copy peak where heat2d, with 128-bit stores, reaches 1.07).
#pragma unroll
    for (int r = 0; r < RY; r++) {
        const int y = y0 + r;
        if (y >= row_end || y >= H - 2) continue;             // warp-uniform: the shuffles below stay converged
        float4 o;
        o.x = blur3(bx[r].x, bx[r + 1].x, bx[r + 2].x);
        o.y = blur3(bx[r].y, bx[r + 1].y, bx[r + 2].y);
        o.z = blur3(bx[r].z, bx[r + 1].z, bx[r + 2].z);
        o.w = blur3(bx[r].w, bx[r + 1].w, bx[r + 2].w);
        const float nx = __shfl_down_sync(0xffffffffu, o.x, 1);
        const float ny = __shfl_down_sync(0xffffffffu, o.y, 1);
        if (!col_ok) continue;
        float *dst = out + (size_t)y * OW + x;
        if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
            if (x + 3 < OW) __stcs(reinterpret_cast<float4 *>(dst), o);
            else __stcs(reinterpret_cast<float2 *>(dst), make_float2(o.x, o.y));          // last thread of the row: 2 columns
        } else {
            if (lane == 0) __stcs(reinterpret_cast<float2 *>(dst), make_float2(o.x, o.y));
            if (lane < 31 && x + 5 < OW) __stcs(reinterpret_cast<float4 *>(dst + 2), make_float4(o.z, o.w, nx, ny));
            else if (x + 3 < OW) __stcs(reinterpret_cast<float2 *>(dst + 2), make_float2(o.z, o.w));
        }
    }
}

template <int RY, bool PEER = false>
__global__ void __launch_bounds__(128, PEER ? 8 : 0)      // the halo-aware copy of the body must not cost the interior CTAs occupancy
tiramisu_sm100_blur_f32_v4(const float *__restrict__ in, float *__restrict__ out, int H, int W,
                           int row_begin, int row_end, const Halo halo = Halo{})
{
    // row blocks whose RY+2 input rows reach past the owned block (the last one, and the one before it
    // when the last block is short) read the lower neighbour's rows
    if (PEER && row_begin + (int)blockIdx.y * RY + RY + 2 > halo.hi) {
        blur_f32_v4_body<RY, true>(in, out, H, W, row_begin, row_end, halo);
        return;
    }
    blur_f32_v4_body<RY, false>(in, out, H, W, row_begin, row_end, halo);
}

__global__ void __launch_bounds__(256)
tiramisu_sm100_blur_f32_scalar(const float *__restrict__ in, float *__restrict__ out, int H, int W,
                               int row_begin, int row_end)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = row_begin + blockIdx.y;
    const int OW = W - 2;
    if (x >= OW || y >= row_end || y >= H - 2) return;
    float bx[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const float *q = in + (size_t)(y + r) * W + x;
        bx[r] = blur3(__ldg(q), __ldg(q + 1), __ldg(q + 2));
    }
    out[(size_t)y * OW + x] = blur3(bx[0], bx[1], bx[2]);
}

// ------------------------------------------------------------------------------------------------
// uint8 blurxy (blurxy_tiramisu.cpp:61,77): every + and / is uint8 arithmetic (sums wrap mod 256).
// One thread produces 4 consecutive output pixels of one row (6 input bytes from each of 3 rows).
__device__ __forceinline__ unsigned blur3_u8(unsigned a, unsigned b, unsigned c)
{
    return (((((a + b) & 0xffu) + c) & 0xffu) / 3u);
}

__global__ void __launch_bounds__(256)
tiramisu_sm100_blurxy_u8(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int C, int H, int W)
{
    const int OH = H - 8, OW = W - 8;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    const int c = blockIdx.z;
    if (x >= OW || y >= OH || c >= C) return;
    const int nx = min(4, OW - x);
    unsigned bx[3][4];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const uint8_t *q = in + ((size_t)c * H + (y + r)) * W + x;
        unsigned p[6];
#pragma unroll
        for (int k = 0; k < 6; k++) p[k] = (k < nx + 2) ? (unsigned)__ldg(q + k) : 0u;
#pragma unroll
        for (int k = 0; k < 4; k++) bx[r][k] = blur3_u8(p[k], p[k + 1], p[k + 2]);
    }
    uint8_t *o = out + ((size_t)c * OH + y) * OW + x;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (k < nx) o[k] = (uint8_t)blur3_u8(bx[0][k], bx[1][k], bx[2][k]);
}

// uint32 box blur, vertical pass first (tests/test_160.cpp:23-29):
//   bx[i][j] = (in[i][j]+in[i+1][j]+in[i+2][j])/3 ; out[i][j] = (bx[i][j]+bx[i][j+1]+bx[i][j+2])/3
__global__ void __launch_bounds__(256)
tiramisu_sm100_boxblur_u32(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int rows, int cols)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= cols || i >= rows) return;
    const int IW = cols + 2;
    uint32_t bx[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const uint32_t *q = in + (size_t)i * IW + j + d;
        bx[d] = ((__ldg(q) + __ldg(q + IW)) + __ldg(q + 2 * (size_t)IW)) / 3u;
    }
    out[(size_t)i * cols + j] = ((bx[0] + bx[1]) + bx[2]) / 3u;
}

__global__ void __launch_bounds__(256) tiramisu_sm100_fill_i32(int32_t *buf, uint64_t n, int32_t v)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        buf[i] = v;
}

// 32x32 smem-tiled transpose (tests/test_168.cpp: B[j][i] = A[i][j]).
__global__ void __launch_bounds__(256)
tiramisu_sm100_transpose_f32(const float *__restrict__ A, float *__restrict__ At, int rows, int cols)
{
    __shared__ float tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int y = by + r, x = bx + threadIdx.x;
        if (y < rows && x < cols) tile[r][threadIdx.x] = A[(size_t)y * cols + x];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int y = bx + r, x = by + threadIdx.x;     // transposed coordinates
        if (y < cols && x < rows) At[(size_t)y * rows + x] = tile[threadIdx.x][r];
    }
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned8(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }

constexpr int kHeatRY = 8;
constexpr int kBlurRY = 6;

template <int FORM>
int launch_heat2d(const float *in, float *out, int H, int W, int rb, int re, const Halo *halo = nullptr)
{
    const int rows = re - rb;
    if (rows <= 0) return 0;
    if (W % 4 == 0 && aligned16(in) && aligned16(out)) {
        dim3 grid((W / 4 + 127) / 128, (rows + kHeatRY - 1) / kHeatRY);
        if (halo) {
            tiramisu_sm100_heat2d_v4<FORM, kHeatRY, true><<<grid, 128, 0, stream()>>>(in, out, H, W, rb, re, *halo);
            return post_launch("tiramisu_sm100_heat2d_v4<peer>");
        }
        tiramisu_sm100_heat2d_v4<FORM, kHeatRY><<<grid, 128, 0, stream()>>>(in, out, H, W, rb, re);
        return post_launch("tiramisu_sm100_heat2d_v4");
    }
    if (halo) return fail("tiramisu_b200_heat2d_peer", "peer halos need W % 4 == 0 and 16-byte aligned rows");
    dim3 grid((W + 255) / 256, rows);
    tiramisu_sm100_heat2d_scalar<FORM><<<grid, 256, 0, stream()>>>(in, out, H, W, rb, re);
    return post_launch("tiramisu_sm100_heat2d_scalar");
}

}  // namespace

TB200_API int tiramisu_b200_heat2d(const float *in, float *out, int32_t H, int32_t W, int32_t form,
                                   int32_t row_begin, int32_t row_end)
{
    if (!in || !out || H < 1 || W < 1 || row_begin < 0 || row_end > H || form < 0 || form > 2)
        return fail(__func__, "bad argument");
    if ((int64_t)row_end - row_begin > 65535LL * kHeatRY && W % 4 != 0)
        return fail(__func__, "grid too tall for the scalar path");
    switch (form) {
    case 0: return launch_heat2d<0>(in, out, H, W, row_begin, row_end);
    case 1: return launch_heat2d<1>(in, out, H, W, row_begin, row_end);
    default: return launch_heat2d<2>(in, out, H, W, row_begin, row_end);
    }
}

static int check_halo(const tiramisu_b200_halo_t *h, int row_begin, int row_end, const char *who)
{
    if (!h) return fail(who, "halo descriptor is null");
    if (h->lo > row_begin || h->hi < row_end || h->lo >= h->hi) return fail(who, "rows to compute lie outside the owned block [lo, hi)");
    if ((h->up && (reinterpret_cast<uintptr_t>(h->up) & 15u)) || (h->dn && (reinterpret_cast<uintptr_t>(h->dn) & 15u)))
        return fail(who, "peer halo rows must be 16-byte aligned");
    if (h->flags == nullptr && (h->sig_up || h->sig_dn)) return fail(who, "signal targets without local flags");
    return 0;
}

TB200_API int tiramisu_b200_heat2d_peer(const float *in, float *out, int32_t H, int32_t W, int32_t form,
                                        int32_t row_begin, int32_t row_end, const tiramisu_b200_halo_t *halo)
{
    if (!in || !out || H < 1 || W < 1 || row_begin < 0 || row_end > H || form < 0 || form > 2)
        return fail(__func__, "bad argument");
    if (check_halo(halo, row_begin, row_end, __func__)) return 1;
    switch (form) {
    case 0: return launch_heat2d<0>(in, out, H, W, row_begin, row_end, halo);
    case 1: return launch_heat2d<1>(in, out, H, W, row_begin, row_end, halo);
    default: return launch_heat2d<2>(in, out, H, W, row_begin, row_end, halo);
    }
}

TB200_API int tiramisu_b200_blur_f32_peer(const float *in, float *out, int32_t H, int32_t W,
                                          int32_t row_begin, int32_t row_end, const tiramisu_b200_halo_t *halo)
{
    if (!in || !out || H < 3 || W < 3 || row_begin < 0 || row_end > H - 2)
        return fail(__func__, "bad argument");
    if (check_halo(halo, row_begin, row_end, __func__)) return 1;
    const int rows = row_end - row_begin;
    if (rows <= 0) return 0;
    if (!(W % 4 == 0 && aligned16(in) && aligned8(out))) return fail(__func__, "peer halos need W % 4 == 0 and aligned rows");
    dim3 grid((W / 4 + 127) / 128, (rows + kBlurRY - 1) / kBlurRY);
    tiramisu_sm100_blur_f32_v4<kBlurRY, true><<<grid, 128, 0, stream()>>>(in, out, H, W, row_begin, row_end, *halo);
    return post_launch("tiramisu_sm100_blur_f32_v4<peer>");
}

TB200_API int tiramisu_b200_blur_f32(const float *in, float *out, int32_t H, int32_t W,
                                     int32_t row_begin, int32_t row_end)
{
    if (!in || !out || H < 3 || W < 3 || row_begin < 0 || row_end > H - 2)
        return fail(__func__, "bad argument");
    const int rows = row_end - row_begin;
    if (rows <= 0) return 0;
    if (W % 4 == 0 && aligned16(in) && aligned8(out)) {
        // rows per thread: tuning knob for bring-up (TB200_BLUR_RY = 4 | 6 | 8 | 10); 6 is the measured best
        static const int ry = [] { const char *e = getenv("TB200_BLUR_RY"); return e ? atoi(e) : kBlurRY; }();
        dim3 grid((W / 4 + 127) / 128, (rows + ry - 1) / ry);
        switch (ry) {
        case 4: tiramisu_sm100_blur_f32_v4<4><<<grid, 128, 0, stream()>>>(in, out, H, W, row_begin, row_end); break;
        case 8: tiramisu_sm100_blur_f32_v4<8><<<grid, 128, 0, stream()>>>(in, out, H, W, row_begin, row_end); break;
        case 10: tiramisu_sm100_blur_f32_v4<10><<<grid, 128, 0, stream()>>>(in, out, H, W, row_begin, row_end); break;
        default:
            grid.y = (rows + kBlurRY - 1) / kBlurRY;
            tiramisu_sm100_blur_f32_v4<kBlurRY><<<grid, 128, 0, stream()>>>(in, out, H, W, row_begin, row_end);
        }
        return post_launch("tiramisu_sm100_blur_f32_v4");
    }
    if (rows > 65535) return fail(__func__, "grid too tall for the scalar path");
    dim3 grid((W - 2 + 255) / 256, rows);
    tiramisu_sm100_blur_f32_scalar<<<grid, 256, 0, stream()>>>(in, out, H, W, row_begin, row_end);
    return post_launch("tiramisu_sm100_blur_f32_scalar");
}

TB200_API int tiramisu_b200_blurxy_u8(const uint8_t *in, uint8_t *out, int32_t C, int32_t H, int32_t W)
{
    if (!in || !out || C < 1 || H < 9 || W < 9 || H - 8 > 65535 || C > 65535)
        return fail(__func__, "bad argument");
    dim3 grid(((W - 8 + 3) / 4 + 255) / 256, H - 8, C);
    tiramisu_sm100_blurxy_u8<<<grid, 256, 0, stream()>>>(in, out, C, H, W);
    return post_launch("tiramisu_sm100_blurxy_u8");
}

TB200_API int tiramisu_b200_boxblur_u32(const uint32_t *in, uint32_t *out, int32_t rows, int32_t cols)
{
    if (!in || !out || rows < 1 || cols < 1 || rows > 65535) return fail(__func__, "bad argument");
    dim3 grid((cols + 255) / 256, rows);
    tiramisu_sm100_boxblur_u32<<<grid, 256, 0, stream()>>>(in, out, rows, cols);
    return post_launch("tiramisu_sm100_boxblur_u32");
}

TB200_API int tiramisu_b200_fill_i32(int32_t *buf, uint64_t n, int32_t value)
{
    if (!buf && n) return fail(__func__, "bad argument");
    if (n == 0) return 0;
    uint64_t want = (n + 255) / 256, cap = (uint64_t)sm_count() * 8;
    const int blocks = (int)(want < cap ? want : cap);
    tiramisu_sm100_fill_i32<<<blocks, 256, 0, stream()>>>(buf, n, value);
    return post_launch("tiramisu_sm100_fill_i32");
}

TB200_API int tiramisu_b200_transpose_f32(const float *A, float *At, int32_t rows, int32_t cols)
{
    if (!A || !At || rows < 1 || cols < 1) return fail(__func__, "bad argument");
    dim3 grid((cols + 31) / 32, (rows + 31) / 32);
    tiramisu_sm100_transpose_f32<<<grid, dim3(32, 8), 0, stream()>>>(A, At, rows, cols);
    return post_launch("tiramisu_sm100_transpose_f32");
}
