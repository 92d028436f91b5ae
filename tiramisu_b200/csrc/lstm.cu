// lstm.cu — the multi-layer LSTM of /root/reference/benchmarks/DNN/blocks/LSTM/{cpu,gpu}/generator.cpp
//   Weights f32[L][2][4F][F] (index 0 = R recurrent, 1 = W input), biases f32[L][4F],
//   x,y f32[T][B][F]; gate rows [0,F)=i sigmoid, [F,2F)=z tanh, [2F,3F)=o sigmoid, [3F,4F)=f sigmoid;
//   c = i*z + f*c_prev; h = tanh(c)*o; y = h of the last layer.            (cpu/generator.cpp:35-53)
//
// Two paths:
//  * exact: fp32 SIMT kernels that follow the reference's single interleaved j loop
//    (s = b; s = (s + R*h_prev) + W*h_below, cpu/generator.cpp:40-41) — any shape.
//  * tensor (F % 32 == 0, B <= 128): per layer
//      1. one large tcgen05 TF32 GEMM for the input contribution of ALL time steps
//         Xg[T*B][4F] = H_below[T*B][F] * Wp^T           (what LSTM/gpu_lib batches over 10 steps, :64-72)
//      2. per time step ONE kernel: recurrent GEMM h_prev[B][F] * Rp^T on tcgen05 with the
//         pointwise gates fused into the epilogue (the reference launches 2 GEMM kernels + 1
//         pointwise kernel per cell, LSTM/gpu/generator.cpp:85-131).  Weight rows are packed
//         gate-interleaved (row 4u+g) so that the i,z,o,f pre-activations of a hidden unit are 4
//         adjacent accumulator columns of the same thread.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tiramisu_b200.h"

#include <cuda_fp16.h>

#include <atomic>
#include <cstdio>
#include <vector>

using namespace tb200;

namespace tb200 {
int gemm_f32_tc(const float *A, const float *B, float *C, uint64_t M, uint64_t N, uint64_t K, float alpha,
                float beta, uint64_t ldA, uint64_t ldB, uint64_t ldC, bool tB);
void gemm_tc_hint(int kernel, int bn);
int gemm_f16_tc(const void *A16, const void *B16, float *C, uint64_t M, uint64_t N, uint64_t K, float alpha, float beta,
                uint64_t ldA, uint64_t ldB, uint64_t ldC, const float *alpha_dev);
}

namespace {

__device__ __forceinline__ float sigmoid_ref(float v) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v))); }

// ------------------------------------------------------------------------------------------------ exact SIMT
// tmp[k][g] = b[g] + sum_j (R[g][j]*hprev[k][j] then W[g][j]*hbelow[k][j]), one warp per (k, 32 g's):
// lanes own consecutive gate rows; j sequential per thread (reference order).
__global__ void __launch_bounds__(256)
tiramisu_sm100_lstm_gates_exact(const float *__restrict__ R, const float *__restrict__ W,
                                const float *__restrict__ b, const float *__restrict__ hprev,
                                const float *__restrict__ hbelow, float *__restrict__ tmp, int F, int B)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (g >= 4 * F || k >= B) return;
    const float *r = R + (size_t)g * F, *w = W + (size_t)g * F;
    const float *hp = hprev + (size_t)k * F, *hb = hbelow + (size_t)k * F;
    float s = b[g];
    for (int j = 0; j < F; j++) {
        s = __fadd_rn(s, __fmul_rn(__ldg(r + j), __ldg(hp + j)));
        s = __fadd_rn(s, __fmul_rn(__ldg(w + j), __ldg(hb + j)));
    }
    tmp[(size_t)k * 4 * F + g] = s;
}

__global__ void __launch_bounds__(256)
tiramisu_sm100_lstm_pointwise(const float *__restrict__ tmp, const float *__restrict__ cprev,
                              float *__restrict__ c, float *__restrict__ h, int F, int B)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * F) return;
    const int k = idx / F, i = idx % F;
    const float *t = tmp + (size_t)k * 4 * F;
    const float gi = sigmoid_ref(t[i]);
    const float gz = tanhf(t[i + F]);
    const float go = sigmoid_ref(t[i + 2 * F]);
    const float gf = sigmoid_ref(t[i + 3 * F]);
    const float cv = __fadd_rn(__fmul_rn(gi, gz), __fmul_rn(gf, cprev[idx]));
    c[idx] = cv;
    h[idx] = __fmul_rn(tanhf(cv), go);
}

// ------------------------------------------------------------------------------------------------ tensor path
// Weights [2][4F][F] -> packed [2][4F][F] with row p = 4*u + g  <-  row g*F + u ; bias likewise.
__global__ void __launch_bounds__(256)
tiramisu_sm100_lstm_pack(const float *__restrict__ Wl, const float *__restrict__ bl, float *__restrict__ Wp,
                         float *__restrict__ bp, int F)
{
    const size_t total4 = (size_t)2 * 4 * F * (F / 4);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
        const int j4 = (int)(i % (F / 4));
        size_t t = i / (F / 4);
        const int p = (int)(t % (4 * F));
        const int m = (int)(t / (4 * F));
        const int u = p >> 2, g = p & 3;
        const float4 v = __ldg(reinterpret_cast<const float4 *>(Wl + ((size_t)m * 4 * F + (size_t)g * F + u) * F) + j4);
        reinterpret_cast<float4 *>(Wp + ((size_t)m * 4 * F + p) * F)[j4] = v;
        if (m == 0 && j4 == 0) bp[p] = bl[g * F + u];
    }
}

constexpr int BM = 128, BK = 32, UMMA_K = 8, BN = 32, STAGES = 8;
constexpr int A_TILE = BM * BK * 4, B_TILE = BN * BK * 4, STAGE_BYTES = A_TILE + B_TILE;
constexpr int kThreads = 192;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;

struct StepParams {
    int B, F;               // batch rows (<= 128), hidden size
    const float *xg;        // [B][4F] packed columns: input contribution of this time step
    const float *bias;      // [4F] packed
    const float *cprev;     // [B][F]
    float *c, *h;           // [B][F]
    float *y;               // optional second destination for h (last layer), may be null
};

// One LSTM cell step: gates[B][4F] = h_prev[B][F] * Rp^T (tcgen05, TF32) ; epilogue = Xg + bias,
// sigmoid/tanh, c/h update.  grid = 4F/32 CTAs, each owning 8 hidden units.
__global__ void __launch_bounds__(kThreads, 1)
tiramisu_sm100_lstm_step_tf32(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_r,
                              const StepParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *smem_a = smem, *smem_b = smem + STAGES * A_TILE;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tmem_full = empty_bar + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN;
    const int num_kb = p.F / BK;

    if (warp == 0 && lane == 0) { ptx::prefetch_tensormap(&map_h); ptx::prefetch_tensormap(&map_r); }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; s++) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
            ptx::mbar_init(tmem_full, 1);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, 32);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        for (int kb = 0; kb < num_kb; kb++) {
            const int s = kb % STAGES;
            ptx::mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
            if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                ptx::tma_load_2d(smem_a + s * A_TILE, &map_h, &full_bar[s], kb * BK, 0);     // rows >= B: zero fill
                ptx::tma_load_2d(smem_b + s * B_TILE, &map_r, &full_bar[s], kb * BK, n0);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // whole warp walks the ring (uniform control flow); one elected lane issues
        constexpr uint32_t idesc = ptx::make_idesc_tf32(BM, BN, false, false);
        const uint64_t da0 = ptx::make_smem_desc(ptx::smem_u32(smem_a), 16, 1024);
        const uint64_t db0 = ptx::make_smem_desc(ptx::smem_u32(smem_b), 16, 1024);
        for (int kb = 0; kb < num_kb; kb++) {
            const int s = kb % STAGES;
            ptx::mbar_wait(&full_bar[s], (kb / STAGES) & 1);
            ptx::tc_fence_after_sync();
            if (ptx::elect_one()) {
                const uint64_t da = ptx::desc_advance(da0, (uint32_t)s * A_TILE), db = ptx::desc_advance(db0, (uint32_t)s * B_TILE);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; k++)
                    ptx::mma_tf32_ss(tmem_base, ptx::desc_advance(da, k * (UMMA_K * 4)),
                                     ptx::desc_advance(db, k * (UMMA_K * 4)), idesc, (kb | k) != 0);
                ptx::mma_commit(&empty_bar[s]);
                if (kb == num_kb - 1) ptx::mma_commit(tmem_full);
            }
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;                 // batch index k
        // operands of the pointwise part do not depend on the MMAs: fetch them while the GEMM runs
        const bool ok = row < p.B;
        float4 xg[8], bs[8], cp[2];
        if (ok) {
            const float4 *xp = reinterpret_cast<const float4 *>(p.xg + (size_t)row * 4 * p.F + n0);
            const float4 *bp = reinterpret_cast<const float4 *>(p.bias + n0);
            const float4 *cq = reinterpret_cast<const float4 *>(p.cprev + (size_t)row * p.F + n0 / 4);
#pragma unroll
            for (int j = 0; j < 8; j++) { xg[j] = __ldg(xp + j); bs[j] = __ldg(bp + j); }
            cp[0] = __ldg(cq); cp[1] = __ldg(cq + 1);
        }
        ptx::mbar_wait(tmem_full, 0);
        ptx::tc_fence_after_sync();
        uint32_t r[32];
        __syncwarp();
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16), r);
        ptx::tmem_ld_wait();
        if (ok) {
            float cn[8], hn[8];
            const float cpv[8] = {cp[0].x, cp[0].y, cp[0].z, cp[0].w, cp[1].x, cp[1].y, cp[1].z, cp[1].w};
#pragma unroll
            for (int u = 0; u < 8; u++) {
                // columns 4u..4u+3 = pre-activations i, z, o, f of hidden unit n0/4 + u
                const float si = __uint_as_float(r[4 * u + 0]) + xg[u].x + bs[u].x;
                const float sz = __uint_as_float(r[4 * u + 1]) + xg[u].y + bs[u].y;
                const float so = __uint_as_float(r[4 * u + 2]) + xg[u].z + bs[u].z;
                const float sf = __uint_as_float(r[4 * u + 3]) + xg[u].w + bs[u].w;
                const float gi = sigmoid_ref(si), gz = tanhf(sz), go = sigmoid_ref(so), gf = sigmoid_ref(sf);
                cn[u] = __fadd_rn(__fmul_rn(gi, gz), __fmul_rn(gf, cpv[u]));
                hn[u] = __fmul_rn(tanhf(cn[u]), go);
            }
            const size_t o = (size_t)row * p.F + n0 / 4;
            reinterpret_cast<float4 *>(p.c + o)[0] = make_float4(cn[0], cn[1], cn[2], cn[3]);
            reinterpret_cast<float4 *>(p.c + o)[1] = make_float4(cn[4], cn[5], cn[6], cn[7]);
            reinterpret_cast<float4 *>(p.h + o)[0] = make_float4(hn[0], hn[1], hn[2], hn[3]);
            reinterpret_cast<float4 *>(p.h + o)[1] = make_float4(hn[4], hn[5], hn[6], hn[7]);
            if (p.y) {
                reinterpret_cast<float4 *>(p.y + o)[0] = make_float4(hn[0], hn[1], hn[2], hn[3]);
                reinterpret_cast<float4 *>(p.y + o)[1] = make_float4(hn[4], hn[5], hn[6], hn[7]);
            }
        }
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem_base, 32);
    }
}


// ------------------------------------------------------------------------------------------------
// Persistent layer kernel: ONE launch runs all T recurrent steps of a layer.
//   grid = 4F/32 CTAs (128 at F = 1024), one per SM, in clusters of CL.
//   * CTA j owns gate columns [32j, 32j+32) = hidden units [8j, 8j+8): its slice of the packed
//     recurrent weights R (32 x F fp32 = 128 KB at F = 1024) is loaded into shared memory ONCE and
//     stays there for the whole layer; the cell state c of its 8 units lives in registers.
//   * per step: h[t-1] (B x F) streams through a shared-memory ring as the A operand (with CL > 1
//     each CTA of the cluster fetches 1/CL of the rows and TMA-multicasts them to its peers, so
//     the all-gather of h costs 1/CL of the L2 reads), tcgen05 TF32 MMAs accumulate the 32 gate
//     pre-activations in TMEM, the epilogue adds the precomputed input contribution Xg + bias,
//     applies the gates and writes h[t] (its 8 columns).
//   * steps are separated by a device-wide barrier on an atomic counter (release/acquire at gpu
//     scope + fence.proxy.async, because the next step reads h through the async proxy = TMA).
// The reference launches 2 GEMM kernels + 1 pointwise kernel per cell (LSTM/gpu/generator.cpp:85-131)
// and r01 of this backend launched one kernel per cell (15 us each, launch + weight re-fetch bound).
constexpr int L_BN = 32;
constexpr int L_RING_BYTES = 64 * 1024;
constexpr int L_PAD_BYTES = 8 * 1024;          // M = 128 MMAs read 16 KB of A even when only 64 rows are live
constexpr int L_MAX_STAGES = 8;

struct LayerParams {
    int B, F, T;
    int a_rows;             // 64 (B <= 64) or 128: rows of h fetched per k-slab
    const float *xg;        // [T*B][4F] packed gate columns: input contribution of every step
    const float *bias;      // [4F] packed
    float *h;               // [T][B][F] output of this layer
    float *c_out;           // null, or [B][F]: the cell state after the last step (LSTM/cpu_lib's `c` argument)
    uint32_t *counter;      // device-wide step barrier, zero at launch
    int dbg;                // timing experiments only (TB200_LSTM_DBG): 1 = producers do not wait for the
                            // step barrier, 2 = no recurrent loads/MMAs at all (results are then wrong)
};

__device__ __forceinline__ void grid_wait(const uint32_t *ctr, uint32_t target)
{
    const long long t0 = clock64();
    while (ptx::ld_acquire_gpu(ctr) < target) {
        if (clock64() - t0 > (1ll << 31)) __trap();       // ~1 s: a protocol bug must not hang the box
    }
}

__device__ __forceinline__ float sigmoid_fast(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }

template <int CL>
__global__ void __launch_bounds__(kThreads, 1)
tiramisu_sm100_lstm_layer_tf32(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_r,
                               const LayerParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int num_kb = p.F / BK;
    uint8_t *smem_r = smem;                                   // num_kb slabs of [32 n][32 k] fp32, 4 KB each
    uint8_t *ring = smem + (size_t)num_kb * 4096;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(ring + L_RING_BYTES + L_PAD_BYTES);
    uint64_t *empty_bar = full_bar + L_MAX_STAGES;
    uint64_t *r_bar = empty_bar + L_MAX_STAGES;
    uint64_t *tmem_full = r_bar + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * L_BN;
    const uint32_t stage_bytes = (uint32_t)p.a_rows * 128u;
    const int stages = L_RING_BYTES / (int)stage_bytes;
    const uint32_t rank = (CL > 1) ? ptx::cluster_ctarank() : 0u;
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);

    if (warp == 0 && lane == 0) { ptx::prefetch_tensormap(&map_h); ptx::prefetch_tensormap(&map_r); }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < L_MAX_STAGES; s++) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], CL); }
            ptx::mbar_init(r_bar, 1);
            ptx::mbar_init(tmem_full, 1);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, 32);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before_sync();
    if (CL > 1) ptx::cluster_sync(); else __syncthreads();    // peers' barriers exist before any multicast lands
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: whole warp, uniform control flow, one elected lane issues =====
        if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(r_bar, (uint32_t)num_kb * 4096u);
            for (int kb = 0; kb < num_kb; kb++) ptx::tma_load_2d(smem_r + kb * 4096, &map_r, r_bar, kb * BK, n0);
        }
        __syncwarp();
        const int rows_per_cta = p.a_rows / CL;
        uint32_t it = 0;
        for (int t = 1; t < p.T; t++) {
            if (!(p.dbg & 1)) grid_wait(p.counter, (uint32_t)gridDim.x * (uint32_t)t);      // every CTA has published h[t-1]
            ptx::fence_proxy_async_all();
            for (int kb = 0; kb < num_kb; kb++, it++) {
                const int s = it % stages;
                ptx::mbar_wait(&empty_bar[s], ((it / stages) & 1) ^ 1);
                if (ptx::elect_one()) {
                    ptx::mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
                    uint8_t *dst = ring + (size_t)s * stage_bytes;
                    if (CL == 1)
                        ptx::tma_load_3d(dst, &map_h, &full_bar[s], kb * BK, 0, t - 1);
                    else
                        ptx::tma_load_3d_multicast(dst + rank * rows_per_cta * 128, &map_h, &full_bar[s], kb * BK,
                                                   (int)rank * rows_per_cta, t - 1, kMask);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: whole warp, uniform control flow, one elected lane issues =====
        constexpr uint32_t idesc = ptx::make_idesc_tf32(BM, L_BN, false, false);
        const uint64_t da0 = ptx::make_smem_desc(ptx::smem_u32(ring), 16, 1024);
        const uint64_t db0 = ptx::make_smem_desc(ptx::smem_u32(smem_r), 16, 1024);
        ptx::mbar_wait(r_bar, 0);
        ptx::tc_fence_after_sync();
        uint32_t it = 0;
        for (int t = 1; t < p.T; t++) {
            for (int kb = 0; kb < num_kb; kb++, it++) {
                const int s = it % stages;
                ptx::mbar_wait(&full_bar[s], (it / stages) & 1);
                ptx::tc_fence_after_sync();
                if (ptx::elect_one()) {
                    const uint64_t da = ptx::desc_advance(da0, (uint32_t)s * stage_bytes);
                    const uint64_t db = ptx::desc_advance(db0, (uint32_t)kb * 4096u);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; k++)
                        ptx::mma_tf32_ss(tmem_base, ptx::desc_advance(da, k * (UMMA_K * 4)),
                                         ptx::desc_advance(db, k * (UMMA_K * 4)), idesc, (kb | k) != 0);
                    if (CL == 1) ptx::mma_commit(&empty_bar[s]); else ptx::mma_commit_multicast(&empty_bar[s], kMask);
                    if (kb == num_kb - 1) ptx::mma_commit(tmem_full);
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue: gates, cell state in registers, h[t] out, publish =====
        const int q = warp & 3;
        const int row = q * 32 + lane;                 // batch index
        const bool ok = row < p.B;
        float c[8];
        float4 bs[8];
#pragma unroll
        for (int u = 0; u < 8; u++) { c[u] = 0.0f; bs[u] = __ldg(reinterpret_cast<const float4 *>(p.bias + n0) + u); }
        for (int t = 0; t < p.T; t++) {
            float4 xg[8];
            if (ok) {
                const float4 *xp = reinterpret_cast<const float4 *>(p.xg + ((size_t)t * p.B + row) * 4 * p.F + n0);
#pragma unroll
                for (int j = 0; j < 8; j++) xg[j] = __ldg(xp + j);
            }
            uint32_t r[32];
            if (t > 0) {
                ptx::mbar_wait(tmem_full, (uint32_t)(t - 1) & 1u);
                ptx::tc_fence_after_sync();
                __syncwarp();
                ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16), r);
                ptx::tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++) r[j] = 0u;       // h[-1] = 0: no recurrent contribution
            }
            if (ok) {
                float hn[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    // columns 4u..4u+3 = pre-activations i, z, o, f of hidden unit n0/4 + u
                    const float si = __uint_as_float(r[4 * u + 0]) + xg[u].x + bs[u].x;
                    const float sz = __uint_as_float(r[4 * u + 1]) + xg[u].y + bs[u].y;
                    const float so = __uint_as_float(r[4 * u + 2]) + xg[u].z + bs[u].z;
                    const float sf = __uint_as_float(r[4 * u + 3]) + xg[u].w + bs[u].w;
                    const float gi = sigmoid_fast(si), gz = ptx::tanh_approx(sz), go = sigmoid_fast(so),
                                gf = sigmoid_fast(sf);
                    c[u] = gi * gz + gf * c[u];
                    hn[u] = ptx::tanh_approx(c[u]) * go;
                }
                float4 *dst = reinterpret_cast<float4 *>(p.h + ((size_t)t * p.B + row) * p.F + n0 / 4);
                dst[0] = make_float4(hn[0], hn[1], hn[2], hn[3]);
                dst[1] = make_float4(hn[4], hn[5], hn[6], hn[7]);
                if (p.c_out && t == p.T - 1) {
                    float4 *cd = reinterpret_cast<float4 *>(p.c_out + (size_t)row * p.F + n0 / 4);
                    cd[0] = make_float4(c[0], c[1], c[2], c[3]);
                    cd[1] = make_float4(c[4], c[5], c[6], c[7]);
                }
                __threadfence();                       // h[t] visible device-wide before the counter moves
            }
            ptx::fence_proxy_async_all();
            ptx::tc_fence_before_sync();               // TMEM reads done before the next step's MMAs overwrite
            asm volatile("bar.sync 1, 128;" ::: "memory");   // the 4 epilogue warps
            if (warp == 2 && lane == 0) ptx::red_release_gpu_add(p.counter, 1u);
        }
    }

    ptx::tc_fence_before_sync();
    if (CL > 1) ptx::cluster_sync(); else __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem_base, 32);
    }
}

// ------------------------------------------------------------------------------------------------
// K-split persistent layer kernel — the default recurrence engine (F % 128 == 0, B <= 64, F/8 <= #SMs).
// The N-split kernel above issues 128 small MMAs (M=128 x N=32) per step behind a 32-slab ring:
// r01 ncu shows the tensor pipe 9 % busy and each step bound by ring round trips + per-slab
// synchronisation (12.8 us/step).  Here the operands are flipped and K is split across a cluster:
//   * cluster j (4 CTAs) owns the 128 packed gate rows [128j, +128) = hidden units [32j, +32);
//     CTA ks of the cluster keeps R[128 rows][K slice ks = F/4 columns] resident in shared memory
//     (128 KB at F = 1024) as the A operand (M = 128);
//   * per step the B operand is h[t-1][:, K slice] (64 x F/4 fp32 = 64 KB): the whole step's input
//     is in flight at once (no ring reuse inside a step), and 4F/128 = 32 MMAs (N = 64) produce
//     the partial pre-activations P_ks[128][64] in TMEM;
//   * the four partials are reduced through distributed shared memory: warp q of every CTA pushes
//     TMEM lane quarter q (gate rows [32q, +32)) into CTA q's receive buffer with st.async
//     (XOR-swizzled rows; every 16-byte store signals complete_tx on CTA q's mbarrier — no fences); CTA q sums them in a fixed order, adds
//     Xg + bias, applies the gates for its 8 units x 64 batch entries (cell state in registers)
//     and writes h[t];
//   * steps are separated by the device-wide counter barrier (one fence per CTA, relaxed polling).
constexpr int KS_CL = 4;
constexpr int KS_RECV_BYTES = 4 * 32 * 64 * 4;     // [src CTA][gate row][batch], 32 KB

struct KSplitParams {
    int B, F, T;
    const float *xg;        // [T*B][4F] packed gate columns
    const float *bias;      // [4F] packed
    float *h;               // [T][B][F]
    float *c_out;           // null, or [B][F]: the cell state after the last step
    uint32_t *counter;
    int dbg;
    unsigned long long *trace;   // optional [grid][T][8] %globaltimer stamps (TB200_LSTM_TRACE, bring-up only)
};

__device__ __forceinline__ unsigned long long globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define KS_TRACE(slot)                                                                              \
    do {                                                                                            \
        if (p.trace) p.trace[((size_t)blockIdx.x * p.T + t) * 8 + (slot)] = globaltimer();          \
    } while (0)

__device__ __forceinline__ void grid_wait_relaxed(const uint32_t *ctr, uint32_t target)
{
    const long long t0 = clock64();
    while (ptx::ld_relaxed_gpu(ctr) < target) {
        if (clock64() - t0 > (1ll << 31)) __trap();
    }
    ptx::fence_acq_rel_gpu();
}

__global__ void __launch_bounds__(kThreads, 1)
tiramisu_sm100_lstm_layer_ksplit(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_r,
                                 const KSplitParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int KB = p.F / (KS_CL * BK);                       // 32-column k-blocks per K slice (<= 8)
    const int half = (KB + 1) / 2;
    uint8_t *smem_r = smem;                                  // KB x [128 rows][32 k]  (16 KB each)
    uint8_t *smem_h = smem_r + (size_t)KB * 16384;           // KB x [64 rows][32 k]   (8 KB each)
    float *recv = reinterpret_cast<float *>(smem_h + (size_t)KB * 8192);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(recv) + KS_RECV_BYTES);   // [2]
    uint64_t *r_bar = full_bar + 2;
    uint64_t *tmem_full = r_bar + 1;
    uint64_t *recv_bar = tmem_full + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(recv_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ks = ptx::cluster_ctarank();              // K slice of this CTA = its rank in the cluster
    const int j = blockIdx.x / KS_CL;                        // gate-row group of the cluster
    const int k_base = (int)ks * (p.F / KS_CL);

    if (warp == 0 && lane == 0) { ptx::prefetch_tensormap(&map_h); ptx::prefetch_tensormap(&map_r); }
    if (warp == 1) {
        if (lane == 0) {
            ptx::mbar_init(&full_bar[0], 1);
            ptx::mbar_init(&full_bar[1], 1);
            ptx::mbar_init(r_bar, 1);
            ptx::mbar_init(tmem_full, 1);
            ptx::mbar_init(recv_bar, 1);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, 64);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before_sync();
    ptx::cluster_sync();                                     // peers' barriers exist before any remote arrive
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(r_bar, (uint32_t)KB * 16384u);
            for (int kb = 0; kb < KB; kb++)
                ptx::tma_load_2d(smem_r + kb * 16384, &map_r, r_bar, k_base + kb * BK, 128 * j);
        }
        __syncwarp();
        for (int t = 1; t < p.T; t++) {
            if (!(p.dbg & 1)) grid_wait_relaxed(p.counter, (uint32_t)gridDim.x * (uint32_t)t);   // h[t-1] is complete
            ptx::fence_proxy_async_all();
            if (ptx::elect_one()) {
                KS_TRACE(0);
                ptx::mbar_arrive_expect_tx(&full_bar[0], (uint32_t)half * 8192u);
                for (int kb = 0; kb < half; kb++)
                    ptx::tma_load_3d(smem_h + kb * 8192, &map_h, &full_bar[0], k_base + kb * BK, 0, t - 1);
                if (KB > half) {
                    ptx::mbar_arrive_expect_tx(&full_bar[1], (uint32_t)(KB - half) * 8192u);
                    for (int kb = half; kb < KB; kb++)
                        ptx::tma_load_3d(smem_h + kb * 8192, &map_h, &full_bar[1], k_base + kb * BK, 0, t - 1);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = ptx::make_idesc_tf32(128, 64, false, false);
        const uint64_t da0 = ptx::make_smem_desc(ptx::smem_u32(smem_r), 16, 1024);
        const uint64_t db0 = ptx::make_smem_desc(ptx::smem_u32(smem_h), 16, 1024);
        ptx::mbar_wait(r_bar, 0);
        ptx::tc_fence_after_sync();
        for (int t = 1; t < p.T; t++) {
            const uint32_t ph = (uint32_t)(t - 1) & 1u;
            ptx::mbar_wait(&full_bar[0], ph);
            ptx::tc_fence_after_sync();
            if (ptx::elect_one()) {
                KS_TRACE(1);
                for (int kb = 0; kb < half; kb++) {
                    const uint64_t da = ptx::desc_advance(da0, (uint32_t)kb * 16384u), db = ptx::desc_advance(db0, (uint32_t)kb * 8192u);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; k++)
                        ptx::mma_tf32_ss(tmem_base, ptx::desc_advance(da, k * (UMMA_K * 4)),
                                         ptx::desc_advance(db, k * (UMMA_K * 4)), idesc, (kb | k) != 0);
                }
                if (KB == half) ptx::mma_commit(tmem_full);
            }
            __syncwarp();
            if (KB > half) {
                ptx::mbar_wait(&full_bar[1], ph);
                ptx::tc_fence_after_sync();
                if (ptx::elect_one()) {
                    KS_TRACE(2);
                    for (int kb = half; kb < KB; kb++) {
                        const uint64_t da = ptx::desc_advance(da0, (uint32_t)kb * 16384u), db = ptx::desc_advance(db0, (uint32_t)kb * 8192u);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; k++)
                            ptx::mma_tf32_ss(tmem_base, ptx::desc_advance(da, k * (UMMA_K * 4)),
                                             ptx::desc_advance(db, k * (UMMA_K * 4)), idesc, true);
                    }
                    ptx::mma_commit(tmem_full);
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue =====
        const uint32_t q = (uint32_t)warp & 3u;        // TMEM lane quarter this warp reads = destination CTA of its rows
        const int e = (int)q * 32 + lane;              // 0..127
        const int b = e & 63;                          // batch entry this thread finalises
        const int ug = e >> 6;                         // which 4 of the CTA's 8 hidden units
        const bool ok = b < p.B;
        const int p0 = 128 * j + 32 * (int)ks + 16 * ug;          // first packed gate column of this thread
        const int u0 = 32 * j + 8 * (int)ks + 4 * ug;             // first hidden unit of this thread
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        float4 bs[4];
#pragma unroll
        for (int u = 0; u < 4; u++) bs[u] = __ldg(reinterpret_cast<const float4 *>(p.bias + p0) + u);
        // where this warp's TMEM rows go: CTA q, slot [src = ks][row = lane]
        const uint32_t push_addr = ptx::mapa_u32(reinterpret_cast<uint8_t *>(recv) + ks * 8192u + (uint32_t)lane * 256u, q);
        const uint32_t push_bar = ptx::mapa_u32(recv_bar, q);
        for (int t = 0; t < p.T; t++) {
            float4 xg[4];
            if (ok) {
                const float4 *xp = reinterpret_cast<const float4 *>(p.xg + ((size_t)t * p.B + b) * 4 * p.F + p0);
#pragma unroll
                for (int u = 0; u < 4; u++) xg[u] = __ldg(xp + u);
            } else {
#pragma unroll
                for (int u = 0; u < 4; u++) xg[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float acc[16];
#pragma unroll
            for (int a = 0; a < 16; a++) acc[a] = 0.0f;
            if (t > 0) {
                // arm this step's receive phase: 4 source CTAs x 32 rows x 256 B arrive as st.async transactions
                if (e == 0) ptx::mbar_arrive_expect_tx(recv_bar, (uint32_t)KS_RECV_BYTES);
                ptx::mbar_wait(tmem_full, (uint32_t)(t - 1) & 1u);
                ptx::tc_fence_after_sync();
                if (e == 0) KS_TRACE(3);
                uint32_t v[64];
                __syncwarp();
                ptx::tmem_ld_32x32_nowait(tmem_base + ((q * 32u) << 16), v);
                ptx::tmem_ld_32x32_nowait(tmem_base + ((q * 32u) << 16) + 32u, v + 32);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before_sync();
                // push gate row (32q + lane) of this CTA's partial to its owner, 16-byte chunks XOR-swizzled by row
#pragma unroll
                for (int ch = 0; ch < 16; ch++)
                    ptx::st_async_cluster_v4(push_addr + (uint32_t)((ch ^ (lane & 7)) * 16), push_bar,
                                             __uint_as_float(v[4 * ch + 0]), __uint_as_float(v[4 * ch + 1]),
                                             __uint_as_float(v[4 * ch + 2]), __uint_as_float(v[4 * ch + 3]));
                if (e == 0) KS_TRACE(4);
                // the four partials of my rows have landed
                ptx::mbar_wait(recv_bar, (uint32_t)(t - 1) & 1u);
                if (e == 0) KS_TRACE(5);
#pragma unroll
                for (int a = 0; a < 16; a++) {
                    const int row = 16 * ug + a;
                    const int col = ((((b >> 2) ^ (row & 7)) << 2) | (b & 3));
                    float sum = recv[0 * 2048 + row * 64 + col];
                    sum += recv[1 * 2048 + row * 64 + col];
                    sum += recv[2 * 2048 + row * 64 + col];
                    sum += recv[3 * 2048 + row * 64 + col];
                    acc[a] = sum;
                }
            }
            float hn[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float si = acc[4 * u + 0] + xg[u].x + bs[u].x;
                const float sz = acc[4 * u + 1] + xg[u].y + bs[u].y;
                const float so = acc[4 * u + 2] + xg[u].z + bs[u].z;
                const float sf = acc[4 * u + 3] + xg[u].w + bs[u].w;
                const float gi = sigmoid_fast(si), gz = ptx::tanh_approx(sz), go = sigmoid_fast(so), gf = sigmoid_fast(sf);
                c[u] = gi * gz + gf * c[u];
                hn[u] = ptx::tanh_approx(c[u]) * go;
            }
            if (ok) {
                *reinterpret_cast<float4 *>(p.h + ((size_t)t * p.B + b) * p.F + u0) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                if (p.c_out && t == p.T - 1) *reinterpret_cast<float4 *>(p.c_out + (size_t)b * p.F + u0) = make_float4(c[0], c[1], c[2], c[3]);
            }
            // publish h[t]: CTA-wide rendezvous of the writers, then ONE gpu-scope fence + counter increment
            if (e == 0) KS_TRACE(6);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (e == 0) {
                ptx::fence_acq_rel_gpu();
                ptx::red_relaxed_gpu_add(p.counter, 1u);
                KS_TRACE(7);
            }
        }
    }

    ptx::tc_fence_before_sync();
    ptx::cluster_sync();                                     // nobody exits while peers may still push into it
    if (warp == 1) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem_base, 64);
    }
}

// ------------------------------------------------------------------------------------------------
// fp16 K-split layer kernel + layer-pair pipeline — the default engine (F % 128 == 0, F <= 1024, B <= 64).
// The TF32 K-split kernel above needs every SM for ONE layer's recurrent weights (16 MB at
// F = 1024), so the L x T cells run strictly one after another (400 dependent steps of ~6.3 us).
// Here R and h are fed to the tensor core as fp16 (kind::f16: 11-bit significands, the same
// operand precision as TF32, fp32 accumulate): a layer's R is 8 MB and fits in the shared memory
// of F/16 = 64 CTAs, so TWO layers' recurrences are resident at once and layer l+1 runs a few
// steps behind layer l:
//   * cluster j (2 CTAs) owns the 128 packed gate rows [128j, +128); CTA ks keeps
//     R16[128 rows][K slice ks = F/2] resident (128 KB) as the A operand; per step the B operand is
//     h16[t-1][:, K slice] (64 x F/2 fp16 = 64 KB, all in flight at once), 32 MMAs (N = 64, K = 16);
//   * the two partials meet through distributed shared memory: TMEM lane half hs goes to CTA hs
//     (st.async + complete_tx), CTA ks finalises gate rows [64 ks, +64) = 16 hidden units x 64 batch
//     entries with 8 epilogue warps, writes h[t] as fp32 (next layer's GEMM, y) and fp16 (own TMA);
//   * steps are separated by the layer's own device-wide counter; the same counter tells the host-side
//     pipeline (a polling kernel on another stream) how far the layer is, and `xg_ready` tells this
//     kernel how many steps of its input contribution Xg the chunk GEMMs have produced so far.
constexpr int H_CL = 2;
constexpr int H_THREADS = 64 + 256;
constexpr int H_RECV_BYTES = 2 * 64 * 64 * 4;      // [src CTA][gate row][batch], 32 KB

struct H16Params {
    int B, F, T;
    const float *xg;            // [T*B][4F] packed gate columns (may still be in production, see xg_ready)
    const float *bias;          // [4F] packed
    float *h;                   // [T][B][F] fp32 output of this layer, or null (only the last layer's is consumed)
    __half *h16;                // [T][B][F] fp16 copy, the recurrence's own operand
    float *c_out;               // null, or [B][F]: the cell state after the last step
    uint32_t *step_flags;       // step barrier words, see sync_mode
    int sync_mode;              // 0: ONE counter, += gridDim.x per finished step (red.add; r01 and the measured best); 3: the same
                                // counter polled with ld.acquire instead of ld.relaxed + fence
                                // 1: [gridDim.x] per-CTA step flags (plain stores; measured slower: r02 run 1, 7.6 vs 5.9 us/step)
                                // 2: 8 counters, CTA c adds to counter c & 7 (shorter same-address atomic queues)
    const float *rscale;        // de-scaling factor of the block-scaled fp16 recurrent weights (a power of two)
    const uint32_t *xg_ready;   // null, or: number of leading time steps of xg that are complete
    int dbg;
    unsigned long long *trace;
    const __half *r16;          // non-null: R16 [4F][F] — the CTA's R slice lives in TENSOR MEMORY (A operand of the MMAs read
                                // from TMEM columns [H_A_COL0, +F/4)) instead of shared memory: the step's MMA phase is bound
                                // by operand bytes through the shared-memory port (~88 B/clk measured, r02 run 22), and R is
                                // 2/3 of them
};
constexpr uint32_t H_A_COL0 = 128;     // first TMEM column of the R slice (accumulator: columns [0, 64))

// sync_mode 3: the single counter polled with acquire loads — no separate fence once the target has been seen
__device__ __forceinline__ void grid_wait_acquire(const uint32_t *ctr, uint32_t target)
{
    const long long t0 = clock64();
    while (ptx::ld_acquire_gpu(ctr) < target) {
        if (clock64() - t0 > (1ll << 31)) __trap();
    }
}

// Every CTA of the layer has published step `target`: lane i watches the flags of CTAs i, i + 32, ...
// (relaxed polls; one acquire fence once all of them have been seen).  Whole warp, converged.
__device__ __forceinline__ void grid_wait_flags(const uint32_t *flags, int n, uint32_t target)
{
    const int lane = threadIdx.x & 31;
    const long long t0 = clock64();
    for (;;) {
        bool ok = true;
        for (int c = lane; c < n; c += 32) ok = ok && ptx::ld_relaxed_gpu(flags + c) >= target;
        if (__all_sync(0xffffffffu, ok)) break;
        if (clock64() - t0 > (1ll << 31)) __trap();
    }
    ptx::fence_acq_rel_gpu();
}

// sync_mode 2: counter k (128 bytes apart) counts the CTAs with (index & 7) == k; lanes 0..7 watch one each
__device__ __forceinline__ void grid_wait_counters8(const uint32_t *ctr, int n, uint32_t step)
{
    const int lane = threadIdx.x & 31;
    const uint32_t mine = lane < 8 ? (uint32_t)((n - lane + 7) / 8) * step : 0u;      // CTAs feeding counter `lane`, times steps
    const long long t0 = clock64();
    for (;;) {
        const bool ok = lane >= 8 || ptx::ld_relaxed_gpu(ctr + 32 * lane) >= mine;
        if (__all_sync(0xffffffffu, ok)) break;
        if (clock64() - t0 > (1ll << 31)) __trap();
    }
    ptx::fence_acq_rel_gpu();
}

__global__ void __launch_bounds__(H_THREADS, 1)
tiramisu_sm100_lstm_layer_h16(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_r,
                              const H16Params p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int KB = p.F / (H_CL * 64);                        // 64-element (128-byte) k-blocks per K slice (<= 8)
    const int half = (KB + 1) / 2;
    uint8_t *smem_r = smem;                                  // KB x [128 rows][64 k] fp16 (16 KB each)
    uint8_t *smem_h = smem_r + (size_t)KB * 16384;           // KB x [64 rows][64 k] fp16  (8 KB each)
    float *recv = reinterpret_cast<float *>(smem_h + (size_t)KB * 8192);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(recv) + H_RECV_BYTES);   // [2]
    uint64_t *r_bar = full_bar + 2;
    uint64_t *tmem_full = r_bar + 1;
    uint64_t *recv_bar = tmem_full + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(recv_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ks = ptx::cluster_ctarank();              // K slice of this CTA = its rank in the cluster
    const int j = blockIdx.x / H_CL;                         // gate-row group of the cluster
    const int k_base = (int)ks * (p.F / H_CL);

    if (warp == 0 && lane == 0) { ptx::prefetch_tensormap(&map_h); ptx::prefetch_tensormap(&map_r); }
    if (warp == 1) {
        if (lane == 0) {
            ptx::mbar_init(&full_bar[0], 1);
            ptx::mbar_init(&full_bar[1], 1);
            ptx::mbar_init(r_bar, p.r16 ? 256 : 1);     // TMEM-resident R: every epilogue thread arrives after its tcgen05.st
            ptx::mbar_init(tmem_full, 1);
            ptx::mbar_init(recv_bar, 1 + 128);          // expect_tx arrival + the 128 lanes that fill the local half
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, p.r16 ? 512 : 64);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before_sync();
    ptx::cluster_sync();                                     // the peer's barriers exist before any remote store
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (!p.r16 && ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(r_bar, (uint32_t)KB * 16384u);
            for (int kb = 0; kb < KB; kb++)
                ptx::tma_load_2d(smem_r + kb * 16384, &map_r, r_bar, k_base + kb * 64, 128 * j);
        }
        __syncwarp();
        for (int t = 1; t < p.T; t++) {
            if (!(p.dbg & 1)) {                                                                   // h16[t-1] is complete
                if (p.sync_mode == 1) grid_wait_flags(p.step_flags, (int)gridDim.x, (uint32_t)t);
                else if (p.sync_mode == 2) grid_wait_counters8(p.step_flags, (int)gridDim.x, (uint32_t)t);
                else if (p.sync_mode == 3) grid_wait_acquire(p.step_flags, (uint32_t)gridDim.x * (uint32_t)t);
                else grid_wait_relaxed(p.step_flags, (uint32_t)gridDim.x * (uint32_t)t);
            }
            ptx::fence_proxy_async_all();
            if (ptx::elect_one()) {
                KS_TRACE(0);
                ptx::mbar_arrive_expect_tx(&full_bar[0], (uint32_t)half * 8192u);
                for (int kb = 0; kb < half; kb++)
                    ptx::tma_load_3d(smem_h + kb * 8192, &map_h, &full_bar[0], k_base + kb * 64, 0, t - 1);
                if (KB > half) {
                    ptx::mbar_arrive_expect_tx(&full_bar[1], (uint32_t)(KB - half) * 8192u);
                    for (int kb = half; kb < KB; kb++)
                        ptx::tma_load_3d(smem_h + kb * 8192, &map_h, &full_bar[1], k_base + kb * 64, 0, t - 1);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = ptx::make_idesc_f16(128, 64, false, false);
        const uint64_t da0 = ptx::make_smem_desc(ptx::smem_u32(smem_r), 16, 1024);
        const uint64_t db0 = ptx::make_smem_desc(ptx::smem_u32(smem_h), 16, 1024);
        const uint32_t a_tmem = p.r16 ? tmem_base + H_A_COL0 : 0u;
        ptx::mbar_wait(r_bar, 0);
        ptx::tc_fence_after_sync();
        for (int t = 1; t < p.T; t++) {
            const uint32_t ph = (uint32_t)(t - 1) & 1u;
            ptx::mbar_wait(&full_bar[0], ph);
            ptx::tc_fence_after_sync();
            if (ptx::elect_one()) {
                KS_TRACE(1);
                for (int kb = 0; kb < half; kb++) {
                    const uint64_t da = ptx::desc_advance(da0, (uint32_t)kb * 16384u), db = ptx::desc_advance(db0, (uint32_t)kb * 8192u);
#pragma unroll
                    for (int k = 0; k < 4; k++) {                // 4 x (K = 16 fp16 = 32 bytes = 8 TMEM columns) per 128-byte row
                        if (a_tmem) ptx::mma_f16_ts(tmem_base, a_tmem + (uint32_t)(kb * 32 + k * 8), ptx::desc_advance(db, k * 32), idesc, (kb | k) != 0);
                        else ptx::mma_f16_ss(tmem_base, ptx::desc_advance(da, k * 32), ptx::desc_advance(db, k * 32), idesc, (kb | k) != 0);
                    }
                }
                if (KB == half) ptx::mma_commit(tmem_full);
            }
            __syncwarp();
            if (KB > half) {
                ptx::mbar_wait(&full_bar[1], ph);
                ptx::tc_fence_after_sync();
                if (ptx::elect_one()) {
                    for (int kb = half; kb < KB; kb++) {
                        const uint64_t da = ptx::desc_advance(da0, (uint32_t)kb * 16384u), db = ptx::desc_advance(db0, (uint32_t)kb * 8192u);
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            if (a_tmem) ptx::mma_f16_ts(tmem_base, a_tmem + (uint32_t)(kb * 32 + k * 8), ptx::desc_advance(db, k * 32), idesc, true);
                            else ptx::mma_f16_ss(tmem_base, ptx::desc_advance(da, k * 32), ptx::desc_advance(db, k * 32), idesc, true);
                        }
                    }
                    ptx::mma_commit(tmem_full);
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue: 8 warps =====
        const uint32_t q = (uint32_t)warp & 3u;            // TMEM lane quarter this warp may read
        const uint32_t ch = (uint32_t)(warp - 2) >> 2;     // which 32 of the 64 batch columns it reads
        const uint32_t dst_cta = q >> 1;                   // gate rows [64 dst, +64) are finalised by CTA dst
        const int e = (warp - 2) * 32 + lane;              // 0..255
        // 4 adjacent lanes finalise the 16 hidden units of one batch entry: their h stores cover whole
        // 32-byte sectors (64 B fp32, 32 B fp16 per batch row).  With one lane per batch row the 512
        // partial-sector writes per step kept the publishing fence waiting ~1.9 us (r01 run 21 timeline).
        const int ug = e & 3;                              // which 4 of the CTA's 16 hidden units
        const int b = e >> 2;                              // batch entry this thread finalises
        const bool ok = b < p.B;
        const int p0 = 128 * j + 64 * (int)ks + 16 * ug;           // first packed gate column of this thread
        const int u0 = 32 * j + 16 * (int)ks + 4 * ug;             // first hidden unit of this thread
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        float4 bs[4];
#pragma unroll
        for (int u = 0; u < 4; u++) bs[u] = __ldg(reinterpret_cast<const float4 *>(p.bias + p0) + u);
        const float rs = __ldg(p.rscale);                  // R was stored as fp16(R * 2^k): accumulators are scaled back
        // where this warp's TMEM rows go: CTA dst, slot [src = ks][row = 32 (q & 1) + lane][batch 32 ch ..]
        const uint32_t row_l = 32u * (q & 1u) + (uint32_t)lane;
        // receive-buffer swizzle: 16-byte chunk c of gate row r lives at chunk c ^ s(r),
        // s(r) = (r & 7) ^ (((r >> 4) & 3) << 1): conflict-free for the row-per-lane writers and for the
        // readers (lanes = 4 unit groups x 8 batch entries -> rows 16 apart, 8 consecutive floats)
        const uint32_t sw_w = (row_l & 7u) ^ (((row_l >> 4) & 3u) << 1);
        const uint32_t push_addr = ptx::mapa_u32(reinterpret_cast<uint8_t *>(recv) + ks * 16384u + row_l * 256u, dst_cta);
        const uint32_t push_bar = ptx::mapa_u32(recv_bar, dst_cta);
        if (p.r16) {
            // R slice -> tensor memory: this thread's TMEM lane is gate row (32 q + lane) of the cluster's 128; warp pair
            // (q, ch) splits the K slice's columns in two.  32 columns = 64 fp16 = 128 bytes of the row per tcgen05.st.
            const __half *rrow = p.r16 + (size_t)(128 * j + 32 * (int)q + lane) * p.F + k_base;
            const int c_lo = ch == 0 ? 0 : 32 * half, c_hi = ch == 0 ? 32 * half : 32 * KB;       // k-blocks [0, half) / [half, KB)
            for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
                uint32_t w[32];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const uint4 v4 = __ldg(reinterpret_cast<const uint4 *>(rrow + 2 * c0) + i);
                    w[4 * i + 0] = v4.x; w[4 * i + 1] = v4.y; w[4 * i + 2] = v4.z; w[4 * i + 3] = v4.w;
                }
                ptx::tmem_st_32x32(tmem_base + ((q * 32u) << 16) + H_A_COL0 + (uint32_t)c0, w);
            }
            ptx::tmem_st_wait();
            ptx::tc_fence_before_sync();
            ptx::mbar_arrive(r_bar);
        }
        uint32_t ready = p.xg_ready ? 0u : 0xffffffffu;
        for (int t = 0; t < p.T; t++) {
            if ((uint32_t)t >= ready) {
                // the chunk GEMM that produces Xg[t] (layer below still running) has not been seen yet
                uint32_t v = 0;
                if (lane == 0) {
                    const long long t0 = clock64();
                    while ((v = ptx::ld_acquire_gpu(p.xg_ready)) <= (uint32_t)t) {
                        if (clock64() - t0 > (1ll << 32)) __trap();
                        __nanosleep(200);
                    }
                }
                ready = __shfl_sync(0xffffffffu, v, 0);
            }
            float4 xg[4];
            if (ok) {
                const float4 *xp = reinterpret_cast<const float4 *>(p.xg + ((size_t)t * p.B + b) * 4 * p.F + p0);
#pragma unroll
                for (int u = 0; u < 4; u++) xg[u] = __ldcg(xp + u);
            } else {
#pragma unroll
                for (int u = 0; u < 4; u++) xg[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float acc[16];
#pragma unroll
            for (int a = 0; a < 16; a++) acc[a] = 0.0f;
            if (t > 0) {
                // arm this step's receive phase: the peer's 64 rows x 256 B arrive as st.async transactions,
                // the local half as plain shared-memory stores + one arrival per writing lane
                if (e == 0) ptx::mbar_arrive_expect_tx(recv_bar, (uint32_t)H_RECV_BYTES / 2);
                ptx::mbar_wait(tmem_full, (uint32_t)(t - 1) & 1u);
                ptx::tc_fence_after_sync();
                if (e == 0) KS_TRACE(3);
                uint32_t v[32];
                __syncwarp();
                ptx::tmem_ld_32x32_nowait(tmem_base + ((q * 32u) << 16) + 32u * ch, v);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before_sync();
                // push batch columns [32 ch, +32) of gate row (32q + lane): 16-byte chunks XOR-swizzled by row
                if (dst_cta == ks) {
                    float4 *lp = reinterpret_cast<float4 *>(reinterpret_cast<uint8_t *>(recv) + ks * 16384u + row_l * 256u);
#pragma unroll
                    for (int cc = 0; cc < 8; cc++)
                        lp[(8u * ch + (uint32_t)cc) ^ sw_w] = make_float4(__uint_as_float(v[4 * cc + 0]), __uint_as_float(v[4 * cc + 1]),
                                                                         __uint_as_float(v[4 * cc + 2]), __uint_as_float(v[4 * cc + 3]));
                    ptx::mbar_arrive(recv_bar);                   // release.cta: the stores above are visible to the waiters
                } else {
#pragma unroll
                    for (int cc = 0; cc < 8; cc++)
                        ptx::st_async_cluster_v4(push_addr + ((8u * ch + (uint32_t)cc) ^ sw_w) * 16u, push_bar,
                                                 __uint_as_float(v[4 * cc + 0]), __uint_as_float(v[4 * cc + 1]),
                                                 __uint_as_float(v[4 * cc + 2]), __uint_as_float(v[4 * cc + 3]));
                }
                if (e == 0) KS_TRACE(4);
                // both partials of my rows have landed
                ptx::mbar_wait(recv_bar, (uint32_t)(t - 1) & 1u);
                if (e == 0) KS_TRACE(5);
#pragma unroll
                for (int a = 0; a < 16; a++) {
                    const int row = 16 * ug + a;
                    const int col = ((((b >> 2) ^ (row & 7) ^ (ug << 1)) << 2) | (b & 3));
                    acc[a] = (recv[0 * 4096 + row * 64 + col] + recv[1 * 4096 + row * 64 + col]) * rs;
                }
            }
            float hn[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float si = acc[4 * u + 0] + xg[u].x + bs[u].x;
                const float sz = acc[4 * u + 1] + xg[u].y + bs[u].y;
                const float so = acc[4 * u + 2] + xg[u].z + bs[u].z;
                const float sf = acc[4 * u + 3] + xg[u].w + bs[u].w;
                const float gi = sigmoid_fast(si), gz = ptx::tanh_approx(sz), go = sigmoid_fast(so), gf = sigmoid_fast(sf);
                c[u] = gi * gz + gf * c[u];
                hn[u] = ptx::tanh_approx(c[u]) * go;
            }
            if (ok) {
                const size_t o = ((size_t)t * p.B + b) * p.F + u0;
                if (p.h) *reinterpret_cast<float4 *>(p.h + o) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                if (p.c_out && t == p.T - 1) *reinterpret_cast<float4 *>(p.c_out + (size_t)b * p.F + u0) = make_float4(c[0], c[1], c[2], c[3]);
                const __half2 lo = __floats2half2_rn(hn[0], hn[1]), hi = __floats2half2_rn(hn[2], hn[3]);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t *>(&lo);
                pk.y = *reinterpret_cast<const uint32_t *>(&hi);
                *reinterpret_cast<uint2 *>(p.h16 + o) = pk;
            }
            // publish h[t]: CTA-wide rendezvous of the writers (also: every reader is done with recv before
            // the next step's pushes are triggered), then ONE gpu-scope fence + this CTA's own step flag
            if (e == 0) KS_TRACE(6);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (e == 0) {
                KS_TRACE(2);
                ptx::fence_acq_rel_gpu();
                if (p.sync_mode == 1) ptx::st_relaxed_gpu(p.step_flags + blockIdx.x, (uint32_t)(t + 1));
                else if (p.sync_mode == 2) ptx::red_relaxed_gpu_add(p.step_flags + 32 * (blockIdx.x & 7), 1u);
                else ptx::red_relaxed_gpu_add(p.step_flags, 1u);
                KS_TRACE(7);
            }
        }
    }

    ptx::tc_fence_before_sync();
    ptx::cluster_sync();                                     // nobody exits while the peer may still push into it
    if (warp == 1) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem_base, p.r16 ? 512 : 64);
    }
}

// ---- block-scaled fp16 weights ----------------------------------------------------------------------
// fp16 has TF32's 11-bit significand but only a 5-bit exponent: |w| > 65504 would become inf and |w| < 6e-8
// zero.  Every weight matrix (R and W of every layer) is therefore stored as fp16(w * 2^k) with k chosen so that
// its largest magnitude lands in [2^13, 2^14); the accumulators are multiplied by 2^-k in the epilogues (exact:
// powers of two).  Overflow cannot happen, every element within 2^-27 of the matrix maximum keeps its 11
// significant bits, and what is smaller still is represented to within 2^-39 of the maximum — far below the
// TF32 rounding (2^-11 relative) of the terms it is added to.
__global__ void __launch_bounds__(256)
tiramisu_sm100_lstm_absmax(const float *__restrict__ W, uint32_t *__restrict__ maxbits, size_t per_matrix4, int n_matrices)
{
    // maxbits[m] = bit pattern of max |w| over matrix m (non-negative floats order like their bit patterns; NaN > inf)
    for (int m = blockIdx.y; m < n_matrices; m += gridDim.y) {
        const float4 *src = reinterpret_cast<const float4 *>(W) + (size_t)m * per_matrix4;
        uint32_t mx = 0;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_matrix4; i += (size_t)gridDim.x * blockDim.x) {
            const float4 v = __ldg(src + i);
            mx = max(max(mx, __float_as_uint(fabsf(v.x))), max(__float_as_uint(fabsf(v.y)), max(__float_as_uint(fabsf(v.z)), __float_as_uint(fabsf(v.w)))));
        }
        mx = __reduce_max_sync(0xffffffffu, mx);
        if ((threadIdx.x & 31) == 0 && mx) atomicMax(maxbits + m, mx);
    }
}

// scale[m] = 2^k, inv[m] = 2^-k; matrix (layer 0, W) stays fp32 (its other operand is the fp32 input x): scale 1
__global__ void tiramisu_sm100_lstm_scales(const uint32_t *__restrict__ maxbits, float *__restrict__ scale, float *__restrict__ inv, int n)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const float mx = __uint_as_float(maxbits[m]);
    int k = 0;
    if (m != 1 && mx > 0.0f && mx < __int_as_float(0x7f800000)) {
        int e;
        frexpf(mx, &e);                                  // mx = f * 2^e, f in [0.5, 1)
        k = 14 - e;
        k = k > 126 ? 126 : (k < -126 ? -126 : k);
    }
    scale[m] = ldexpf(1.0f, k);
    inv[m] = ldexpf(1.0f, -k);
}

// R (fp32 [4F][F], row g*F+u) -> fp16 packed rows p = 4u+g ; W likewise: fp32 for layer 0 (B operand of the
// TF32 input GEMM on x), fp16 for the layers above (their A operand is the fp16 copy of h) ; bias packed.
// scale2: the two power-of-two factors (R, W) of this layer.
__global__ void __launch_bounds__(256)
tiramisu_sm100_lstm_pack16(const float *__restrict__ Wl, const float *__restrict__ bl, __half *__restrict__ R16,
                           float *__restrict__ Wp32, __half *__restrict__ Wp16, float *__restrict__ bp,
                           const float *__restrict__ scale2, int F)
{
    const size_t total4 = (size_t)2 * 4 * F * (F / 4);
    const float sR = __ldg(scale2), sW = __ldg(scale2 + 1);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
        const int j4 = (int)(i % (F / 4));
        size_t t = i / (F / 4);
        const int pr = (int)(t % (4 * F));
        const int m = (int)(t / (4 * F));
        const int u = pr >> 2, g = pr & 3;
        const float4 v = __ldg(reinterpret_cast<const float4 *>(Wl + ((size_t)m * 4 * F + (size_t)g * F + u) * F) + j4);
        const float sc = m == 0 ? sR : sW;
        const __half2 lo = __floats2half2_rn(v.x * sc, v.y * sc), hi = __floats2half2_rn(v.z * sc, v.w * sc);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t *>(&lo);
        pk.y = *reinterpret_cast<const uint32_t *>(&hi);
        if (m == 0) {
            reinterpret_cast<uint2 *>(R16 + (size_t)pr * F)[j4] = pk;
            if (j4 == 0) bp[pr] = bl[g * F + u];
        } else if (Wp16) {
            reinterpret_cast<uint2 *>(Wp16 + (size_t)pr * F)[j4] = pk;
        } else {
            reinterpret_cast<float4 *>(Wp32 + (size_t)pr * F)[j4] = v;
        }
    }
}

// Stream-side halves of the layer pipeline: a one-warp kernel that returns once *ctr >= target (the
// layer below has published enough steps) and a one-thread kernel that publishes a value.
__global__ void tiramisu_sm100_wait_counter(const uint32_t *flags, int n, uint32_t step, int sync_mode)
{
    // returns once all n CTAs of the layer have published `step` steps (same three encodings as H16Params::sync_mode)
    const int words = sync_mode == 1 ? n : (sync_mode == 2 ? 8 : 1);
    for (int c = threadIdx.x; c < words; c += blockDim.x) {
        const uint32_t *w = sync_mode == 2 ? flags + 32 * c : flags + c;
        const uint32_t target = sync_mode == 1 ? step : (sync_mode == 2 ? (uint32_t)((n - c + 7) / 8) * step : (uint32_t)n * step);
        const long long t0 = clock64();
        while (ptx::ld_acquire_gpu(w) < target) {
            if (clock64() - t0 > (1ll << 32)) __trap();
            __nanosleep(500);
        }
    }
}
__global__ void tiramisu_sm100_publish(uint32_t *flag, uint32_t value)
{
    // stream order already guarantees that the producing GEMM has completed; release for the pollers
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

std::atomic<int> g_lstm_mode{0};
std::atomic<int> g_lstm_kernel{0};
thread_local const char *t_last_engine = "";     // which recurrence engine the calling thread's last tiramisu_b200_lstm used
inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int lstm_exact(int F, int B, int L, int T, const float *Wt, const float *bias, const float *x, float *y, float *c_last)
{
    const size_t BF = (size_t)B * F;
    // buf_h[T+1][L+1][B][F], buf_c[T+1][L][B][F] as in cpu/generator.cpp:91-92
    float *h = reinterpret_cast<float *>(pool_alloc((size_t)(T + 1) * (L + 1) * BF * 4));
    float *c = reinterpret_cast<float *>(pool_alloc((size_t)(T + 1) * L * BF * 4));
    float *tmp = reinterpret_cast<float *>(pool_alloc((size_t)B * 4 * F * 4));
    if (!h || !c || !tmp) return 1;
    auto H = [&](int m, int l) { return h + (((size_t)m + 1) * (L + 1) + ((size_t)l + 1)) * BF; };
    auto C = [&](int m, int l) { return c + (((size_t)m + 1) * L + (size_t)l) * BF; };
    TB200_CHECK(cudaMemsetAsync(h, 0, (size_t)(L + 1) * BF * 4, stream()));        // h(-1, l) = 0
    TB200_CHECK(cudaMemsetAsync(c, 0, (size_t)L * BF * 4, stream()));              // c(-1, l) = 0
    for (int m = 0; m < T; m++)                                                     // h(m, -1) = x(m)
        TB200_CHECK(cudaMemcpyAsync(H(m, -1), x + (size_t)m * BF, BF * 4, cudaMemcpyDeviceToDevice, stream()));
    for (int m = 0; m < T; m++)
        for (int l = 0; l < L; l++) {
            const float *R = Wt + ((size_t)l * 2 + 0) * 4 * F * F, *W = Wt + ((size_t)l * 2 + 1) * 4 * F * F;
            dim3 grid((4 * F + 255) / 256, B);
            tiramisu_sm100_lstm_gates_exact<<<grid, 256, 0, stream()>>>(R, W, bias + (size_t)l * 4 * F, H(m - 1, l),
                                                                         H(m, l - 1), tmp, F, B);
            if (post_launch("tiramisu_sm100_lstm_gates_exact")) return 1;
            tiramisu_sm100_lstm_pointwise<<<(int)((BF + 255) / 256), 256, 0, stream()>>>(tmp, C(m - 1, l), C(m, l),
                                                                                        H(m, l), F, B);
            if (post_launch("tiramisu_sm100_lstm_pointwise")) return 1;
        }
    for (int m = 0; m < T; m++)
        TB200_CHECK(cudaMemcpyAsync(y + (size_t)m * BF, H(m, L - 1), BF * 4, cudaMemcpyDeviceToDevice, stream()));
    if (c_last) TB200_CHECK(cudaMemcpyAsync(c_last, C(T - 1, L - 1), BF * 4, cudaMemcpyDeviceToDevice, stream()));
    pool_free(h); pool_free(c); pool_free(tmp);
    return 0;
}


size_t layer_smem_bytes(int F) { return (size_t)(F / BK) * 4096 + L_RING_BYTES + L_PAD_BYTES + 256 + 1024; }

// Can the persistent layer kernel run this shape?  Its grid (4F/32 CTAs, one per SM) must be
// co-resident for the device-wide step barrier, and the R slice + ring must fit in shared memory.
bool layer_kernel_ok(int F, int B)
{
    return F % 32 == 0 && B <= 128 && 4 * F / L_BN <= sm_count() && layer_smem_bytes(F) <= 227 * 1024;
}

template <int CL>
int launch_layer(const CUtensorMap &map_h, const CUtensorMap &map_r, const LayerParams &p, bool probe_only, bool *fits)
{
    auto kern = tiramisu_sm100_lstm_layer_tf32<CL>;
    const size_t smem = layer_smem_bytes(p.F);
    static bool attr_set = false;
    if (!attr_set) {
        TB200_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(4 * p.F / L_BN));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (probe_only) {
        int n = 0;
        if (CL > 1) {
            if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
        } else {
            n = sm_count();
        }
        *fits = n * CL >= (int)cfg.gridDim.x;
        return 0;
    }
    TB200_CHECK(cudaLaunchKernelEx(&cfg, kern, map_h, map_r, p));
    return post_launch("tiramisu_sm100_lstm_layer_tf32");
}

int layer_dispatch(int cl, const CUtensorMap &map_h, const CUtensorMap &map_r, const LayerParams &p, bool probe_only,
                   bool *fits)
{
    switch (cl) {
    case 4: return launch_layer<4>(map_h, map_r, p, probe_only, fits);
    case 2: return launch_layer<2>(map_h, map_r, p, probe_only, fits);
    default: return launch_layer<1>(map_h, map_r, p, probe_only, fits);
    }
}

// Largest cluster size (4, 2, 1) whose whole grid is co-resident; TB200_LSTM_CLUSTER caps it.
int pick_cluster(int F, int a_rows)
{
    int want = 4;
    if (const char *e = getenv("TB200_LSTM_CLUSTER")) want = atoi(e);
    LayerParams probe = {};
    probe.F = F;
    CUtensorMap dummy = {};
    for (int cl = 4; cl >= 2; cl >>= 1) {
        if (cl > want || (4 * F / L_BN) % cl || (a_rows / cl) % 8) continue;
        bool fits = false;
        if (layer_dispatch(cl, dummy, dummy, probe, true, &fits) == 0 && fits) return cl;
    }
    return 1;
}

size_t ksplit_smem_bytes(int F) { return (size_t)(F / (KS_CL * BK)) * (16384 + 8192) + KS_RECV_BYTES + 256 + 1024; }

// K-split engine: shape + co-residency (all F/8 CTAs, clusters of 4, must be resident for the step barrier)
bool ksplit_ok(int F, int B)
{
    if (F % 128 != 0 || B > 64 || F / 8 > sm_count() || ksplit_smem_bytes(F) > 227 * 1024) return false;
    static int cached_F = 0;
    static bool cached = false;
    if (cached_F == F) return cached;
    auto kern = tiramisu_sm100_lstm_layer_ksplit;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(F / 8));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = ksplit_smem_bytes(F);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = KS_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
    cached_F = F;
    cached = n * KS_CL >= F / 8;
    return cached;
}

int launch_ksplit(const CUtensorMap &map_h, const CUtensorMap &map_r, const KSplitParams &p)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(p.F / 8));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = ksplit_smem_bytes(p.F);
    cfg.stream = stream();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = KS_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TB200_CHECK(cudaLaunchKernelEx(&cfg, tiramisu_sm100_lstm_layer_ksplit, map_h, map_r, p));
    return post_launch("tiramisu_sm100_lstm_layer_ksplit");
}

// ---- fp16 engine: host side -----------------------------------------------------------------------
size_t h16_smem_bytes(int F) { return (size_t)(F / (H_CL * 64)) * (16384 + 8192) + H_RECV_BYTES + 256 + 1024; }

cudaLaunchConfig_t h16_config(int F, cudaLaunchAttribute *attr, cudaStream_t st)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(F / 16));
    cfg.blockDim = dim3(H_THREADS);
    cfg.dynamicSmemBytes = h16_smem_bytes(F);
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = H_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cfg;
}

// Shape + co-residency.  *pair_ok: two layers' grids (2 x F/16 CTAs) fit next to each other with
// SMs to spare for the chunk GEMMs.
bool h16_ok(int F, int B, bool *pair_ok)
{
    *pair_ok = false;
    if (F % 128 != 0 || F > 1024 || B > 64 || h16_smem_bytes(F) > 227 * 1024) return false;
    static int cached_F = 0, cached_clusters = 0;
    if (cached_F != F) {
        auto kern = tiramisu_sm100_lstm_layer_h16;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        cudaLaunchAttribute attr[1];
        cudaLaunchConfig_t cfg = h16_config(F, attr, nullptr);
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
        cached_F = F;
        cached_clusters = n;
    }
    const int need = F / 16;
    *pair_ok = cached_clusters * H_CL >= 2 * need + 8 && sm_count() >= 2 * need + 8;
    return cached_clusters * H_CL >= need;
}

// Auxiliary streams + events of the layer pipeline, one set per host thread.
struct PipeStreams {
    cudaStream_t rec = nullptr, gemm = nullptr;
    cudaEvent_t fork = nullptr, join_rec = nullptr, join_gemm = nullptr;
    bool ok = false;
    PipeStreams()
    {
        ok = cudaStreamCreateWithFlags(&rec, cudaStreamNonBlocking) == cudaSuccess &&
             cudaStreamCreateWithFlags(&gemm, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&join_rec, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&join_gemm, cudaEventDisableTiming) == cudaSuccess;
        if (!ok) cudaGetLastError();
    }
};

int lstm_h16(int F, int B, int L, int T, const float *Wt, const float *bias, const float *x, float *y, float *c_last, bool pipeline)
{
    const size_t TB = (size_t)T * B;
    const size_t wsz = (size_t)4 * F * F;
    __half *R16 = reinterpret_cast<__half *>(pool_alloc((size_t)L * wsz * 2));             // packed recurrent weights, all layers
    float *Wp = reinterpret_cast<float *>(pool_alloc(wsz * 4));                            // packed input weights of layer 0 (fp32: x is fp32)
    __half *Wp16 = reinterpret_cast<__half *>(pool_alloc((size_t)L * wsz * 2));            // packed input weights of layers >= 1 (fp16)
    float *bp = reinterpret_cast<float *>(pool_alloc((size_t)L * 4 * F * 4));
    float *xg = reinterpret_cast<float *>(pool_alloc((size_t)2 * TB * 4 * F * 4));         // input contributions, by layer parity
    __half *h16 = reinterpret_cast<__half *>(pool_alloc((size_t)2 * TB * F * 2));          // h fp16, by layer parity
    // [L][256] step barrier words (one counter, 64 per-CTA flags or 8 counters 128 B apart), [L] xg_ready, [2L] weight
    // max-abs bit patterns, scratch word; then [2L] scale, [2L] 1/scale
    const size_t n_words = (size_t)256 * L + L + 2 * L + 1;
    int sync_mode = 0;
    if (const char *e = getenv("TB200_LSTM_SYNC")) sync_mode = atoi(e) >= 1 && atoi(e) <= 3 ? atoi(e) : 0;
    // R slice in tensor memory (H16Params::r16): K slice = F/2 <= 512 fp16 = 256 columns next to the 64 accumulator columns
    bool r_in_tmem = true;
    if (const char *e = getenv("TB200_LSTM_RTMEM")) r_in_tmem = r_in_tmem && atoi(e) != 0;
    uint32_t *flags = reinterpret_cast<uint32_t *>(pool_alloc((n_words + 4 * (size_t)L) * 4));
    if (!R16 || !Wp || !Wp16 || !bp || !xg || !h16 || !flags) return 1;
    uint32_t *step_flags = flags, *xg_ready = flags + 256 * L, *maxbits = xg_ready + L, *scratch = maxbits + 2 * L;
    float *wscale = reinterpret_cast<float *>(flags + n_words), *winv = wscale + 2 * L;
    TB200_CHECK(cudaMemsetAsync(flags, 0, n_words * 4, stream()));
    const int nctas = F / 16;                                    // <= 64 (F <= 1024)
    int chunk = 10;
    if (const char *e = getenv("TB200_LSTM_CHUNK")) chunk = atoi(e) > 0 ? atoi(e) : chunk;
    int dbg = 0;
    if (const char *e = getenv("TB200_LSTM_DBG")) dbg = atoi(e);

    {
        // power-of-two block scales of the 2L weight matrices (no host round trip: the pack kernels and the
        // epilogues read them from device memory)
        const unsigned bx = (unsigned)((wsz / 4 + 255) / 256 < (size_t)sm_count() * 2 ? (wsz / 4 + 255) / 256 : (size_t)sm_count() * 2);
        tiramisu_sm100_lstm_absmax<<<dim3(bx, (unsigned)(2 * L)), 256, 0, stream()>>>(Wt, maxbits, wsz / 4, 2 * L);
        if (post_launch("tiramisu_sm100_lstm_absmax")) return 1;
        tiramisu_sm100_lstm_scales<<<(unsigned)((2 * L + 63) / 64), 64, 0, stream()>>>(maxbits, wscale, winv, 2 * L);
        if (post_launch("tiramisu_sm100_lstm_scales")) return 1;
    }
    for (int l = 0; l < L; l++) {
        const size_t pk = (size_t)2 * 4 * F * (F / 4);
        const size_t want = (pk + 255) / 256, cap = (size_t)sm_count() * 16;
        tiramisu_sm100_lstm_pack16<<<(unsigned)(want < cap ? want : cap), 256, 0, stream()>>>(
            Wt + (size_t)l * 2 * wsz, bias + (size_t)l * 4 * F, R16 + (size_t)l * wsz, Wp,
            l == 0 ? nullptr : Wp16 + (size_t)l * wsz, bp + (size_t)l * 4 * F, wscale + 2 * l, F);
        if (post_launch("tiramisu_sm100_lstm_pack16")) return 1;
    }

    static thread_local PipeStreams *ps = nullptr;
    if (pipeline && !ps) ps = new PipeStreams();
    if (pipeline && !ps->ok) pipeline = false;

    auto out_of = [&](int l) -> float * { return l == L - 1 ? y : nullptr; };      // upper layers read the fp16 copy
    auto xg_of = [&](int l) -> float * { return xg + (size_t)(l & 1) * TB * 4 * F; };

    // recurrence of layer l on stream st
    auto launch_rec = [&](int l, cudaStream_t st, const uint32_t *ready, unsigned long long *trace) -> int {
        __half *hh = h16 + (size_t)(l & 1) * TB * F;
        CUtensorMap map_h, map_r;
        {
            uint64_t dims[3] = {(uint64_t)F, (uint64_t)B, (uint64_t)T};
            uint64_t str[2] = {(uint64_t)F * 2, (uint64_t)B * F * 2};
            uint32_t box[3] = {64, 64, 1};
            if (make_tensor_map_f16(&map_h, hh, 3, dims, str, box)) return 1;
        }
        {
            uint64_t dims[2] = {(uint64_t)F, 4 * (uint64_t)F};
            uint64_t str[1] = {(uint64_t)F * 2};
            uint32_t box[2] = {64, 128};
            if (make_tensor_map_f16(&map_r, R16 + (size_t)l * wsz, 2, dims, str, box)) return 1;
        }
        H16Params hp;
        hp.B = B; hp.F = F; hp.T = T;
        hp.xg = xg_of(l); hp.bias = bp + (size_t)l * 4 * F; hp.h = out_of(l); hp.h16 = hh; hp.c_out = l == L - 1 ? c_last : nullptr;
        hp.step_flags = step_flags + 256 * l; hp.sync_mode = sync_mode; hp.rscale = winv + 2 * l; hp.xg_ready = ready; hp.dbg = dbg; hp.trace = trace;
        hp.r16 = r_in_tmem ? R16 + (size_t)l * wsz : nullptr;
        cudaLaunchAttribute attr[1];
        cudaLaunchConfig_t cfg = h16_config(F, attr, st);
        TB200_CHECK(cudaLaunchKernelEx(&cfg, tiramisu_sm100_lstm_layer_h16, map_h, map_r, hp));
        return post_launch("tiramisu_sm100_lstm_layer_h16");
    };
    // Xg rows [t0*B, t1*B) of layer l = below * Wp(l)^T on the calling thread's current stream()
    auto launch_xg = [&](int l, int t0, int t1) -> int {
        float *dst = xg_of(l) + (size_t)t0 * B * 4 * F;
        if (l == 0)
            return gemm_f32_tc(x + (size_t)t0 * B * F, Wp, dst, (uint64_t)(t1 - t0) * B, 4 * (uint64_t)F, F, 1.0f, 0.0f, F, F,
                               4 * (uint64_t)F, true);
        const __half *below16 = h16 + (size_t)((l - 1) & 1) * TB * F + (size_t)t0 * B * F;
        return gemm_f16_tc(below16, Wp16 + (size_t)l * wsz, dst, (uint64_t)(t1 - t0) * B, 4 * (uint64_t)F, F, 1.0f, 0.0f, F, F,
                           4 * (uint64_t)F, winv + 2 * l + 1);
    };

    const char *trace_path = getenv("TB200_LSTM_TRACE");                 // bring-up: step timeline of layer 0
    const size_t trace_n = (size_t)nctas * T * 8;
    unsigned long long *trace = nullptr;
    if (trace_path) {
        trace = reinterpret_cast<unsigned long long *>(pool_alloc(trace_n * 8));
        if (!trace) return 1;
        TB200_CHECK(cudaMemsetAsync(trace, 0, trace_n * 8, stream()));
    }

    cudaStream_t main_st = stream();
    if (pipeline) {
        // CUDA loads kernels lazily, and loading one may synchronise the context: every kernel the
        // overlapped phase launches must already be loaded before a recurrence kernel starts spinning on
        // their results (the first pipelined run of r01 deadlocked exactly there and hit the wait
        // traps).  One tiny launch of each, once per process, on the main stream.
        static std::atomic<bool> warmed{false};
        if (!warmed.load()) {
            tiramisu_sm100_wait_counter<<<1, 64, 0, main_st>>>(scratch, 1, 0u, 0);
            if (post_launch("tiramisu_sm100_wait_counter")) return 1;
            tiramisu_sm100_publish<<<1, 1, 0, main_st>>>(scratch, 0u);
            if (post_launch("tiramisu_sm100_publish")) return 1;
            if (gemm_f16_tc(h16, Wp16, xg, (uint64_t)B, 4 * (uint64_t)F, F, 1.0f, 0.0f, F, F, 4 * (uint64_t)F, nullptr)) return 1;
            TB200_CHECK(cudaStreamSynchronize(main_st));
            warmed.store(true);
        }
    }
    for (int l = 0; l < L;) {
        // input contribution of layer l for all T steps: one full-chip GEMM
        if (launch_xg(l, 0, T)) return 1;
        const bool pair = pipeline && l + 1 < L;
        if (!pair) {
            if (launch_rec(l, main_st, nullptr, l == 0 ? trace : nullptr)) return 1;
            l += 1;
            continue;
        }
        // layers l and l+1 together: l on the main stream; the chunk GEMMs of Xg(l+1) follow l's step
        // counter on ps->gemm (they run on the SMs the two recurrences leave free); l+1 on ps->rec
        // consumes the chunks as they are published.
        TB200_CHECK(cudaEventRecord(ps->fork, main_st));
        TB200_CHECK(cudaStreamWaitEvent(ps->rec, ps->fork, 0));
        TB200_CHECK(cudaStreamWaitEvent(ps->gemm, ps->fork, 0));
        if (launch_rec(l, main_st, nullptr, l == 0 ? trace : nullptr)) return 1;
        if (launch_rec(l + 1, ps->rec, xg_ready + (l + 1), nullptr)) return 1;
        {
            ScopedStream on(ps->gemm);
            int rc = 0;
            for (int t0 = 0, n = 0; t0 < T && !rc; n++) {
                const int len = n == 0 ? (chunk < 4 ? chunk : 4) : (n == 1 ? (chunk < 6 ? chunk : 6) : chunk);
                const int t1 = t0 + len < T ? t0 + len : T;
                tiramisu_sm100_wait_counter<<<1, 64, 0, ps->gemm>>>(step_flags + 256 * l, nctas, (uint32_t)t1, sync_mode);
                rc = post_launch("tiramisu_sm100_wait_counter");
                if (!rc) rc = launch_xg(l + 1, t0, t1);
                if (!rc) {
                    tiramisu_sm100_publish<<<1, 1, 0, ps->gemm>>>(xg_ready + (l + 1), (uint32_t)t1);
                    rc = post_launch("tiramisu_sm100_publish");
                }
                t0 = t1;
            }
            if (rc) return 1;
        }
        TB200_CHECK(cudaEventRecord(ps->join_rec, ps->rec));
        TB200_CHECK(cudaEventRecord(ps->join_gemm, ps->gemm));
        TB200_CHECK(cudaStreamWaitEvent(main_st, ps->join_rec, 0));
        TB200_CHECK(cudaStreamWaitEvent(main_st, ps->join_gemm, 0));
        l += 2;
    }
    if (trace_path) {
        std::vector<unsigned long long> host(trace_n);
        TB200_CHECK(cudaMemcpyAsync(host.data(), trace, trace_n * 8, cudaMemcpyDeviceToHost, stream()));
        TB200_CHECK(cudaStreamSynchronize(stream()));
        if (FILE *f = fopen(trace_path, "wb")) { fwrite(host.data(), 8, trace_n, f); fclose(f); }
        pool_free(trace);
    }
    pool_free(flags); pool_free(h16); pool_free(xg); pool_free(bp); pool_free(Wp16); pool_free(Wp); pool_free(R16);
    return 0;
}

int lstm_tensor(int F, int B, int L, int T, const float *Wt, const float *bias, const float *x, float *y, float *c_last)
{
    // recurrence engine (g_lstm_kernel: 0 auto, 1 = one launch per cell, 2 = persistent N-split layer
    // kernel, 3 = persistent TF32 K-split layer kernel, 4 = fp16 K-split kernel with layer pairs
    // overlapped (TB200_LSTM_PIPE=0: one layer at a time)); auto prefers 4, then 3, 2, 1
    int which = g_lstm_kernel.load();
    if (const char *e = getenv("TB200_LSTM_KERNEL")) which = atoi(e);
    if (which == 0 || which == 4) {
        bool pair_ok = false;
        const bool ok = h16_ok(F, B, &pair_ok);
        if (which == 4 && !ok) return fail("tiramisu_b200_lstm", "shape not supported by the fp16 K-split layer kernel");
        if (ok) {
            bool pipe = pair_ok;
            if (const char *e = getenv("TB200_LSTM_PIPE")) pipe = pipe && atoi(e) != 0;
            t_last_engine = pipe ? "h16: block-scaled fp16 operands (11-bit significand = TF32's), R in tensor memory, K split over 2-CTA clusters, layer pairs overlapped"
                                 : "h16: block-scaled fp16 operands (11-bit significand = TF32's), R in tensor memory, K split over 2-CTA clusters, one layer at a time";
            return lstm_h16(F, B, L, T, Wt, bias, x, y, c_last, pipe);
        }
    }
    const size_t BF = (size_t)B * F, TB = (size_t)T * B;
    float *Wp = reinterpret_cast<float *>(pool_alloc((size_t)2 * 4 * F * F * 4));          // packed weights of one layer
    float *bp = reinterpret_cast<float *>(pool_alloc((size_t)4 * F * 4));
    float *xg = reinterpret_cast<float *>(pool_alloc(TB * 4 * F * 4));                     // input contributions, all steps
    float *hs = reinterpret_cast<float *>(pool_alloc((size_t)2 * TB * F * 4));             // h of layer l (ping-pong by layer)
    float *cs = reinterpret_cast<float *>(pool_alloc((size_t)2 * BF * 4));                 // c ping-pong by step
    float *zero = reinterpret_cast<float *>(pool_alloc(BF * 4));
    uint32_t *counters = reinterpret_cast<uint32_t *>(pool_alloc((size_t)L * 4));
    if (!Wp || !bp || !xg || !hs || !cs || !zero || !counters) return 1;
    TB200_CHECK(cudaMemsetAsync(zero, 0, BF * 4, stream()));
    TB200_CHECK(cudaMemsetAsync(counters, 0, (size_t)L * 4, stream()));
    const bool use_ksplit = (which == 0 || which == 3) && ksplit_ok(F, B);
    if (which == 3 && !use_ksplit) return fail("tiramisu_b200_lstm", "shape not supported by the K-split layer kernel");
    const bool use_layer_kernel = !use_ksplit && which != 1 && which != 3 && layer_kernel_ok(F, B);
    if (which == 2 && !use_layer_kernel) return fail("tiramisu_b200_lstm", "shape not supported by the persistent layer kernel");
    const int a_rows = B <= 64 ? 64 : 128;
    const int cl = use_layer_kernel ? pick_cluster(F, a_rows) : 1;
    t_last_engine = use_ksplit ? "ksplit: TF32 operands, K split over 4-CTA clusters (DSMEM reduction)"
                    : use_layer_kernel ? "layer: TF32 operands, gate columns split across CTAs, persistent per layer"
                                       : "step: TF32 operands, one launch per cell";
    static bool attr_set = false;
    if (!attr_set) {
        TB200_CHECK(cudaFuncSetAttribute(tiramisu_sm100_lstm_step_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         SMEM_BYTES));
        attr_set = true;
    }
    for (int l = 0; l < L; l++) {
        const float *below = (l == 0) ? x : hs + (size_t)((l - 1) & 1) * TB * F;           // [T*B][F]
        float *hl = (l == L - 1) ? y : hs + (size_t)(l & 1) * TB * F;                       // last layer writes y directly
        const size_t pk = (size_t)2 * 4 * F * (F / 4);
        const size_t want = (pk + 255) / 256, cap = (size_t)sm_count() * 16;
        tiramisu_sm100_lstm_pack<<<(unsigned)(want < cap ? want : cap), 256, 0, stream()>>>(
            Wt + (size_t)l * 2 * 4 * F * F, bias + (size_t)l * 4 * F, Wp, bp, F);
        if (post_launch("tiramisu_sm100_lstm_pack")) return 1;
        // 1. input contribution of all T steps: Xg[T*B][4F] = below * Wp(W)^T
        if (gemm_f32_tc(below, Wp + (size_t)4 * F * F, xg, TB, 4 * (uint64_t)F, F, 1.0f, 0.0f, F, F, 4 * (uint64_t)F,
                        true))
            return 1;
        // 2. the recurrence
        CUtensorMap map_r;
        {
            uint64_t dims[2] = {(uint64_t)F, 4 * (uint64_t)F};
            uint64_t str[1] = {(uint64_t)F * 4};
            uint32_t box[2] = {BK, BN};
            if (make_tensor_map_f32(&map_r, Wp, 2, dims, str, box)) return 1;
        }
        if (use_ksplit) {
            CUtensorMap map_h3, map_ra;
            {
                uint64_t dims[3] = {(uint64_t)F, (uint64_t)B, (uint64_t)T};
                uint64_t str[2] = {(uint64_t)F * 4, (uint64_t)B * F * 4};
                uint32_t box[3] = {BK, 64, 1};
                if (make_tensor_map_f32(&map_h3, hl, 3, dims, str, box)) return 1;
            }
            {
                uint64_t dims[2] = {(uint64_t)F, 4 * (uint64_t)F};
                uint64_t str[1] = {(uint64_t)F * 4};
                uint32_t box[2] = {BK, 128};
                if (make_tensor_map_f32(&map_ra, Wp, 2, dims, str, box)) return 1;
            }
            KSplitParams kp;
            kp.B = B; kp.F = F; kp.T = T; kp.xg = xg; kp.bias = bp; kp.h = hl; kp.counter = counters + l;
            kp.c_out = l == L - 1 ? c_last : nullptr;
            kp.dbg = 0;
            kp.trace = nullptr;
            if (const char *e = getenv("TB200_LSTM_DBG")) kp.dbg = atoi(e);
            const char *trace_path = (l == 0) ? getenv("TB200_LSTM_TRACE") : nullptr;    // bring-up: timeline of layer 0
            const size_t trace_n = (size_t)(F / 8) * T * 8;
            if (trace_path) {
                kp.trace = reinterpret_cast<unsigned long long *>(pool_alloc(trace_n * 8));
                if (!kp.trace) return 1;
                TB200_CHECK(cudaMemsetAsync(kp.trace, 0, trace_n * 8, stream()));
            }
            if (launch_ksplit(map_h3, map_ra, kp)) return 1;
            if (trace_path) {
                std::vector<unsigned long long> host(trace_n);
                TB200_CHECK(cudaMemcpyAsync(host.data(), kp.trace, trace_n * 8, cudaMemcpyDeviceToHost, stream()));
                TB200_CHECK(cudaStreamSynchronize(stream()));
                if (FILE *f = fopen(trace_path, "wb")) { fwrite(host.data(), 8, trace_n, f); fclose(f); }
                pool_free(kp.trace);
            }
            continue;
        }
        if (use_layer_kernel) {
            // one persistent launch for all T steps
            CUtensorMap map_h3;
            uint64_t dims[3] = {(uint64_t)F, (uint64_t)B, (uint64_t)T};
            uint64_t str[2] = {(uint64_t)F * 4, (uint64_t)B * F * 4};
            uint32_t box[3] = {BK, (uint32_t)(a_rows / cl), 1};
            if (make_tensor_map_f32(&map_h3, hl, 3, dims, str, box)) return 1;
            LayerParams lp;
            lp.B = B; lp.F = F; lp.T = T; lp.a_rows = a_rows;
            lp.xg = xg; lp.bias = bp; lp.h = hl; lp.counter = counters + l; lp.c_out = l == L - 1 ? c_last : nullptr;
            lp.dbg = 0;
            if (const char *e = getenv("TB200_LSTM_DBG")) lp.dbg = atoi(e);
            bool unused = false;
            if (layer_dispatch(cl, map_h3, map_r, lp, false, &unused)) return 1;
            continue;
        }
        for (int m = 0; m < T; m++) {
            const float *hprev = (m == 0) ? zero : hl + (size_t)(m - 1) * BF;
            CUtensorMap map_h;
            uint64_t dims[2] = {(uint64_t)F, (uint64_t)B};
            uint64_t str[1] = {(uint64_t)F * 4};
            uint32_t box[2] = {BK, BM};
            if (make_tensor_map_f32(&map_h, hprev, 2, dims, str, box)) return 1;
            StepParams p;
            p.B = B; p.F = F;
            p.xg = xg + (size_t)m * B * 4 * F;
            p.bias = bp;
            p.cprev = (m == 0) ? zero : cs + (size_t)((m - 1) & 1) * BF;
            p.c = cs + (size_t)(m & 1) * BF;
            p.h = hl + (size_t)m * BF;
            p.y = nullptr;
            tiramisu_sm100_lstm_step_tf32<<<4 * F / BN, kThreads, SMEM_BYTES, stream()>>>(map_h, map_r, p);
            if (post_launch("tiramisu_sm100_lstm_step_tf32")) return 1;
        }
        if (c_last && l == L - 1)              // per-cell engine: the cell state lives in the ping-pong buffers
            TB200_CHECK(cudaMemcpyAsync(c_last, cs + (size_t)((T - 1) & 1) * BF, BF * 4, cudaMemcpyDeviceToDevice, stream()));
    }
    pool_free(counters);
    pool_free(Wp); pool_free(bp); pool_free(xg); pool_free(hs); pool_free(cs); pool_free(zero);
    return 0;
}

}  // namespace

TB200_API int tiramisu_b200_set_lstm_mode(int mode)
{
    if (mode < 0 || mode > 2) mode = 0;
    return g_lstm_mode.exchange(mode);
}

TB200_API int tiramisu_b200_set_lstm_kernel(int kernel)
{
    if (kernel < 0 || kernel > 4) kernel = 0;
    return g_lstm_kernel.exchange(kernel);
}

static int lstm_dispatch(const float *Weights, const float *biases, const float *x, float *y, float *c_last, int F, int B, int L, int T,
                         const char *who)
{
    int mode = g_lstm_mode.load();
    if (const char *e = getenv("TB200_LSTM_MODE")) mode = atoi(e);     // for harnesses that cannot call the setter
    const bool tc_ok = F % 32 == 0 && B <= 128 && aligned16(Weights) && aligned16(biases) && aligned16(x) && aligned16(y) &&
                       (!c_last || aligned16(c_last));
    if (mode == 2 && !tc_ok) return fail(who, "shape not supported by the tcgen05 LSTM path");
    if (mode != 1 && tc_ok) return lstm_tensor(F, B, L, T, Weights, biases, x, y, c_last);
    t_last_engine = "exact: fp32 SIMT in the reference's accumulation order";
    return lstm_exact(F, B, L, T, Weights, biases, x, y, c_last);
}

TB200_API int tiramisu_b200_lstm(const float *Weights, const float *biases, const float *x, float *y, int32_t F,
                                 int32_t B, int32_t L, int32_t T)
{
    if (!Weights || !biases || !x || !y || F < 1 || B < 1 || L < 1 || T < 1) return fail(__func__, "bad argument");
    return lstm_dispatch(Weights, biases, x, y, nullptr, F, B, L, T, __func__);
}

// ---- the other two `lstm` ABIs of the reference: same cell, other parameter layouts ---------------------
// layout 1 — LSTM/gpu_lib/generator.cpp:34-37,66-88: Weights [L][2][4F][F] with index 0 = W (input), 1 = R
//            (recurrent), gate row blocks i, f, z, o; biases [L][4F] in the same gate order
// layout 2 — LSTM/cpu_lib/generator.cpp:27-30,50-76: weights [L][2][F][4F] (TRANSPOSED: input feature major), index
//            0 = W, 1 = R, gate column blocks i, f, z, o
// Both are converted once per call into the layout of LSTM/cpu (layout 0: [L][2][4F][F], 0 = R, 1 = W, blocks
// i, z, o, f) by one pass over the weights and then run on the same engines.
namespace {
__device__ __forceinline__ int src_gate_block(int g_ours) { return g_ours == 0 ? 0 : (g_ours == 1 ? 2 : (g_ours == 2 ? 3 : 1)); }   // i, z, o, f <- i, f, z, o

__global__ void __launch_bounds__(256)
tiramisu_sm100_lstm_relayout_rows(const float *__restrict__ src, float *__restrict__ dst, int F, int L)
{
    // layout 1 -> 0: dst[l][m][g*F + u][:] = src[l][1 - m][src_gate_block(g)*F + u][:], float4 per thread
    const size_t total4 = (size_t)L * 2 * 4 * F * (F / 4);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
        const int j4 = (int)(i % (F / 4));
        size_t t = i / (F / 4);
        const int row = (int)(t % (4 * F)); t /= 4 * F;
        const int m = (int)(t % 2), l = (int)(t / 2);
        const int g = row / F, u = row % F;
        const float4 *s4 = reinterpret_cast<const float4 *>(src + (((size_t)l * 2 + (1 - m)) * 4 * F + (size_t)src_gate_block(g) * F + u) * F);
        reinterpret_cast<float4 *>(dst + (((size_t)l * 2 + m) * 4 * F + row) * F)[j4] = __ldg(s4 + j4);
    }
}

__global__ void __launch_bounds__(256)
tiramisu_sm100_lstm_relayout_transposed(const float *__restrict__ src, float *__restrict__ dst, int F)
{
    // layout 2 -> 0 for one (layer, matrix) pair per blockIdx.z: dst[g*F + u][j] = src[j][src_gate_block(g)*F + u];
    // 32 x 32 tiles through shared memory so that both sides are coalesced
    __shared__ float tile[32][33];
    const int lm = blockIdx.z, l = lm / 2, m = lm % 2;
    const float *S = src + ((size_t)l * 2 + (1 - m)) * 4 * F * F;       // [F (j)][4F (their rows)]
    float *D = dst + ((size_t)l * 2 + m) * 4 * F * F;                    // [4F (our rows)][F (j)]
    const int r0 = blockIdx.y * 32, j0 = blockIdx.x * 32;                // tile of OUR rows x j
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;              // 32 x 8 threads
    for (int k = ty; k < 32; k += 8) {
        const int row = r0 + tx, j = j0 + k;                             // read src[j][their(row)]: consecutive tx -> consecutive columns
        if (row < 4 * F && j < F) tile[k][tx] = S[(size_t)j * 4 * F + (size_t)src_gate_block(row / F) * F + row % F];
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int row = r0 + k, j = j0 + tx;
        if (row < 4 * F && j < F) D[(size_t)row * F + j] = tile[tx][k];
    }
}

__global__ void tiramisu_sm100_lstm_relayout_bias(const float *__restrict__ src, float *__restrict__ dst, int F, int L)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L * 4 * F) return;
    const int l = i / (4 * F), row = i % (4 * F);
    dst[i] = src[(size_t)l * 4 * F + src_gate_block(row / F) * F + row % F];
}
}  // namespace

TB200_API int tiramisu_b200_lstm_ex(const float *Weights, const float *biases, const float *x, float *y, float *c_last, int32_t F,
                                    int32_t B, int32_t L, int32_t T, int32_t layout)
{
    if (!Weights || !biases || !x || !y || F < 1 || B < 1 || L < 1 || T < 1 || layout < 0 || layout > 2) return fail(__func__, "bad argument");
    if (layout == 0) return lstm_dispatch(Weights, biases, x, y, c_last, F, B, L, T, __func__);
    const size_t wn = (size_t)L * 2 * 4 * F * F, bn = (size_t)L * 4 * F;
    float *W0 = reinterpret_cast<float *>(pool_alloc(wn * 4)), *b0 = reinterpret_cast<float *>(pool_alloc(bn * 4));
    if (!W0 || !b0) return 1;
    if (layout == 1 && F % 4 == 0 && aligned16(Weights)) {
        const size_t want = (wn / 4 + 255) / 256, cap = (size_t)sm_count() * 16;
        tiramisu_sm100_lstm_relayout_rows<<<(unsigned)(want < cap ? want : cap), 256, 0, stream()>>>(Weights, W0, F, L);
        if (post_launch("tiramisu_sm100_lstm_relayout_rows")) return 1;
    } else if (layout == 1) {
        // rows that cannot be moved as float4: per (layer, matrix, gate block) strided copies
        for (int l = 0; l < L; l++)
            for (int m = 0; m < 2; m++)
                for (int g = 0; g < 4; g++) {
                    const int sg = g == 0 ? 0 : (g == 1 ? 2 : (g == 2 ? 3 : 1));
                    TB200_CHECK(cudaMemcpyAsync(W0 + (((size_t)l * 2 + m) * 4 + g) * F * F, Weights + (((size_t)l * 2 + (1 - m)) * 4 + sg) * F * F,
                                                (size_t)F * F * 4, cudaMemcpyDeviceToDevice, stream()));
                }
    } else {
        dim3 grid((unsigned)((F + 31) / 32), (unsigned)((4 * F + 31) / 32), (unsigned)(2 * L));
        tiramisu_sm100_lstm_relayout_transposed<<<grid, 256, 0, stream()>>>(Weights, W0, F);
        if (post_launch("tiramisu_sm100_lstm_relayout_transposed")) return 1;
    }
    tiramisu_sm100_lstm_relayout_bias<<<(unsigned)((bn + 255) / 256), 256, 0, stream()>>>(biases, b0, F, L);
    if (post_launch("tiramisu_sm100_lstm_relayout_bias")) return 1;
    const int rc = lstm_dispatch(W0, b0, x, y, c_last, F, B, L, T, __func__);
    pool_free(W0); pool_free(b0);
    return rc;
}

TB200_API const char *tiramisu_b200_lstm_last_engine(void) { return t_last_engine; }
