// gemm_tc.cu — fp32 GEMM on the 5th-generation tensor cores (tcgen05, kind::tf32, fp32
// accumulators in TMEM), the sm_100a lowering of the contraction pattern
//     c_acc(i,j,k) = c_acc(i,j,0) + A(i,k)*B(k,j);  C = acc*alpha + C*beta
// (reference: benchmarks/linear_algebra/blas/level3/sgemm/gpu/generator1.cpp:68-72, which the
// cuda_ast emitter turns into a 16x16-thread SIMT kernel with hand-staged __shared__ tiles, and
// tiramisu_cublas_sgemm, src/tiramisu_cuda_wrappers.cpp:74-115).
//
// Structure (one 128 x BN output tile per CTA, 192 threads):
//   warp 0   : TMA producer — cp.async.bulk.tensor loads of the A (128x32 fp32) and B (32xBN)
//              k-slabs into a STAGES-deep ring of 128B-swizzled smem tiles, mbarrier expect_tx.
//   warp 1   : MMA issuer — one thread issues 4 x tcgen05.mma (M=128,N=BN,K=8) per slab, frees
//              each slab with tcgen05.commit, signals the epilogue with a final commit.
//   warps 2-5: epilogue — tcgen05.ld the accumulator rows, apply alpha/beta, 128-bit stores.
// fp32 operands are consumed in place as TF32 (the tensor core ignores the low 13 mantissa
// bits), so there is no conversion pass over HBM.  A is K-major; B is either K-major ([N][K],
// the transposeB / LSTM-weights case) or MN-major ([K][N], the plain row-major sgemm case), the
// latter fetched as BN/32 boxes of 32(k) x 32(n) so that each box is one MN-major swizzle slab.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tiramisu_b200.h"

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

using namespace tb200;

namespace tb200 {

// ---- cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link) ----------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// Descriptors are pure functions of (base, shape, pitches, box, swizzle, element type): a small per-thread cache
// keyed by exactly those fields removes cuTensorMapEncodeTiled from the launch path of repeated calls — at 1024^3
// the two encodes per GEMM were a visible part of a ~12 us host-side launch period (r02 run 2), and generated
// host loops (the lowered LSTM: 1600 GEMM calls on the same few operands) re-encoded the same maps every iteration.
struct TmapKey {
    const void *base;
    uint64_t dims[5], strides[4];
    uint32_t box[5];
    int rank, swizzle, dtype;
    bool operator==(const TmapKey &o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapSlot { TmapKey key; CUtensorMap map; bool valid = false; };
constexpr int kTmapSlots = 256;

static int make_tensor_map_cached(CUtensorMap *map, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                                  const uint32_t *box, int swizzle, int dtype /* 0 fp32, 1 fp16 */, const char *who)
{
    static thread_local TmapSlot *slots = new TmapSlot[kTmapSlots]();
    TmapKey k;
    memset(&k, 0, sizeof(k));
    k.base = base; k.rank = rank; k.swizzle = swizzle; k.dtype = dtype;
    uint64_t h = reinterpret_cast<uintptr_t>(base) * 0x9E3779B97F4A7C15ull + (uint64_t)rank * 31 + (uint64_t)swizzle * 7 + (uint64_t)dtype;
    for (int i = 0; i < rank; i++) { k.dims[i] = dims[i]; k.box[i] = box[i]; h = (h ^ dims[i]) * 0x100000001B3ull; h = (h ^ box[i]) * 0x100000001B3ull; }
    for (int i = 0; i + 1 < rank; i++) { k.strides[i] = strides_bytes[i]; h = (h ^ strides_bytes[i]) * 0x100000001B3ull; }
    TmapSlot &s = slots[(h >> 17) % kTmapSlots];
    if (s.valid && s.key == k) { *map = s.map; return 0; }
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(who, "cuTensorMapEncodeTiled unavailable");
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], estr[5];
    for (int i = 0; i < rank; i++) { gdim[i] = dims[i]; bx[i] = box[i]; estr[i] = 1; }
    for (int i = 0; i + 1 < rank; i++) gstr[i] = strides_bytes[i];
    CUresult r = fn(map, dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                    const_cast<void *>(base), gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle == kTmaSwizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B
                    : swizzle == kTmaSwizzle128Atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                                      : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(who, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    s.key = k; s.map = *map; s.valid = true;
    return 0;
}

int make_tensor_map_f32(CUtensorMap *map, const void *base, int rank, const uint64_t *dims,
                        const uint64_t *strides_bytes, const uint32_t *box, int swizzle)
{
    return make_tensor_map_cached(map, base, rank, dims, strides_bytes, box, swizzle, 0, "make_tensor_map_f32");
}

int make_tensor_map_f16(CUtensorMap *map, const void *base, int rank, const uint64_t *dims,
                        const uint64_t *strides_bytes, const uint32_t *box)
{
    return make_tensor_map_cached(map, base, rank, dims, strides_bytes, box, kTmaSwizzle128, 1, "make_tensor_map_f16");
}

// Per-thread overrides used by callers that place GEMMs next to other resident kernels (the LSTM
// layer pipeline runs its chunk GEMMs on the SMs the recurrence kernels leave free).
static thread_local int t_hint_kernel = 0, t_hint_bn = 0;
void gemm_tc_hint(int kernel, int bn) { t_hint_kernel = kernel; t_hint_bn = bn; }

// Arrival gate for the NEXT gemm_f32_tc call of this thread (consumed by it): device address of a uint32 that some
// other stream raises to >= value once the B operand is in place.  Persistent kernels poll it in the TMA producer
// warp (the launch, prologue and A loads overlap the tail of the transfer); the other kernels get a stream-side
// cuStreamWaitValue32 in front of the launch.
static thread_local const uint32_t *t_gate = nullptr;
static thread_local uint32_t t_gate_value = 0, t_gate_last = 0;
static thread_local int t_gate_cols = 0;
// cols == 0: the whole operand has arrived once *flag >= value.  cols > 0: columns [j cols, (j+1) cols) of B have arrived
// once *flag >= value + j; `last` = the token of the final panel (what a kernel without an in-kernel gate waits for).
void gemm_tc_gate(const uint32_t *flag, uint32_t value, int cols, uint32_t last)
{
    t_gate = flag; t_gate_value = value; t_gate_cols = cols; t_gate_last = cols ? last : value;
}
// Layout of the (MN-major) B operand of the NEXT gemm_f32_tc call of this thread: panel-major, column panels of `cols`
// columns, consecutive panels `panel_stride` elements apart, ldB == cols.  Persistent kernels only (the caller asks
// gemm_tc_kernel_choice first).
static thread_local int t_b_panel_cols = 0;
static thread_local uint64_t t_b_panel_stride = 0;
void gemm_tc_b_panels(int cols, uint64_t panel_stride) { t_b_panel_cols = cols; t_b_panel_stride = panel_stride; }
int stream_wait_value32(cudaStream_t st, const uint32_t *flag, uint32_t value);       // comm.cu

}  // namespace tb200

namespace {

constexpr int BM = 128;
constexpr int BK = 32;                 // 32 fp32 = one 128-byte swizzle row
constexpr int UMMA_K = 8;              // kind::tf32
constexpr int kThreads = 192;
constexpr int A_TILE_BYTES = BM * BK * 4;

template <int BN> struct Cfg {
    static constexpr int B_TILE_BYTES = BN * BK * 4;
    static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmParams {
    int M, N, K;
    float alpha, beta;
    float *C;
    int64_t ldC;
    int num_m, num_n, group_m;
    uint32_t b_lbo, b_sbo;              // MN-major B descriptor offsets (bytes)
    const float *alpha_dev;             // optional device-resident factor multiplied into alpha (LSTM: the
                                        // power-of-two de-scaling of block-scaled fp16 weights), else null
    int prefetch_kb;                    // L2 prefetch distance in k-slabs beyond the shared-memory ring (0 = off)
    const uint32_t *gate = nullptr;     // optional arrival word of the B operand (row-sharded sgemm: B is being copied in by
    uint32_t gate_value = 0;            // DMA engines while this kernel runs).  The TMA producer of a persistent kernel polls
    int gate_cols = 0;                  // *gate >= gate_value + n0 / gate_cols before the B loads of a tile whose first column
                                        // is n0 (column panels arrive in order); gate_cols == 0: one wait before the first load
    int b_panel_cols = 0;               // > 0: MN-major B is stored panel-major, [N / cols][K][cols] (each column panel a
                                        // contiguous row-major matrix — what the DMA broadcast delivers); map_b is then 3-D
    // stream-K (persistent kernels): the first sk_dp_waves * clusters tiles are whole-tile work, one per cluster and wave;
    // the k-slabs of the remaining tiles (from tile sk_first_tile on) are cut into equal contiguous runs of sk_units per
    // cluster.  A run that starts inside a tile is a CONTRIBUTION (raw accumulator -> sk_ws slot of the cluster, per-warp
    // flag); the cluster whose run holds the tile's first slab OWNS it and adds the contributions in its epilogue.
    float *sk_ws = nullptr;             // [clusters][TILE_M][256] fp32, null = plain persistent tiles
    uint32_t *sk_flags = nullptr;       // [clusters][2 * kEpiWarps]: 1 = contribution of that warp is in place (owner resets to 0)
    int sk_dp_waves = 0, sk_first_tile = 0, sk_units = 0, sk_total = 0;
    // sk_split = s >= 2: k-ALIGNED split instead of contiguous runs — remainder tile r belongs to clusters [r s, r s + s), member
    // i takes slabs [i sk_units, (i + 1) sk_units) (member 0 owns the tile).  The clusters of one k-phase still sweep K in
    // lockstep, so concurrent tiles keep sharing their A / B slabs in L2; contiguous runs put every cluster at a different k
    // offset and the remainder phase became HBM-bound (r02 run 21: 4096^3 2.6x slower per slab, 662 vs 693 TFLOP/s overall).
    int sk_split = 0;
};

// F16: operands are fp16 (kind::f16, K = 16 per instruction, 64 elements per 128-byte swizzle row):
// same tile bytes, ring and descriptors; only the K bookkeeping and the MMA kind differ.  Used by the
// LSTM layer pipeline, whose upper layers take the fp16 copy of h as the A operand.
template <int BN, bool B_MN, bool F16 = false>
__global__ void __launch_bounds__(kThreads, 1)
tiramisu_sm100_gemm_tf32(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         const GemmParams p)
{
    static_assert(!(F16 && B_MN), "fp16 operands are K-major only");
    constexpr int BKE = F16 ? 64 : BK;                 // elements per k-slab (128 bytes)
    using C_ = Cfg<BN>;
    constexpr int STAGES = C_::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *smem_a = smem;
    uint8_t *smem_b = smem + STAGES * A_TILE_BYTES;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * C_::STAGE_BYTES);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tmem_full_bar = empty_bar + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // grouped rasterisation: GROUP_M m-tiles share the same B panels while they are L2-resident
    const int t = blockIdx.x;
    const int tiles_per_group = p.group_m * p.num_n;
    const int group = t / tiles_per_group;
    const int first_m = group * p.group_m;
    const int gsz = min(p.num_m - first_m, p.group_m);
    const int tile_m = first_m + (t % tiles_per_group) % gsz;
    const int tile_n = (t % tiles_per_group) / gsz;
    const int m0 = tile_m * BM, n0 = tile_n * BN;
    const int num_kb = (p.K + BKE - 1) / BKE;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_b);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; s++) {
                ptx::mbar_init(&full_bar[s], 1);
                ptx::mbar_init(&empty_bar[s], 1);
            }
            ptx::mbar_init(tmem_full_bar, 1);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, BN);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: whole warp walks the ring (uniform control flow), one elected lane issues =====
        for (int kb = 0; kb < num_kb; kb++) {
            const int s = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            ptx::mbar_wait(&empty_bar[s], ph ^ 1);
            if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(&full_bar[s], C_::STAGE_BYTES);
                const int k0 = kb * BKE;
                ptx::tma_load_2d(smem_a + s * A_TILE_BYTES, &map_a, &full_bar[s], k0, m0);
                uint8_t *bdst = smem_b + s * C_::B_TILE_BYTES;
                if (B_MN) {
#pragma unroll
                    for (int j = 0; j < BN / 32; j++)
                        ptx::tma_load_2d(bdst + j * (BK * 128), &map_b, &full_bar[s], n0 + 32 * j, k0);
                } else {
                    ptx::tma_load_2d(bdst, &map_b, &full_bar[s], k0, n0);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the ring (uniform control flow), one elected lane issues =====
        constexpr uint32_t idesc = F16 ? ptx::make_idesc_f16(BM, BN, false, false) : ptx::make_idesc_tf32(BM, BN, false, B_MN);
        // K-major, 128B swizzle: rows are 128 B apart, 8-row groups 1024 B apart; a K step of 8 fp32
        // moves the start address 32 B inside the swizzled row.
        // MN-major tf32 (128B swizzle, 32B atoms): rows are k (128 B = 32 columns each), 4-row groups
        // 512 B apart (stride offset), 32-column slabs BK*128 B apart (leading offset); a K step of
        // 8 rows moves the start address 1024 B.
        const uint64_t da0 = ptx::make_smem_desc(ptx::smem_u32(smem_a), 16, 1024);
        const uint64_t db0 = B_MN ? ptx::make_smem_desc(ptx::smem_u32(smem_b), p.b_lbo, p.b_sbo, ptx::kLayoutSw128Base32)
                                  : ptx::make_smem_desc(ptx::smem_u32(smem_b), 16, 1024);
        constexpr uint32_t kStepA = UMMA_K * 4, kStepB = B_MN ? UMMA_K * 128 : UMMA_K * 4;
        for (int kb = 0; kb < num_kb; kb++) {
            const int s = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            ptx::mbar_wait(&full_bar[s], ph);
            ptx::tc_fence_after_sync();
            if (ptx::elect_one()) {
                const uint64_t da = ptx::desc_advance(da0, (uint32_t)s * A_TILE_BYTES);
                const uint64_t db = ptx::desc_advance(db0, (uint32_t)s * C_::B_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; k++) {
                    if (F16)
                        ptx::mma_f16_ss(tmem_base, ptx::desc_advance(da, k * kStepA), ptx::desc_advance(db, k * kStepB),
                                        idesc, (kb | k) != 0);
                    else
                        ptx::mma_tf32_ss(tmem_base, ptx::desc_advance(da, k * kStepA), ptx::desc_advance(db, k * kStepB),
                                         idesc, (kb | k) != 0);
                }
                ptx::mma_commit(&empty_bar[s]);        // slab free once these MMAs retire
                if (kb == num_kb - 1) ptx::mma_commit(tmem_full_bar);     // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue: TMEM -> registers -> global =====
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        const int row = m0 + q * 32 + lane;
        const bool row_ok = row < p.M;
        float *crow = p.C + (size_t)row * p.ldC;
        const bool use_c = p.beta != 0.0f;             // cuBLAS semantics: C not read when beta == 0
        const float alpha = p.alpha_dev ? __fmul_rn(p.alpha, __ldg(p.alpha_dev)) : p.alpha;
        // The beta*C operand does not depend on the MMAs: fetch the first 32-column chunk of this
        // thread's C row while the main loop runs, and each further chunk one iteration ahead (at
        // 1024^3 the whole kernel lasts a few microseconds and this latency was exposed, r01 ncu).
        auto load_c = [&](int c0, float4 (&cv)[8]) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int col = n0 + c0 + 4 * j;
                cv[j] = (use_c && row_ok && col + 3 < p.N) ? *reinterpret_cast<const float4 *>(crow + col)
                                                           : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        float4 cnext[8];
        load_c(0, cnext);
        ptx::mbar_wait(tmem_full_bar, 0);
        ptx::tc_fence_after_sync();
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            float4 ccur[8];
#pragma unroll
            for (int j = 0; j < 8; j++) ccur[j] = cnext[j];
            if (c0 + 32 < BN) load_c(c0 + 32, cnext);
            uint32_t r[32];
            __syncwarp();                              // tcgen05.ld is .sync.aligned: whole warp, converged
            ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            ptx::tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int col = n0 + c0 + j;
                    if (col + 3 < p.N) {
                        float4 o;
                        o.x = __fmul_rn(__uint_as_float(r[j + 0]), alpha);
                        o.y = __fmul_rn(__uint_as_float(r[j + 1]), alpha);
                        o.z = __fmul_rn(__uint_as_float(r[j + 2]), alpha);
                        o.w = __fmul_rn(__uint_as_float(r[j + 3]), alpha);
                        if (use_c) {
                            const float4 c = ccur[j >> 2];
                            o.x = __fadd_rn(o.x, __fmul_rn(c.x, p.beta));
                            o.y = __fadd_rn(o.y, __fmul_rn(c.y, p.beta));
                            o.z = __fadd_rn(o.z, __fmul_rn(c.z, p.beta));
                            o.w = __fadd_rn(o.w, __fmul_rn(c.w, p.beta));
                        }
                        *reinterpret_cast<float4 *>(crow + col) = o;
                    } else {
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            if (col + u < p.N) {
                                float o = __fmul_rn(__uint_as_float(r[j + u]), alpha);
                                if (use_c) o = __fadd_rn(o, __fmul_rn(crow[col + u], p.beta));
                                crow[col + u] = o;
                            }
                        }
                    }
                }
            }
        }
    }

    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem_base, BN);
    }
}


// ================================================================================================
// Split-K variant for SMALL problems (1024^3: BASELINE configs[0]).  With one 128 x 64 tile per CTA a 1024^3
// product moves 4*M*N*K*(1/BM + 1/BN) = 100 MB from L2 into shared memory — 8 us at the chip's ~6300 B/clk L2
// throughput, against a 2.6 us roofline — because tiles that small re-read A and B 16 and 8 times.  Here a
// CLUSTER OF TWO CTAs shares one 128 x 128 output tile and each takes half of K: 64 tiles x 2 = 128 CTAs keep the
// machine as full as before while the L2 -> SMEM traffic drops to 64 MB and every MMA is N = 128 wide (one
// shared-memory operand byte per tensor-pipe cycle instead of 1.5).  The two partial accumulators are combined
// through distributed shared memory: each CTA finalises 64 of the 128 columns; it pushes the OTHER 64 columns of
// its partial (128 rows x 256 B) into its peer's receive buffer with st.async — every 16-byte store signals
// complete_tx on the peer's mbarrier, no fence, no global-memory round trip, deterministic summation order —
// then adds what it received to its own half, applies alpha/beta and stores.
// The prologue (barrier init, TMEM allocation, descriptor prefetch) runs under programmatic dependent launch:
// it overlaps the tail of the previous kernel in the stream; griddepcontrol.wait orders every global access.
constexpr int SK_BN = 128;
constexpr int SK_STAGES = 5;
constexpr int SK_STAGE_BYTES = A_TILE_BYTES + SK_BN * BK * 4;          // 32 KB
constexpr int SK_RECV_BYTES = BM * (SK_BN / 2) * 4;                     // 128 rows x 64 fp32 = 32 KB
constexpr int SK_SMEM_BYTES = SK_STAGES * SK_STAGE_BYTES + SK_RECV_BYTES + 1024 + 256;

template <bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
tiramisu_sm100_gemm_tf32_splitk(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                const GemmParams p)
{
    constexpr int BN = SK_BN, STAGES = SK_STAGES, B_TILE_BYTES = BN * BK * 4;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *smem_a = smem;
    uint8_t *smem_b = smem + STAGES * A_TILE_BYTES;
    float *recv = reinterpret_cast<float *>(smem + STAGES * SK_STAGE_BYTES);      // [128 rows][16 x 16-byte chunks], chunk XOR-swizzled by row
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(recv) + SK_RECV_BYTES);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tmem_full_bar = empty_bar + STAGES;
    uint64_t *recv_bar = tmem_full_bar + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(recv_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t ks = ptx::cluster_ctarank();              // which half of K (and which 64 output columns it finalises)

    const int t = blockIdx.x >> 1;
    const int tiles_per_group = p.group_m * p.num_n;
    const int group = t / tiles_per_group;
    const int first_m = group * p.group_m;
    const int gsz = min(p.num_m - first_m, p.group_m);
    const int tile_m = first_m + (t % tiles_per_group) % gsz;
    const int tile_n = (t % tiles_per_group) / gsz;
    const int m0 = tile_m * BM, n0 = tile_n * BN;
    const int num_kb = (p.K + BK - 1) / BK;                  // >= 2 (launcher)
    const int kb_half = (num_kb + 1) / 2;
    const int kb_begin = ks == 0 ? 0 : kb_half, kb_end = ks == 0 ? kb_half : num_kb;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_b);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; s++) {
                ptx::mbar_init(&full_bar[s], 1);
                ptx::mbar_init(&empty_bar[s], 1);
            }
            ptx::mbar_init(tmem_full_bar, 1);
            ptx::mbar_init(recv_bar, 1);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, BN);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before_sync();
    ptx::cluster_sync();                                     // the peer's receive barrier exists before any remote store
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    // everything above overlapped the previous kernel's tail; from here on global memory is touched
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        // the k-slabs beyond the ring go to L2 right away: the whole K half of this CTA is requested from DRAM in
        // the first microsecond instead of ring-depth slabs at a time
        if (p.prefetch_kb && ptx::elect_one()) {
            for (int kb = kb_begin + STAGES; kb < kb_end && kb < kb_begin + STAGES + p.prefetch_kb; kb++) {
                ptx::tma_prefetch_2d(&map_a, kb * BK, m0);
                if (B_MN) {
#pragma unroll
                    for (int j = 0; j < BN / 32; j++) ptx::tma_prefetch_2d(&map_b, n0 + 32 * j, kb * BK);
                } else {
                    ptx::tma_prefetch_2d(&map_b, kb * BK, n0);
                }
            }
        }
        __syncwarp();
        for (int kb = kb_begin, it = 0; kb < kb_end; kb++, it++) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            ptx::mbar_wait(&empty_bar[s], ph ^ 1);
            if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(&full_bar[s], SK_STAGE_BYTES);
                const int k0 = kb * BK;
                ptx::tma_load_2d(smem_a + s * A_TILE_BYTES, &map_a, &full_bar[s], k0, m0);
                uint8_t *bdst = smem_b + s * B_TILE_BYTES;
                if (B_MN) {
#pragma unroll
                    for (int j = 0; j < BN / 32; j++)
                        ptx::tma_load_2d(bdst + j * (BK * 128), &map_b, &full_bar[s], n0 + 32 * j, k0);
                } else {
                    ptx::tma_load_2d(bdst, &map_b, &full_bar[s], k0, n0);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = ptx::make_idesc_tf32(BM, BN, false, B_MN);
        const uint64_t da0 = ptx::make_smem_desc(ptx::smem_u32(smem_a), 16, 1024);
        const uint64_t db0 = B_MN ? ptx::make_smem_desc(ptx::smem_u32(smem_b), p.b_lbo, p.b_sbo, ptx::kLayoutSw128Base32)
                                  : ptx::make_smem_desc(ptx::smem_u32(smem_b), 16, 1024);
        constexpr uint32_t kStepA = UMMA_K * 4, kStepB = B_MN ? UMMA_K * 128 : UMMA_K * 4;
        for (int kb = kb_begin, it = 0; kb < kb_end; kb++, it++) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            ptx::mbar_wait(&full_bar[s], ph);
            ptx::tc_fence_after_sync();
            if (ptx::elect_one()) {
                const uint64_t da = ptx::desc_advance(da0, (uint32_t)s * A_TILE_BYTES);
                const uint64_t db = ptx::desc_advance(db0, (uint32_t)s * B_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; k++)
                    ptx::mma_tf32_ss(tmem_base, ptx::desc_advance(da, k * kStepA), ptx::desc_advance(db, k * kStepB), idesc,
                                     (it | k) != 0);
                ptx::mma_commit(&empty_bar[s]);
                if (kb == kb_end - 1) ptx::mma_commit(tmem_full_bar);
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue: exchange the halves over DSMEM, then TMEM + received -> alpha/beta -> global =====
        const uint32_t q = (uint32_t)warp & 3u;
        const uint32_t row_l = q * 32u + (uint32_t)lane;       // row of the tile this thread owns
        const int row = m0 + (int)row_l;
        const bool row_ok = row < p.M;
        const int c_mine = 64 * (int)ks, c_peer = 64 * (1 - (int)ks);
        float *crow = p.C + (size_t)row * p.ldC;
        const bool use_c = p.beta != 0.0f;
        const float alpha = p.alpha_dev ? __fmul_rn(p.alpha, __ldg(p.alpha_dev)) : p.alpha;
        if (threadIdx.x == 64) ptx::mbar_arrive_expect_tx(recv_bar, (uint32_t)SK_RECV_BYTES);     // the one arrival of this phase
        auto load_c = [&](int c0, float4 (&cv)[8]) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int col = n0 + c0 + 4 * j;
                cv[j] = (use_c && row_ok && col + 3 < p.N) ? *reinterpret_cast<const float4 *>(crow + col)
                                                           : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        float4 cnext[8];
        load_c(c_mine, cnext);
        const uint32_t sw = row_l & 15u;
        const uint32_t push_row = ptx::mapa_u32(reinterpret_cast<uint8_t *>(recv) + row_l * 256u, ks ^ 1u);
        const uint32_t push_bar = ptx::mapa_u32(recv_bar, ks ^ 1u);
        ptx::mbar_wait(tmem_full_bar, 0);
        ptx::tc_fence_after_sync();
        // 1. the peer's 64 columns of my partial -> its receive buffer
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            uint32_t r[32];
            __syncwarp();
            ptx::tmem_ld_32x32(tmem_base + ((q * 32u) << 16) + (uint32_t)(c_peer + 32 * h), r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int cc = 0; cc < 8; cc++)
                ptx::st_async_cluster_v4(push_row + (((uint32_t)(8 * h + cc)) ^ sw) * 16u, push_bar, __uint_as_float(r[4 * cc + 0]),
                                         __uint_as_float(r[4 * cc + 1]), __uint_as_float(r[4 * cc + 2]), __uint_as_float(r[4 * cc + 3]));
        }
        // 2. my 64 columns: own partial + the peer's, in a fixed order (K half 0 + K half 1)
        ptx::mbar_wait_cluster(recv_bar, 0);
        const float4 *rrow = reinterpret_cast<const float4 *>(reinterpret_cast<uint8_t *>(recv) + row_l * 256u);
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            const int c0 = c_mine + 32 * h;
            float4 ccur[8];
#pragma unroll
            for (int j = 0; j < 8; j++) ccur[j] = cnext[j];
            if (h == 0) load_c(c0 + 32, cnext);
            uint32_t r[32];
            __syncwarp();
            ptx::tmem_ld_32x32(tmem_base + ((q * 32u) << 16) + (uint32_t)c0, r);
            ptx::tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 other = rrow[((uint32_t)(8 * h + (j >> 2))) ^ sw];
                    float v[4] = {__uint_as_float(r[j + 0]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])};
                    const float o4[4] = {other.x, other.y, other.z, other.w};
#pragma unroll
                    for (int u = 0; u < 4; u++) v[u] = ks == 0 ? __fadd_rn(v[u], o4[u]) : __fadd_rn(o4[u], v[u]);
                    const int col = n0 + c0 + j;
                    if (col + 3 < p.N) {
                        float4 o;
                        o.x = __fmul_rn(v[0], alpha); o.y = __fmul_rn(v[1], alpha);
                        o.z = __fmul_rn(v[2], alpha); o.w = __fmul_rn(v[3], alpha);
                        if (use_c) {
                            const float4 c = ccur[j >> 2];
                            o.x = __fadd_rn(o.x, __fmul_rn(c.x, p.beta));
                            o.y = __fadd_rn(o.y, __fmul_rn(c.y, p.beta));
                            o.z = __fadd_rn(o.z, __fmul_rn(c.z, p.beta));
                            o.w = __fadd_rn(o.w, __fmul_rn(c.w, p.beta));
                        }
                        *reinterpret_cast<float4 *>(crow + col) = o;
                    } else {
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            if (col + u < p.N) {
                                float o = __fmul_rn(v[u], alpha);
                                if (use_c) o = __fadd_rn(o, __fmul_rn(crow[col + u], p.beta));
                                crow[col + u] = o;
                            }
                        }
                    }
                }
            }
        }
    }

    ptx::tc_fence_before_sync();
    ptx::cluster_sync();                                     // nobody exits while its peer may still push into it
    if (warp == 1) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem_base, BN);
    }
}


// ================================================================================================
// Persistent variant for large problems: one CTA (CG = 1) or one CTA pair (CG = 2, cta_group::2)
// per SM / SM pair loops over 128*CG x 256 output tiles.
//   warp 0    : TMA producer      warp 1 : MMA issuer (leader CTA only when CG = 2) + TMEM owner
//   warps 2-9 : epilogue, two warps per TMEM lane quarter (each takes 128 of the 256 columns)
// Two 256-column accumulator buffers fill TMEM (512 columns): the epilogue of tile i (TMEM ->
// registers -> alpha/beta -> global) overlaps the MMAs of tile i+1, which the one-tile-per-CTA
// kernel above cannot do (r01 ncu: tensor pipe idle during every epilogue, 4096^3 at 56 %).
// CG = 2: the pair shares one 256 x 256 tile; each CTA loads its own 128 rows of A and 128 of the
// 256 columns of B, so the L2 -> shared-memory traffic per FLOP drops by a third (TF32 operands
// are 4 bytes wide: at 128 x 256 per CTA the fill rate, not the tensor pipe, is the limit).
constexpr int P_BN = 256;
constexpr int kThreadsP = 320;
constexpr int kEpiWarps = 8;

template <int CG> struct PCfg {
    static constexpr int B_COLS = P_BN / CG;                       // B columns held by one CTA
    static constexpr int B_TILE_BYTES = B_COLS * BK * 4;
    static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES; // per CTA
    static constexpr int STAGES = (CG == 2) ? 6 : 4;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int CG, bool B_MN>
__global__ void __launch_bounds__(kThreadsP, 1)
tiramisu_sm100_gemm_tf32_persistent(const __grid_constant__ CUtensorMap map_a,
                                    const __grid_constant__ CUtensorMap map_b, const GemmParams p)
{
    using C_ = PCfg<CG>;
    constexpr int STAGES = C_::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *smem_a = smem;
    uint8_t *smem_b = smem + STAGES * A_TILE_BYTES;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * C_::STAGE_BYTES);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tmem_full = empty_bar + STAGES;       // [2]
    uint64_t *tmem_empty = tmem_full + 2;           // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = (CG == 2) ? ptx::cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    const int cluster_id = blockIdx.x / CG;
    const int num_clusters = gridDim.x / CG;
    const int num_tiles = p.num_m * p.num_n;
    const int num_kb = (p.K + BK - 1) / BK;
    constexpr int TILE_M = BM * CG;

    // Tile order: groups of group_m tile rows sweep the columns (A rows of a group and the B columns under the sweep stay
    // in L2).  A gated launch (B arriving column panel by column panel) walks panel by panel — all the tiles of panel j
    // before any tile of panel j+1 — with the same grouped order inside each panel, so that every CTA has work as soon as
    // a panel is in and nothing sweeps ahead into columns that are still on the wire.
    const int sweep_n = p.gate_cols ? p.gate_cols / P_BN : p.num_n;         // tile columns per sweep
    auto tile_origin = [&](int t, int &m0, int &n0) {
        const int per_sweep = p.num_m * sweep_n;
        const int sweep = t / per_sweep;
        t -= sweep * per_sweep;
        const int tiles_per_group = p.group_m * sweep_n;
        const int group = t / tiles_per_group;
        const int first_m = group * p.group_m;
        const int gsz = min(p.num_m - first_m, p.group_m);
        m0 = (first_m + (t % tiles_per_group) % gsz) * TILE_M;
        n0 = (sweep * sweep_n + (t % tiles_per_group) / gsz) * P_BN;
    };

    // The work of this cluster as a sequence of segments (tile, k-slab range); every role walks the same sequence.
    int sk_u_begin = 0, sk_u_end = 0;
    if (p.sk_ws && p.sk_split >= 2) {
        const int r = cluster_id / p.sk_split, i = cluster_id % p.sk_split;          // remainder tile and member
        sk_u_begin = min(p.sk_total, r * num_kb + min(num_kb, i * p.sk_units));
        sk_u_end = min(p.sk_total, r * num_kb + min(num_kb, (i + 1) * p.sk_units));
    } else if (p.sk_ws) {
        sk_u_begin = min(p.sk_total, cluster_id * p.sk_units);
        sk_u_end = min(p.sk_total, sk_u_begin + p.sk_units);
    }
    auto next_segment = [&](uint32_t &i, int &u, int &tile, int &kb0, int &kb1) -> bool {
        if (!p.sk_ws) {
            tile = cluster_id + (int)i * num_clusters;
            if (tile >= num_tiles) return false;
            kb0 = 0; kb1 = num_kb; i++;
            return true;
        }
        if ((int)i < p.sk_dp_waves) { tile = cluster_id + (int)i * num_clusters; kb0 = 0; kb1 = num_kb; i++; return true; }
        if (u >= sk_u_end) return false;
        tile = p.sk_first_tile + u / num_kb;
        kb0 = u % num_kb;
        kb1 = min(num_kb, kb0 + (sk_u_end - u));
        u += kb1 - kb0;
        i++;
        return true;
    };

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_b);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; s++) {
                ptx::mbar_init(&full_bar[s], 1);
                ptx::mbar_init(&empty_bar[s], 1);
            }
            for (int b = 0; b < 2; b++) {
                ptx::mbar_init(&tmem_full[b], 1);
                ptx::mbar_init(&tmem_empty[b], CG * kEpiWarps);
            }
            ptx::fence_barrier_init();
        }
        __syncwarp();
        if (CG == 2) { ptx::tmem_alloc_cg2(tmem_slot, 512); ptx::tmem_relinquish_cg2(); }
        else { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
    }
    ptx::tc_fence_before_sync();
    if (CG == 2) ptx::cluster_sync(); else __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (every CTA loads its own A rows and its own share of B); whole warp,
        // uniform control flow, one elected lane issues =====
        {
            uint32_t it = 0;
            uint32_t arrived = 0;                                   // highest arrival token seen so far (gated launches)
            uint32_t seg = 0;
            int seg_u = sk_u_begin, tile, kb0, kb1;
            while (next_segment(seg, seg_u, tile, kb0, kb1)) {
                int m0, n0;
                tile_origin(tile, m0, n0);
                if (p.gate) {
                    // B is still arriving (peer copies by DMA engines, comm.cu: sgemm_bcast_b): barrier init, TMEM allocation
                    // and descriptor prefetch already ran; wait for the arrival token of this tile's column panel, then order
                    // the generic-proxy acquire before the async-proxy (TMA) reads of the freshly written bytes
                    const uint32_t need = p.gate_value + (p.gate_cols ? (uint32_t)(n0 / p.gate_cols) : 0u);
                    if (arrived < need) {
                        uint32_t v = 0;
                        if (lane == 0) {
                            const long long t0 = clock64();
                            do {
                                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.gate) : "memory");
                                if (v >= need) break;
                                __nanosleep(200);
                                if (clock64() - t0 > 12000000000LL) __trap();      // seconds: a protocol bug must not hang the box
                            } while (true);
                        }
                        arrived = __shfl_sync(0xffffffffu, v, 0);
                        ptx::fence_proxy_async_all();
                    }
                }
                m0 += (int)cta_rank * BM;
                n0 += (int)cta_rank * C_::B_COLS;
                for (int kb = kb0; kb < kb1; kb++, it++) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    ptx::mbar_wait(&empty_bar[s], ph ^ 1);
                    if (ptx::elect_one()) {
                    // L2 prefetch of the slab `prefetch_kb` iterations ahead of the ring (same tile only)
                    if (p.prefetch_kb && !p.b_panel_cols && kb + STAGES + p.prefetch_kb - 1 < kb1) {
                        const int kp = (kb + STAGES + p.prefetch_kb - 1) * BK;
                        ptx::tma_prefetch_2d(&map_a, kp, m0);
                        if (B_MN) {
#pragma unroll
                            for (int j = 0; j < C_::B_COLS / 32; j++) ptx::tma_prefetch_2d(&map_b, n0 + 32 * j, kp);
                        } else {
                            ptx::tma_prefetch_2d(&map_b, kp, n0);
                        }
                    }
                    // the pair's full barrier lives in the leader and expects both CTAs' bytes
                    if (leader) ptx::mbar_arrive_expect_tx(&full_bar[s], CG * C_::STAGE_BYTES);
                    const int k0 = kb * BK;
                    uint8_t *adst = smem_a + s * A_TILE_BYTES;
                    uint8_t *bdst = smem_b + s * C_::B_TILE_BYTES;
                    if (CG == 2) {
                        ptx::tma_load_2d_cg2(adst, &map_a, &full_bar[s], k0, m0);
                        if (B_MN && p.b_panel_cols) {
                            const int pj = n0 / p.b_panel_cols, pn = n0 - pj * p.b_panel_cols;
#pragma unroll
                            for (int j = 0; j < C_::B_COLS / 32; j++)
                                ptx::tma_load_3d_cg2(bdst + j * (BK * 128), &map_b, &full_bar[s], pn + 32 * j, k0, pj);
                        } else if (B_MN) {
#pragma unroll
                            for (int j = 0; j < C_::B_COLS / 32; j++)
                                ptx::tma_load_2d_cg2(bdst + j * (BK * 128), &map_b, &full_bar[s], n0 + 32 * j, k0);
                        } else {
                            ptx::tma_load_2d_cg2(bdst, &map_b, &full_bar[s], k0, n0);
                        }
                    } else {
                        ptx::tma_load_2d(adst, &map_a, &full_bar[s], k0, m0);
                        if (B_MN && p.b_panel_cols) {
                            const int pj = n0 / p.b_panel_cols, pn = n0 - pj * p.b_panel_cols;
#pragma unroll
                            for (int j = 0; j < C_::B_COLS / 32; j++)
                                ptx::tma_load_3d(bdst + j * (BK * 128), &map_b, &full_bar[s], pn + 32 * j, k0, pj);
                        } else if (B_MN) {
#pragma unroll
                            for (int j = 0; j < C_::B_COLS / 32; j++)
                                ptx::tma_load_2d(bdst + j * (BK * 128), &map_b, &full_bar[s], n0 + 32 * j, k0);
                        } else {
                            ptx::tma_load_2d(bdst, &map_b, &full_bar[s], k0, n0);
                        }
                    }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: whole warp, uniform control flow, one elected lane issues =====
        if (leader) {
            constexpr uint32_t idesc = ptx::make_idesc_tf32(TILE_M, P_BN, false, B_MN);
            const uint64_t da0 = ptx::make_smem_desc(ptx::smem_u32(smem_a), 16, 1024);
            const uint64_t db0 = B_MN ? ptx::make_smem_desc(ptx::smem_u32(smem_b), p.b_lbo, p.b_sbo, ptx::kLayoutSw128Base32)
                                      : ptx::make_smem_desc(ptx::smem_u32(smem_b), 16, 1024);
            constexpr uint32_t kStepA = UMMA_K * 4, kStepB = B_MN ? UMMA_K * 128 : UMMA_K * 4;
            uint32_t it = 0, local = 0, seg = 0;
            int seg_u = sk_u_begin, tile, kb0, kb1;
            for (; next_segment(seg, seg_u, tile, kb0, kb1); local++) {
                const uint32_t buf = local & 1;
                ptx::mbar_wait(&tmem_empty[buf], ((local >> 1) & 1) ^ 1);      // epilogue drained this buffer
                ptx::tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + buf * P_BN;
                for (int kb = kb0; kb < kb1; kb++, it++) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    ptx::mbar_wait(&full_bar[s], ph);
                    ptx::tc_fence_after_sync();
                    if (ptx::elect_one()) {
                        const uint64_t da = ptx::desc_advance(da0, (uint32_t)s * A_TILE_BYTES);
                        const uint64_t db = ptx::desc_advance(db0, (uint32_t)s * C_::B_TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; k++) {
                            if (CG == 2)
                                ptx::mma_tf32_ss_cg2(d_tmem, ptx::desc_advance(da, k * kStepA),
                                                     ptx::desc_advance(db, k * kStepB), idesc, ((kb - kb0) | k) != 0);
                            else
                                ptx::mma_tf32_ss(d_tmem, ptx::desc_advance(da, k * kStepA),
                                                 ptx::desc_advance(db, k * kStepB), idesc, ((kb - kb0) | k) != 0);
                        }
                        if (CG == 2) ptx::mma_commit_cg2(&empty_bar[s], 3); else ptx::mma_commit(&empty_bar[s]);
                        if (kb == kb1 - 1) {
                            if (CG == 2) ptx::mma_commit_cg2(&tmem_full[buf], 3); else ptx::mma_commit(&tmem_full[buf]);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> global, overlapped with the next tile's MMAs =====
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;              // which 128 of the 256 accumulator columns
        const bool use_c = p.beta != 0.0f;
        uint32_t local = 0, seg = 0;
        int seg_u = sk_u_begin, tile, kb0, kb1;
        const int wslot = (int)cta_rank * kEpiWarps + (warp - 2);             // this warp's flag / position inside a cluster
        for (; next_segment(seg, seg_u, tile, kb0, kb1); local++) {
            const uint32_t buf = local & 1;
            int m0, n0;
            tile_origin(tile, m0, n0);
            const int row = m0 + (int)cta_rank * BM + q * 32 + lane;
            const bool row_ok = row < p.M;
            float *crow = p.C + (size_t)row * p.ldC;
            const int ws_row = (int)cta_rank * BM + q * 32 + lane;              // row inside the cluster's 256-column slot
            if (kb0 > 0) {
                // ---- stream-K contribution: the run started inside this tile; leave the raw accumulator in this cluster's
                // workspace slot for the owner (the same warp position there reads exactly what this warp writes)
                float *wrow = p.sk_ws + ((size_t)cluster_id * TILE_M + ws_row) * P_BN;
                ptx::mbar_wait(&tmem_full[buf], (local >> 1) & 1);
                ptx::tc_fence_after_sync();
#pragma unroll 1
                for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
                    uint32_t r[32];
                    __syncwarp();
                    ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * P_BN + (uint32_t)c0, r);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4 *>(wrow + c0 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                                  __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                }
                ptx::tc_fence_before_sync();
                __threadfence();
                __syncwarp();
                if (lane == 0) {
                    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.sk_flags + (size_t)cluster_id * (2 * kEpiWarps) + wslot), "r"(1u) : "memory");
                    if (CG == 2) ptx::mbar_arrive_remote(&tmem_empty[buf], 0);
                    else ptx::mbar_arrive(&tmem_empty[buf]);
                }
                continue;
            }
            // owner of a tile other clusters continue: their contributions are added below (clusters c+1 .. sk_last)
            int sk_last = cluster_id;
            if (p.sk_ws && kb1 < num_kb)
                sk_last = p.sk_split >= 2 ? cluster_id + (num_kb - 1) / p.sk_units
                                          : min(num_clusters - 1, ((tile - p.sk_first_tile + 1) * num_kb - 1) / p.sk_units);
            // beta*C does not depend on the MMAs: fetch each 32-column chunk one iteration ahead, the
            // first one before waiting for the accumulator (matters for the last tile of a CTA,
            // whose epilogue nothing overlaps)
            auto load_c = [&](int c0, float4 (&cv)[8]) {
#pragma unroll
                for (int jj = 0; jj < 8; jj++) {
                    const int col = n0 + c0 + 4 * jj;
                    cv[jj] = (use_c && row_ok && col + 3 < p.N) ? *reinterpret_cast<const float4 *>(crow + col)
                                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            float4 cnext[8];
            load_c(half * 128, cnext);
            ptx::mbar_wait(&tmem_full[buf], (local >> 1) & 1);
            ptx::tc_fence_after_sync();
            for (int c = cluster_id + 1; c <= sk_last; c++) {       // this warp's counterpart in every contributing cluster is done
                if (lane == 0) {
                    const uint32_t *f = p.sk_flags + (size_t)c * (2 * kEpiWarps) + wslot;
                    const long long t0 = clock64();
                    uint32_t v;
                    do {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                        if (clock64() - t0 > (1ll << 32)) __trap();
                    } while (v == 0);
                }
                __syncwarp();
            }
#pragma unroll 1
            for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
                float4 ccur[8];
#pragma unroll
                for (int jj = 0; jj < 8; jj++) ccur[jj] = cnext[jj];
                if (c0 + 32 < half * 128 + 128) load_c(c0 + 32, cnext);
                uint32_t r[32];
                __syncwarp();
                ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * P_BN + (uint32_t)c0, r);
                ptx::tmem_ld_wait();
                for (int c = cluster_id + 1; c <= sk_last; c++) {   // partial sums in a fixed order: owner, then c+1, c+2, ...
                    const float *wrow = p.sk_ws + ((size_t)c * TILE_M + ws_row) * P_BN + c0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 w = __ldcg(reinterpret_cast<const float4 *>(wrow + j));
                        r[j + 0] = __float_as_uint(__fadd_rn(__uint_as_float(r[j + 0]), w.x));
                        r[j + 1] = __float_as_uint(__fadd_rn(__uint_as_float(r[j + 1]), w.y));
                        r[j + 2] = __float_as_uint(__fadd_rn(__uint_as_float(r[j + 2]), w.z));
                        r[j + 3] = __float_as_uint(__fadd_rn(__uint_as_float(r[j + 3]), w.w));
                    }
                }
                if (row_ok) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const int col = n0 + c0 + j;
                        if (col + 3 < p.N) {
                            float4 o;
                            o.x = __fmul_rn(__uint_as_float(r[j + 0]), p.alpha);
                            o.y = __fmul_rn(__uint_as_float(r[j + 1]), p.alpha);
                            o.z = __fmul_rn(__uint_as_float(r[j + 2]), p.alpha);
                            o.w = __fmul_rn(__uint_as_float(r[j + 3]), p.alpha);
                            if (use_c) {
                                const float4 c = ccur[j >> 2];
                                o.x = __fadd_rn(o.x, __fmul_rn(c.x, p.beta));
                                o.y = __fadd_rn(o.y, __fmul_rn(c.y, p.beta));
                                o.z = __fadd_rn(o.z, __fmul_rn(c.z, p.beta));
                                o.w = __fadd_rn(o.w, __fmul_rn(c.w, p.beta));
                            }
                            *reinterpret_cast<float4 *>(crow + col) = o;
                        } else {
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                if (col + u < p.N) {
                                    float o = __fmul_rn(__uint_as_float(r[j + u]), p.alpha);
                                    if (use_c) o = __fadd_rn(o, __fmul_rn(crow[col + u], p.beta));
                                    crow[col + u] = o;
                                }
                            }
                        }
                    }
                }
            }
            // this warp has read its part of the buffer: let the MMA issuer (in the leader) reuse it
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                for (int c = cluster_id + 1; c <= sk_last; c++)      // consumed: the flags are 0 again when the kernel ends
                    p.sk_flags[(size_t)c * (2 * kEpiWarps) + wslot] = 0u;
                if (CG == 2) ptx::mbar_arrive_remote(&tmem_empty[buf], 0);
                else ptx::mbar_arrive(&tmem_empty[buf]);
            }
        }
    }

    ptx::tc_fence_before_sync();
    if (CG == 2) ptx::cluster_sync(); else __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after_sync();
        if (CG == 2) ptx::tmem_dealloc_cg2(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
    }
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// L2 prefetch distance (k-slabs beyond the ring); TB200_GEMM_PREFETCH overrides the kernel's default
inline int gemm_prefetch_kb(int dflt)
{
    const char *e = getenv("TB200_GEMM_PREFETCH");
    const int env = e ? atoi(e) : -1;
    return env >= 0 ? env : dflt;
}

template <int BN, bool B_MN>
int launch_tc(const float *A, const float *B, float *C, int M, int N, int K, float alpha, float beta, int64_t ldA,
              int64_t ldB, int64_t ldC)
{
    CUtensorMap map_a, map_b;
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
        uint64_t str[1] = {(uint64_t)ldA * 4};
        uint32_t box[2] = {BK, BM};
        if (make_tensor_map_f32(&map_a, A, 2, dims, str, box)) return 1;
    }
    if (B_MN) {
        uint64_t dims[2] = {(uint64_t)N, (uint64_t)K};
        uint64_t str[1] = {(uint64_t)ldB * 4};
        uint32_t box[2] = {32, BK};
        // MN-major 32-bit operands need the 128B swizzle with 32-byte atoms (UMMA layout
        // SWIZZLE_128B_BASE32B); the plain 128B swizzle is only valid for K-major tf32.
        if (make_tensor_map_f32(&map_b, B, 2, dims, str, box, kTmaSwizzle128Atom32)) return 1;
    } else {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
        uint64_t str[1] = {(uint64_t)ldB * 4};
        uint32_t box[2] = {BK, (uint32_t)BN};
        if (make_tensor_map_f32(&map_b, B, 2, dims, str, box)) return 1;
    }
    GemmParams p;
    p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.beta = beta; p.C = C; p.ldC = ldC;
    p.num_m = (M + BM - 1) / BM;
    p.num_n = (N + BN - 1) / BN;
    p.group_m = 16;
    p.b_lbo = BK * 128;
    p.b_sbo = 512;
    p.alpha_dev = nullptr;
    p.prefetch_kb = 0;
    if (const char *e = getenv("TB200_DBG_B_LBO")) p.b_lbo = (uint32_t)atoi(e);     // descriptor bring-up only
    if (const char *e = getenv("TB200_DBG_B_SBO")) p.b_sbo = (uint32_t)atoi(e);
    auto kern = tiramisu_sm100_gemm_tf32<BN, B_MN>;
    static bool attr_set = false;
    if (!attr_set) {
        TB200_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_BYTES));
        attr_set = true;
    }
    kern<<<p.num_m * p.num_n, kThreads, Cfg<BN>::SMEM_BYTES, stream()>>>(map_a, map_b, p);
    return post_launch("tiramisu_sm100_gemm_tf32");
}

template <bool B_MN>
int launch_tc_splitk(const float *A, const float *B, float *C, int M, int N, int K, float alpha, float beta, int64_t ldA,
                     int64_t ldB, int64_t ldC)
{
    CUtensorMap map_a, map_b;
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
        uint64_t str[1] = {(uint64_t)ldA * 4};
        uint32_t box[2] = {BK, BM};
        if (make_tensor_map_f32(&map_a, A, 2, dims, str, box)) return 1;
    }
    if (B_MN) {
        uint64_t dims[2] = {(uint64_t)N, (uint64_t)K};
        uint64_t str[1] = {(uint64_t)ldB * 4};
        uint32_t box[2] = {32, BK};
        if (make_tensor_map_f32(&map_b, B, 2, dims, str, box, kTmaSwizzle128Atom32)) return 1;
    } else {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
        uint64_t str[1] = {(uint64_t)ldB * 4};
        uint32_t box[2] = {BK, (uint32_t)SK_BN};
        if (make_tensor_map_f32(&map_b, B, 2, dims, str, box)) return 1;
    }
    GemmParams p;
    p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.beta = beta; p.C = C; p.ldC = ldC;
    p.num_m = (M + BM - 1) / BM;
    p.num_n = (N + SK_BN - 1) / SK_BN;
    p.group_m = 16;
    p.b_lbo = BK * 128;
    p.b_sbo = 512;
    p.alpha_dev = nullptr;
    p.prefetch_kb = gemm_prefetch_kb(0);        // measured (r02 run 2): prefetching the K half ahead of the ring is slower at 1024^3
    auto kern = tiramisu_sm100_gemm_tf32_splitk<B_MN>;
    static bool attr_set = false;
    if (!attr_set) {
        TB200_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM_BYTES));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(p.num_m * p.num_n * 2));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = SK_SMEM_BYTES;
    cfg.stream = stream();
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = getenv("TB200_GEMM_NO_PDL") ? 1 : 2;
    TB200_CHECK(cudaLaunchKernelEx(&cfg, kern, map_a, map_b, p));
    return post_launch("tiramisu_sm100_gemm_tf32_splitk");
}

template <int BN>
int launch_tc_f16(const void *A, const void *B, float *C, int M, int N, int K, float alpha, float beta, int64_t ldA,
                  int64_t ldB, int64_t ldC, const float *alpha_dev)
{
    CUtensorMap map_a, map_b;
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
        uint64_t str[1] = {(uint64_t)ldA * 2};
        uint32_t box[2] = {64, BM};
        if (make_tensor_map_f16(&map_a, A, 2, dims, str, box)) return 1;
    }
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
        uint64_t str[1] = {(uint64_t)ldB * 2};
        uint32_t box[2] = {64, (uint32_t)BN};
        if (make_tensor_map_f16(&map_b, B, 2, dims, str, box)) return 1;
    }
    GemmParams p;
    p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.beta = beta; p.C = C; p.ldC = ldC;
    p.num_m = (M + BM - 1) / BM;
    p.num_n = (N + BN - 1) / BN;
    p.group_m = 16;
    p.b_lbo = 0;
    p.b_sbo = 0;
    p.alpha_dev = alpha_dev;
    p.prefetch_kb = 0;
    auto kern = tiramisu_sm100_gemm_tf32<BN, false, true>;
    static bool attr_set = false;
    if (!attr_set) {
        TB200_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_BYTES));
        attr_set = true;
    }
    kern<<<p.num_m * p.num_n, kThreads, Cfg<BN>::SMEM_BYTES, stream()>>>(map_a, map_b, p);
    return post_launch("tiramisu_sm100_gemm_f16");
}

// Stream-K workspace: one fp32 slot of TILE_M x 256 per cluster + one flag per epilogue warp and cluster, per (thread, stream)
// — launches on one stream are ordered, launches on different streams must not share slots.  Flags are 0 between launches.
struct StreamKWorkspace { float *partial = nullptr; uint32_t *flags = nullptr; int clusters = 0, tile_m = 0; };
StreamKWorkspace *streamk_workspace(cudaStream_t st, int clusters, int tile_m)
{
    static thread_local std::vector<std::pair<cudaStream_t, StreamKWorkspace>> *all = new std::vector<std::pair<cudaStream_t, StreamKWorkspace>>();
    for (auto &e : *all)
        if (e.first == st && e.second.clusters >= clusters && e.second.tile_m >= tile_m) return &e.second;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) { cudaGetLastError(); return nullptr; }   // no allocation inside a capture
    StreamKWorkspace w;
    w.clusters = clusters; w.tile_m = tile_m;
    const size_t bytes = (size_t)clusters * tile_m * P_BN * 4, fbytes = (size_t)clusters * 2 * kEpiWarps * 4;
    if (cudaMalloc(&w.partial, bytes) != cudaSuccess || cudaMalloc(&w.flags, fbytes) != cudaSuccess || cudaMemset(w.flags, 0, fbytes) != cudaSuccess) {
        cudaGetLastError();
        if (w.partial) cudaFree(w.partial);
        return nullptr;
    }
    all->push_back({st, w});
    return &all->back().second;
}

template <int CG, bool B_MN>
int launch_tc_persistent(const float *A, const float *B, float *C, int M, int N, int K, float alpha, float beta,
                         int64_t ldA, int64_t ldB, int64_t ldC)
{
    using C_ = PCfg<CG>;
    CUtensorMap map_a, map_b;
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
        uint64_t str[1] = {(uint64_t)ldA * 4};
        uint32_t box[2] = {BK, BM};
        if (make_tensor_map_f32(&map_a, A, 2, dims, str, box)) return 1;
    }
    const int panel_cols = B_MN ? tb200::t_b_panel_cols : 0;
    if (B_MN && panel_cols) {
        if (panel_cols % P_BN || N % panel_cols || ldB != panel_cols) return fail("gemm_f32_tc", "panel-major B: bad panel width");
        uint64_t dims[3] = {(uint64_t)panel_cols, (uint64_t)K, (uint64_t)(N / panel_cols)};
        uint64_t str[2] = {(uint64_t)panel_cols * 4, tb200::t_b_panel_stride * 4};
        uint32_t box[3] = {32, BK, 1};
        if (make_tensor_map_f32(&map_b, B, 3, dims, str, box, kTmaSwizzle128Atom32)) return 1;
    } else if (B_MN) {
        uint64_t dims[2] = {(uint64_t)N, (uint64_t)K};
        uint64_t str[1] = {(uint64_t)ldB * 4};
        uint32_t box[2] = {32, BK};
        if (make_tensor_map_f32(&map_b, B, 2, dims, str, box, kTmaSwizzle128Atom32)) return 1;
    } else {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
        uint64_t str[1] = {(uint64_t)ldB * 4};
        uint32_t box[2] = {BK, (uint32_t)C_::B_COLS};
        if (make_tensor_map_f32(&map_b, B, 2, dims, str, box)) return 1;
    }
    GemmParams p;
    p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.beta = beta; p.C = C; p.ldC = ldC;
    p.num_m = (M + BM * CG - 1) / (BM * CG);
    p.num_n = (N + P_BN - 1) / P_BN;
    p.group_m = 16 / CG;
    p.b_panel_cols = panel_cols;
    p.b_lbo = BK * 128;
    p.b_sbo = 512;
    p.alpha_dev = nullptr;
    p.prefetch_kb = gemm_prefetch_kb(0);
    p.gate = tb200::t_gate; p.gate_value = tb200::t_gate_value; p.gate_cols = tb200::t_gate_cols;
    tb200::t_gate = nullptr;
    auto kern = tiramisu_sm100_gemm_tf32_persistent<CG, B_MN>;
    static bool attr_set = false;
    if (!attr_set) {
        TB200_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM_BYTES));
        attr_set = true;
    }
    const int num_tiles = p.num_m * p.num_n;
    int clusters = sm_count() / CG;
    // The last, partly filled wave: with T tiles on P clusters it holds T mod P tiles (4096^3: 256 = 3 waves + 34) and the
    // kernel lasts ceil(T / P) tile times.  When at least two clusters are free per remainder tile, the k-slabs of each
    // remainder tile are split between s = min(4, P / rem) clusters (GemmParams::sk_split: k-aligned, the clusters of one
    // k-phase stay in lockstep); member 0 owns the tile and adds the others' raw accumulators in its epilogue through one
    // workspace slot per cluster.  TB200_GEMM_STREAMK=0: off; =1: classic stream-K (equal contiguous runs over all remainder
    // slabs; measured slower, see GemmParams) — kept for the record and for its parity tests.
    const int num_kb = (K + BK - 1) / BK;
    const int rem = num_tiles % clusters;
    int sk_mode = 2;
    if (const char *e = getenv("TB200_GEMM_STREAMK")) sk_mode = atoi(e);
    if (getenv("TB200_GEMM_NO_STREAMK")) sk_mode = 0;
    int split = rem ? clusters / rem : 0;
    if (split > 4) split = 4;
    bool streamk = false;
    if (!p.gate && rem != 0 && sk_mode == 1)
        streamk = (int64_t)rem * num_kb >= (int64_t)clusters * 16 && (num_tiles < clusters ? rem * 10 < clusters * 9 + clusters / 2 : true);
    else if (!p.gate && rem != 0 && sk_mode == 2)
        streamk = split >= 2 && num_kb >= 16 * split;
    if (streamk) {
        StreamKWorkspace *ws = streamk_workspace(stream(), clusters, BM * CG);
        if (!ws) streamk = false;
        else {
            p.sk_ws = ws->partial; p.sk_flags = ws->flags;
            p.sk_dp_waves = num_tiles / clusters;
            p.sk_first_tile = p.sk_dp_waves * clusters;
            p.sk_total = rem * num_kb;
            if (sk_mode == 2) { p.sk_split = split; p.sk_units = (num_kb + split - 1) / split; }
            else p.sk_units = (p.sk_total + clusters - 1) / clusters;
        }
    }
    if (!streamk && clusters > num_tiles) clusters = num_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * CG));
    cfg.blockDim = dim3(kThreadsP);
    cfg.dynamicSmemBytes = C_::SMEM_BYTES;
    cfg.stream = stream();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TB200_CHECK(cudaLaunchKernelEx(&cfg, kern, map_a, map_b, p));
    return post_launch("tiramisu_sm100_gemm_tf32_persistent");
}

std::atomic<int> g_gemm_mode{0};
std::atomic<int> g_gemm_kernel{0};       // 0 auto, 1 one-tile-per-CTA, 2 persistent CG=1, 3 persistent CG=2, 4 split-K 2-CTA clusters

}  // namespace

namespace tb200 {

int gemm_f32_exact(const float *A, const float *B, float *C, uint64_t M, uint64_t N, uint64_t K, float alpha,
                   float beta, uint64_t ldA, uint64_t ldB, uint64_t ldC, bool tA, bool tB);

// Can the tensor-core kernel address these operands?  TMA needs 16-byte aligned bases and row
// pitches; the epilogue needs 16-byte aligned C rows.  A must be K-major (no transposeA).
bool gemm_tc_addressable(const float *A, const float *B, const float *C, uint64_t M, uint64_t N, uint64_t K,
                         uint64_t ldA, uint64_t ldB, uint64_t ldC, bool tA, bool tB)
{
    (void)tB;
    if (tA) return false;
    if (M == 0 || N == 0 || K == 0) return false;
    if (M > (1u << 30) || N > (1u << 30) || K > (1u << 30)) return false;
    if (!aligned16(A) || !aligned16(B) || !aligned16(C)) return false;
    if (ldA % 4 || ldB % 4 || ldC % 4) return false;
    return true;
}

// Kernel choice: the persistent CTA-pair kernel once there is at least ~one 256x256 tile per SM
// pair, the persistent single-CTA kernel for mid sizes, the one-tile-per-CTA kernel below that.
// 1 one tile per CTA (or 4, split-K, decided later), 2 persistent, 3 persistent CTA pairs.
int gemm_tc_kernel_choice(uint64_t M, uint64_t N, uint64_t K, bool *was_auto)
{
    int kernel = g_gemm_kernel.load();
    if (const char *e = getenv("TB200_GEMM_KERNEL")) kernel = atoi(e);
    if (t_hint_kernel) kernel = t_hint_kernel;
    if (was_auto) *was_auto = kernel == 0;
    if (kernel == 0) {
        // measured (profiles/r01_results.md): the pair kernel wins from ~2048^2 outputs upwards even
        // when it cannot fill every SM pair (2048^3: 416 vs 339 TFLOP/s); below that the one-tile-per-CTA
        // kernel with narrow tiles (more CTAs, shorter critical path) is faster (1024^3: 147 vs 73).
        const uint64_t sms = (uint64_t)sm_count();
        const uint64_t pair_tiles = ((M + 255) / 256) * ((N + 255) / 256);
        kernel = (pair_tiles * 2 * 5 >= sms * 4 && K >= 256) ? 3 : 1;
    }
    return kernel;
}

int gemm_f32_tc(const float *A, const float *B, float *C, uint64_t M, uint64_t N, uint64_t K, float alpha,
                float beta, uint64_t ldA, uint64_t ldB, uint64_t ldC, bool tB)
{
    // Very long K on a large output: run K panels of 4096 with beta = 1 accumulation.  Without a
    // device-wide sync the persistent CTAs drift apart along K until the A/B slabs that concurrent
    // tiles share fall out of L2 (16384^3: 547 TFLOP/s in one launch, 742 in four panels, r01);
    // a panel boundary re-aligns them at the price of one extra read+write of C per panel.
    if (K >= 12288 && M >= 4096 && N >= 4096 && !getenv("TB200_GEMM_NO_KPANELS")) {
        const uint64_t panel = 4096;
        for (uint64_t k0 = 0; k0 < K; k0 += panel) {
            const uint64_t kk = (K - k0 < panel + panel / 2) ? K - k0 : panel;     // fold a short tail into the last panel
            const float *Ap = A + k0;                                            // A is K-major (no transposeA here)
            const float *Bp = tB ? B + k0 : B + k0 * ldB;
            if (gemm_f32_tc(Ap, Bp, C, M, N, kk, alpha, k0 == 0 ? beta : 1.0f, ldA, ldB, ldC, tB)) return 1;
            if (kk != panel) break;
        }
        return 0;
    }
    const int m = (int)M, n = (int)N, k = (int)K;
    const uint64_t sms = (uint64_t)sm_count();
    bool auto_kernel = false;
    int kernel = gemm_tc_kernel_choice(M, N, K, &auto_kernel);
    if (t_gate && ((kernel != 3 && kernel != 2) || getenv("TB200_SGEMM_GATE_ON_STREAM"))) {   // no in-kernel gate in the one-tile kernels (or diagnosis): wait on the stream
        const uint32_t *g = t_gate;
        t_gate = nullptr;
        if (stream_wait_value32(stream(), g, t_gate_last)) return 1;
    }
    if (kernel == 3)
        return tB ? launch_tc_persistent<2, false>(A, B, C, m, n, k, alpha, beta, ldA, ldB, ldC)
                  : launch_tc_persistent<2, true>(A, B, C, m, n, k, alpha, beta, ldA, ldB, ldC);
    if (kernel == 2)
        return tB ? launch_tc_persistent<1, false>(A, B, C, m, n, k, alpha, beta, ldA, ldB, ldC)
                  : launch_tc_persistent<1, true>(A, B, C, m, n, k, alpha, beta, ldA, ldB, ldC);
    // Small outputs with a long K (1024^3): 2-CTA split-K clusters on 128 x 128 tiles when that fills the machine at
    // least as well as narrow tiles would (kernel 4; TB200_GEMM_KERNEL=1 keeps the one-tile-per-CTA kernel)
    {
        const uint64_t t128 = ((M + BM - 1) / BM) * ((N + SK_BN - 1) / SK_BN);
        const bool fits = 2 * t128 <= sms && K >= 512 && 2 * t128 * 4 >= sms;      // at most one wave, at least a quarter of the SMs
        if (kernel == 4 || (auto_kernel && kernel == 1 && !t_hint_bn && fits))
            if (K >= 2 * BK)
                return tB ? launch_tc_splitk<false>(A, B, C, m, n, k, alpha, beta, ldA, ldB, ldC)
                          : launch_tc_splitk<true>(A, B, C, m, n, k, alpha, beta, ldA, ldB, ldC);
    }
    // Tile width: the widest BN that still yields at least one CTA per SM (148), else 64.
    int bn = 64;
    const uint64_t mt = (M + BM - 1) / BM;
    if (mt * ((N + 255) / 256) >= sms) bn = 256;
    else if (mt * ((N + 127) / 128) >= sms) bn = 128;
    if (const char *e = getenv("TB200_GEMM_BN")) {
        int v = atoi(e);
        if (v == 64 || v == 128 || v == 256) bn = v;
    }
    if (t_hint_bn == 64 || t_hint_bn == 128 || t_hint_bn == 256) bn = t_hint_bn;
    if (tB) {
        switch (bn) {
        case 256: return launch_tc<256, false>(A, B, C, m, n, k, alpha, beta, ldA, ldB, ldC);
        case 128: return launch_tc<128, false>(A, B, C, m, n, k, alpha, beta, ldA, ldB, ldC);
        default: return launch_tc<64, false>(A, B, C, m, n, k, alpha, beta, ldA, ldB, ldC);
        }
    }
    switch (bn) {
    case 256: return launch_tc<256, true>(A, B, C, m, n, k, alpha, beta, ldA, ldB, ldC);
    case 128: return launch_tc<128, true>(A, B, C, m, n, k, alpha, beta, ldA, ldB, ldC);
    default: return launch_tc<64, true>(A, B, C, m, n, k, alpha, beta, ldA, ldB, ldC);
    }
}

// C[M][N] (fp32) = alpha * A[M][K] * B[N][K]^T + beta * C with fp16 operands (both K-major, pitches in
// elements, multiples of 8), fp32 accumulation; one 128 x 256 (or 128 x 128) tile per CTA.  alpha_dev: optional
// device-resident factor multiplied into alpha (read by the epilogue, no host synchronisation).
int gemm_f16_tc(const void *A16, const void *B16, float *C, uint64_t M, uint64_t N, uint64_t K, float alpha, float beta,
                uint64_t ldA, uint64_t ldB, uint64_t ldC, const float *alpha_dev)
{
    if (!aligned16(A16) || !aligned16(B16) || !aligned16(C) || ldA % 8 || ldB % 8 || ldC % 4 || M == 0 || N == 0 || K == 0)
        return fail("gemm_f16_tc", "operands not addressable by TMA");
    if (N % 256 == 0 || N > 1024)
        return launch_tc_f16<256>(A16, B16, C, (int)M, (int)N, (int)K, alpha, beta, ldA, ldB, ldC, alpha_dev);
    return launch_tc_f16<128>(A16, B16, C, (int)M, (int)N, (int)K, alpha, beta, ldA, ldB, ldC, alpha_dev);
}

int gemm_f32(const float *A, const float *B, float *C, uint64_t M, uint64_t N, uint64_t K, float alpha, float beta,
             uint64_t ldA, uint64_t ldB, uint64_t ldC, bool tA, bool tB)
{
    int mode = g_gemm_mode.load();
    if (const char *e = getenv("TB200_GEMM_MODE")) mode = atoi(e);     // for harnesses that cannot call the setter
    const bool ok = gemm_tc_addressable(A, B, C, M, N, K, ldA, ldB, ldC, tA, tB);
    if (mode == 2 && !ok) return fail("tiramisu_cublas_sgemm", "operands not addressable by the tcgen05 path");
    if (mode != 1 && ok) return gemm_f32_tc(A, B, C, M, N, K, alpha, beta, ldA, ldB, ldC, tB);
    if (t_gate) {                                            // arrival gate of the B operand: stream-side on this path
        const uint32_t *g = t_gate;
        t_gate = nullptr;
        if (stream_wait_value32(stream(), g, t_gate_last)) return 1;
    }
    return gemm_f32_exact(A, B, C, M, N, K, alpha, beta, ldA, ldB, ldC, tA, tB);
}

}  // namespace tb200

TB200_API int tiramisu_b200_set_gemm_mode(int mode)
{
    if (mode < 0 || mode > 2) mode = 0;
    return g_gemm_mode.exchange(mode);
}

TB200_API int tiramisu_b200_set_gemm_kernel(int kernel)
{
    if (kernel < 0 || kernel > 4) kernel = 0;
    return g_gemm_kernel.exchange(kernel);
}

TB200_API int tiramisu_cublas_sgemm(float *A, float *B, float *C, uint64_t M, uint64_t N, uint64_t K, float alpha,
                                    float beta, uint64_t ldA, uint64_t ldB, uint64_t ldC, uint64_t offsetA,
                                    uint64_t offsetB, uint64_t offsetC, bool transposeA, bool transposeB)
{
    if (!A || !B || !C) return fail(__func__, "null operand");
    // src/tiramisu_cuda_wrappers.cpp:90-98: 0 means tightly packed
    if (ldA == 0) ldA = transposeA ? M : K;
    if (ldB == 0) ldB = transposeB ? K : N;
    if (ldC == 0) ldC = N;
    if (M == 0 || N == 0) return 0;
    return gemm_f32(A + offsetA, B + offsetB, C + offsetC, M, N, K, alpha, beta, ldA, ldB, ldC, transposeA,
                    transposeB);
}
