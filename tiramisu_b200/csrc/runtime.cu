// runtime.cu — thin C-ABI runtime replacing /root/reference/src/tiramisu_cuda_wrappers.cpp and
// src/tiramisu_externs.cpp: stream-ordered pooled allocation, async copies, per-thread launch
// stream, launch accounting.  Same symbol names and argument meaning as the reference so that
// Tiramisu-generated host code links against it unchanged.
#include "common.cuh"
#include "tiramisu_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace tb200 {

std::atomic<uint64_t> g_kernel_launches{0};

static std::mutex g_err_mu;
static std::string g_last_error;
static std::atomic<int> g_error_mode{0};     // 0 = exit(1) like the reference, 1 = return codes

static thread_local cudaStream_t t_stream = nullptr;
static thread_local bool t_stream_set = false;

void set_last_error(const std::string &msg)
{
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_last_error = msg;
}

int fail(const char *where, const std::string &msg)
{
    std::string full = std::string("Error at ") + where + ": " + msg;
    set_last_error(full);
    if (g_error_mode.load() == 0) {
        // src/tiramisu_cuda_wrappers.cpp:18-21: message on stderr, exit(1)
        fprintf(stderr, "%s\n", full.c_str());
        exit(1);
    }
    return 1;
}

int check(cudaError_t e, const char *where)
{
    if (e == cudaSuccess) return 0;
    return fail(where, cudaGetErrorString(e));
}

cudaStream_t stream() { return t_stream_set ? t_stream : cudaStreamPerThread; }

ScopedStream::ScopedStream(cudaStream_t s) : prev_(t_stream), prev_set_(t_stream_set)
{
    t_stream = s;
    t_stream_set = true;
}
ScopedStream::~ScopedStream()
{
    t_stream = prev_;
    t_stream_set = prev_set_;
}

int post_launch(const char *kernel_name)
{
    count_launch();
    return check(cudaGetLastError(), kernel_name);
}

int sm_count()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// ---- pooled, stream-ordered allocation --------------------------------------------------------
// The reference cudaMalloc/cudaFree's every temporary on every call
// (src/tiramisu_codegen_halide.cpp:2588-2603, :2488-2503).  Here allocations come from the
// device's default memory pool with an unbounded release threshold, so steady-state calls
// neither hit the driver nor synchronise.
static void configure_pool_once()
{
    static std::once_flag once;
    std::call_once(once, [] {
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    });
}

void *pool_alloc(size_t bytes)
{
    configure_pool_once();
    void *p = nullptr;
    if (bytes == 0) bytes = 1;
    if (check(cudaMallocAsync(&p, bytes, stream()), "tiramisu_cuda_malloc")) return nullptr;
    return p;
}

void pool_free(void *p)
{
    if (p) check(cudaFreeAsync(p, stream()), "tiramisu_cuda_free");
}

// ---- halide_buffer_t helpers ------------------------------------------------------------------
bool view_of(const halide_buffer_t *b, BufView *v)
{
    if (!b) return false;
    *v = BufView();
    if (b->host) {
        v->ptr = b->host;
        v->on_device = false;
    } else if (b->device) {
        v->ptr = reinterpret_cast<void *>(b->device);
        v->on_device = true;
    } else {
        return false;
    }
    // Wrappers such as benchmarks/halide/wrapper_heat2d_dist.cpp:52-66 pass a halide_buffer_t with
    // only .host set and garbage elsewhere, so metadata is advisory: use it when it is sane.
    if (b->dimensions > 0 && b->dimensions <= 8 && b->dim != nullptr) {
        v->ndim = b->dimensions;
        for (int i = 0; i < b->dimensions; i++) v->extent[i] = b->dim[i].extent;
        v->elem_bits = b->type.bits;
        v->type_code = b->type.code;
    }
    return true;
}

DeviceArg::~DeviceArg()
{
    if (owned_ && dev_) pool_free(dev_);
}

int DeviceArg::bind(const BufView &v, size_t bytes, bool copy_in)
{
    bytes_ = bytes;
    if (v.on_device) {
        dev_ = v.ptr;
        owned_ = false;
        return 0;
    }
    host_ = v.ptr;
    dev_ = pool_alloc(bytes);
    if (!dev_) return 1;
    owned_ = true;
    if (copy_in) return tiramisu_cuda_memcpy_to_device_async(dev_, host_, bytes);
    return 0;
}

int DeviceArg::copy_out()
{
    if (!owned_) return 0;
    return tiramisu_cuda_memcpy_to_host_async(host_, dev_, bytes_);
}

}  // namespace tb200

using namespace tb200;

// =============================================================================== library control
TB200_API const char *tiramisu_b200_version(void) { return "tiramisu_b200 0.1 (sm_100a)"; }

TB200_API const char *tiramisu_b200_last_error(void)
{
    static thread_local std::string copy;
    std::lock_guard<std::mutex> lk(g_err_mu);
    copy = g_last_error;
    return copy.c_str();
}

TB200_API int tiramisu_b200_set_error_mode(int m) { return g_error_mode.exchange(m ? 1 : 0); }

TB200_API void tiramisu_b200_set_stream(void *s)
{
    t_stream = reinterpret_cast<cudaStream_t>(s);
    t_stream_set = (s != nullptr);
}

TB200_API void *tiramisu_b200_get_stream(void) { return reinterpret_cast<void *>(stream()); }

TB200_API uint64_t tiramisu_b200_kernel_launches(void) { return g_kernel_launches.load(); }

TB200_API int tiramisu_b200_post_launch(const char *kernel_name) { return post_launch(kernel_name ? kernel_name : "generated kernel"); }

// Measurement aid (bench.py): an EMPTY kernel with a given launch shape — grid, block, dynamic shared memory and
// cluster width — so that the period of back-to-back launches of that shape (the launch floor a few-microsecond
// kernel such as the 1024^3 GEMM sits on) can be measured next to the kernel itself.
namespace { __global__ void tiramisu_sm100_null_kernel(int *sink) { if (sink && threadIdx.x == 9999) *sink = 0; } }
TB200_API int tiramisu_b200_null_launch(int32_t grid, int32_t block, int32_t smem_bytes, int32_t cluster)
{
    if (grid < 1 || block < 1 || block > 1024 || smem_bytes < 0 || cluster < 1 || grid % cluster) return fail(__func__, "bad launch shape");
    static int max_set = -1;
    if (smem_bytes > max_set) {
        TB200_CHECK(cudaFuncSetAttribute(tiramisu_sm100_null_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        max_set = smem_bytes;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cfg.stream = stream();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int *none = nullptr;
    TB200_CHECK(cudaLaunchKernelEx(&cfg, tiramisu_sm100_null_kernel, none));
    return post_launch("tiramisu_sm100_null_kernel");
}

TB200_API int tiramisu_b200_pool_trim(void)
{
    int dev = 0;
    cudaMemPool_t pool;
    TB200_CHECK(cudaGetDevice(&dev));
    TB200_CHECK(cudaDeviceGetDefaultMemPool(&pool, dev));
    TB200_CHECK(cudaStreamSynchronize(stream()));
    TB200_CHECK(cudaMemPoolTrimTo(pool, 0));
    return 0;
}

// =============================================================================== B2 shims
TB200_API int tiramisu_cuda_memcpy_to_device_async(void *dst, const void *src, uint64_t bytes)
{
    TB200_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream()));
    return 0;
}

TB200_API int tiramisu_cuda_memcpy_to_host_async(void *dst, const void *src, uint64_t bytes)
{
    TB200_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream()));
    return 0;
}

// Blocking, like the reference's cudaMemcpy (the caller may reuse/read the host buffer on return).
TB200_API int tiramisu_cuda_memcpy_to_device(void *dst, void *src, uint64_t bytes)
{
    TB200_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream()));
    TB200_CHECK(cudaStreamSynchronize(stream()));
    return 0;
}

TB200_API int tiramisu_cuda_memcpy_to_host(void *dst, void *src, uint64_t bytes)
{
    TB200_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream()));
    TB200_CHECK(cudaStreamSynchronize(stream()));
    return 0;
}

TB200_API int tiramisu_cuda_memcpy_to_symbol(void *symbol, void *src, uint64_t bytes)
{
    // `symbol` is what <buf>_get_symbol() returned (src/tiramisu_codegen_cuda.cpp:817-823); our
    // kernels keep small read-only operands in global memory, so a symbol is a device address.
    TB200_CHECK(cudaMemcpyAsync(symbol, src, bytes, cudaMemcpyHostToDevice, stream()));
    TB200_CHECK(cudaStreamSynchronize(stream()));
    return 0;
}

TB200_API void *tiramisu_cuda_malloc(uint64_t bytes) { return pool_alloc(bytes); }

TB200_API int tiramisu_cuda_free(void *p)
{
    if (!p) return 0;
    TB200_CHECK(cudaFreeAsync(p, stream()));
    return 0;
}

TB200_API int32_t tiramisu_cuda_stream_synchronize(int32_t)
{
    TB200_CHECK(cudaStreamSynchronize(stream()));
    return 0;
}

// src/tiramisu_externs.cpp:8-54
#define TB200_ADDRESS_OF(NAME, T)                                                        \
    TB200_API T *tiramisu_address_of_##NAME(halide_buffer_t *b, unsigned long idx)      \
    {                                                                                    \
        return &(reinterpret_cast<T *>(b->host))[idx];                                   \
    }
TB200_ADDRESS_OF(int8, int8_t)
TB200_ADDRESS_OF(uint8, uint8_t)
TB200_ADDRESS_OF(int16, int16_t)
TB200_ADDRESS_OF(uint16, uint16_t)
TB200_ADDRESS_OF(int32, int32_t)
TB200_ADDRESS_OF(uint32, uint32_t)
TB200_ADDRESS_OF(int64, int64_t)
TB200_ADDRESS_OF(uint64, uint64_t)
TB200_ADDRESS_OF(float32, float)
TB200_ADDRESS_OF(float64, double)
