// common.cuh — shared plumbing of the sm_100a backend (streams, errors, launch accounting).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "tiramisu_b200/halide_buffer.h"

#define TB200_API extern "C" __attribute__((visibility("default")))

namespace tb200 {

// ---- error handling -------------------------------------------------------------------------
// The reference's runtime shims print and exit(1) on any CUDA failure
// (src/tiramisu_cuda_wrappers.cpp:14-31).  We keep that as the default for the tiramisu_cuda_*
// symbols and additionally record the message; tiramisu_b200_set_error_mode(1) turns the exit
// into a non-zero return so that a host language can raise instead of dying.
void set_last_error(const std::string &msg);
int fail(const char *where, const std::string &msg);          // records, maybe exits, returns 1
int check(cudaError_t e, const char *where);                  // 0 on success

#define TB200_CHECK(expr)                                                   \
    do {                                                                    \
        if (int _rc = ::tb200::check((expr), __func__)) return _rc;         \
    } while (0)

// ---- stream ---------------------------------------------------------------------------------
// One launch stream per host thread (the reference compiles with --default-stream per-thread,
// src/tiramisu_codegen_cuda.cpp:1745).  tiramisu_b200_set_stream() overrides it for the calling
// thread, e.g. with torch's current stream.
cudaStream_t stream();
// Redirects stream() of the calling thread for a scope (fork/join onto auxiliary streams).
class ScopedStream {
public:
    explicit ScopedStream(cudaStream_t s);
    ~ScopedStream();
    ScopedStream(const ScopedStream &) = delete;
    ScopedStream &operator=(const ScopedStream &) = delete;
private:
    cudaStream_t prev_;
    bool prev_set_;
};

// ---- launch accounting ----------------------------------------------------------------------
extern std::atomic<uint64_t> g_kernel_launches;
inline void count_launch(int n = 1) { g_kernel_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
int post_launch(const char *kernel_name);                     // cudaGetLastError + count

int sm_count();

// ---- halide_buffer_t views ------------------------------------------------------------------
// A generated function receives halide_buffer_t*.  The reference reads only ->host (sizes are
// baked in at generator time).  We accept either a host pointer (->host) or, when ->host is
// null and ->device is not, a raw CUDA device pointer in ->device (device-resident call: no
// copies).  Extents come from ->dim when present (innermost first), else from the caller's
// defaults (the reference's compile-time sizes).
struct BufView {
    void *ptr = nullptr;
    bool on_device = false;
    int ndim = 0;
    int64_t extent[8] = {0};
    int elem_bits = 0;          // 0 when the caller left type zeroed
    int type_code = -1;
};

bool view_of(const halide_buffer_t *b, BufView *v);           // false if b is null / has no pointer

// Scoped device mirror of an argument buffer: allocates from the stream-ordered pool and copies
// host->device when the argument lives on the host; no-op for device-resident arguments.
class DeviceArg {
public:
    DeviceArg() = default;
    ~DeviceArg();
    DeviceArg(const DeviceArg &) = delete;
    DeviceArg &operator=(const DeviceArg &) = delete;
    int bind(const BufView &v, size_t bytes, bool copy_in);   // 0 on success
    int copy_out();                                           // device->host if host-resident (async)
    void *dev() const { return dev_; }
    template <class T> T *as() const { return reinterpret_cast<T *>(dev_); }
    bool staged() const { return owned_; }
private:
    void *dev_ = nullptr;
    void *host_ = nullptr;
    size_t bytes_ = 0;
    bool owned_ = false;
};

void *pool_alloc(size_t bytes);                               // stream-ordered, pooled; exits/returns null on failure
void pool_free(void *p);

}  // namespace tb200
