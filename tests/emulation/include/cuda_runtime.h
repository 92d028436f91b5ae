// tests/emulation/include/cuda_runtime.h — TEST INFRASTRUCTURE, never part of the product.
//
// Lets a *generated* source (x.o.b200.cu, after tests/emulation/emulate.py has rewritten its `<<<...>>>` launches)
// be compiled by g++ and executed on the host, one "thread" after the other, so that the lowering's index
// arithmetic, guards, loop structure and host code can be checked without a GPU (the build container has none).
// Kernels that need real SIMT semantics — __syncthreads, shared memory exchange, warp shuffles — cannot be
// emulated this way and are rejected by emulate.py.  The GPU tests (-m gpu) remain the parity gate; this only
// keeps the CPU-side suite sensitive to lowering regressions.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <type_traits>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct tb_emu_idx { unsigned x = 0, y = 0, z = 0; };
inline thread_local tb_emu_idx blockIdx, threadIdx;
inline thread_local dim3 gridDim, blockDim;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

typedef void *cudaStream_t;
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
inline int cudaMemcpyAsync(void *d, const void *s, size_t n, int, cudaStream_t) { std::memcpy(d, s, n); return 0; }
// CUDA graphs (generated host loops are captured once and replayed): the emulation has no streams to capture —
// cudaStreamBeginCapture reports failure and the generated code runs its statements directly
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorNotSupported = 801 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1, cudaStreamCaptureModeRelaxed = 2 };
typedef void *cudaGraph_t;
typedef void *cudaGraphExec_t;
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaErrorNotSupported; }
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t *g) { *g = nullptr; return cudaErrorNotSupported; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t *, cudaGraph_t, unsigned long long) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }

template <typename K, typename... A> inline void tb_emu_launch(K kernel, dim3 grid, dim3 block, A... args)
{
    gridDim = grid;
    blockDim = block;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++)
                for (unsigned tz = 0; tz < block.z; tz++)
                    for (unsigned ty = 0; ty < block.y; ty++)
                        for (unsigned tx = 0; tx < block.x; tx++) {
                            blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
                            threadIdx.x = tx; threadIdx.y = ty; threadIdx.z = tz;
                            kernel(args...);
                        }
}
