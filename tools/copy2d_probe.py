"""Peer-copy probe (2 GPUs, one process): bandwidth of cudaMemcpyAsync / cudaMemcpy2DAsync from GPU 1 into GPU 0
as a function of the row width, with and without a persistent GEMM running on GPU 0 — decides how the B panels of
the row-sharded sgemm travel (tools/runs/r02_run8.sh).  Prints one JSON line per case."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiramisu_b200 import api  # noqa: E402


def main():
    rt = C.CDLL("libcudart.so.12")
    n = 16384
    with torch.cuda.device(1):
        src = torch.rand((n, n), device="cuda:1")
    torch.cuda.set_device(0)
    dst = torch.empty((n, n), device="cuda:0")
    rt.cudaDeviceEnablePeerAccess(1, 0)
    s = torch.cuda.Stream()
    st = C.c_void_p(s.cuda_stream)
    D2D = 3
    rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]

    def timed(fn, reps=5):
        fn(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            fn()
        e1.record(s)
        s.synchronize()
        return e0.elapsed_time(e1) / reps

    A = torch.rand((4096, 8192), device="cuda:0")
    Bm = torch.rand((8192, 8192), device="cuda:0")
    Cm = torch.zeros((4096, 8192), device="cuda:0")
    gs = torch.cuda.Stream()

    def with_gemm(flag):
        if flag:
            api.set_stream(gs.cuda_stream)
            for _ in range(12):
                api.cublas_sgemm(A.data_ptr(), Bm.data_ptr(), Cm.data_ptr(), 4096, 8192, 8192, 1.0, 0.0)

    for busy in (False, True):
        for cols in (n, 8192, 4096, 2048, 1024, 512):
            rows = n if cols < n else n // 4
            nbytes = rows * cols * 4

            def fn():
                if cols == n:
                    rc = rt.cudaMemcpyAsync(dst.data_ptr(), src.data_ptr(), nbytes, D2D, st)
                else:
                    rc = rt.cudaMemcpy2DAsync(dst.data_ptr(), n * 4, src.data_ptr(), n * 4, cols * 4, rows, D2D, st)
                assert rc == 0, rc
            with_gemm(busy)
            ms = timed(fn)
            gs.synchronize()
            print(json.dumps({"peer_copy": "1D" if cols == n else "2D", "row_bytes": cols * 4, "rows": rows, "pitch": n * 4,
                              "gemm_running_on_dst": busy, "ms": ms, "GBps": nbytes / ms / 1e6}), flush=True)
    ok = bool(torch.equal(dst[:, :512].cpu(), src[:, :512].cpu()))
    print(json.dumps({"last_2d_copy_bit_exact": ok}))

    # K-panel GEMM cost on one GPU: 4096 x 8192 x K with beta = 1 against one launch over the whole K
    api.set_stream(s.cuda_stream)
    for K, beta in ((8192, 0.0), (2048, 1.0), (2048, 0.0), (4096, 1.0)):
        ms = timed(lambda: api.cublas_sgemm(A.data_ptr(), Bm.data_ptr(), Cm.data_ptr(), 4096, 8192, K, 1.0, beta, ldA=8192, ldB=8192, ldC=8192), 10)
        print(json.dumps({"gemm": f"4096x8192x{K}", "beta": beta, "ms": ms, "TFLOPs": 2 * 4096 * 8192 * K / ms / 1e9}), flush=True)


if __name__ == "__main__":
    main()
