"""Run one workload a few times (for ncu captures).  usage: python tools/run_one.py sgemm4096|sgemm1024|conv|lstm|heat2d|blur"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiramisu_b200 import _lib, api  # noqa: E402
from tools import bench_others  # noqa: E402


def main():
    what = sys.argv[1]
    stream = torch.cuda.Stream()
    api.set_stream(stream.cuda_stream)
    peaks = {"bf16_tflops": 1670.0, "hbm_gbs": 6550.0}
    if what.startswith("sgemm"):
        print(bench_others.sgemm(torch, stream, peaks, sizes=(int(what[5:]),)))
    elif what == "conv":
        print(bench_others.conv(torch, stream, peaks))
    elif what == "lstm":
        print(bench_others.lstm(torch, stream, peaks))
    elif what == "tf32peak":
        import json
        torch.backends.cuda.matmul.allow_tf32 = True
        n = 8192
        a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
        for _ in range(3):
            a @ b
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res = {"tf32_tflops": 2.0 * n ** 3 / (best * 1e-3) / 1e12, "how": "torch.matmul fp32 with allow_tf32, 8192^3, best of 10 (cuBLAS)"}
        torch.backends.cuda.matmul.allow_tf32 = False
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res["fp32_simt_tflops"] = 2.0 * n ** 3 / (best * 1e-3) / 1e12
        print(json.dumps(res))
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(res, open("gpurun_out/tf32_peak.json", "w"))
    else:
        L = _lib.lib()
        N = 16384
        x = torch.rand((N, N), device="cuda")
        o = torch.empty_like(x)
        for _ in range(5):
            if what == "heat2d":
                L.tiramisu_b200_heat2d(x.data_ptr(), o.data_ptr(), N, N, 0, 0, N)
            else:
                L.tiramisu_b200_blur_f32(x.data_ptr(), o.data_ptr(), N, N, 0, N - 2)
        api.synchronize()


if __name__ == "__main__":
    main()
