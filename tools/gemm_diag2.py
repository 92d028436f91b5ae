"""GPU diagnostic for the tcgen05 GEMM kernels: correctness (integer operands are exact in TF32) and
device time per kernel variant.  usage: python tools/gemm_diag2.py <kernel 1|2|3> [check|time|all]
Run each variant in its own process (a device-side trap poisons the CUDA context)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiramisu_b200 import api  # noqa: E402


def check(kernel, M, N, K, tB, alpha=1.0, beta=0.0):
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.integers(-5, 5, (M, K)).astype(np.float32)
    B = rng.integers(-5, 5, (N, K) if tB else (K, N)).astype(np.float32)
    C0 = rng.integers(-5, 5, (M, N)).astype(np.float32)
    want = alpha * (A.astype(np.float64) @ (B.T if tB else B).astype(np.float64)) + beta * C0
    dA, dB, dC = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), torch.from_numpy(C0).cuda()
    api.cublas_sgemm(dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), M, N, K, alpha, beta, tB=tB)
    api.synchronize()
    got = dC.cpu().numpy().astype(np.float64)
    bad = np.argwhere(got != want)
    return len(bad), (bad[:3].tolist() if len(bad) else [])


def timeit(n, iters, tB=False, sets=1):
    g = torch.Generator(device="cuda").manual_seed(5)
    ops = []
    for _ in range(sets):
        A = torch.rand((n, n), device="cuda", generator=g) * 2 - 1
        B = torch.rand((n, n), device="cuda", generator=g) * 2 - 1
        C = torch.zeros((n, n), device="cuda")
        ops.append((A, B, C))
    i = [0]

    def fn():
        a, b, c = ops[i[0] % sets]
        i[0] += 1
        api.cublas_sgemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), n, n, n, 3.0, 2.0, tB=tB)

    for _ in range(3):
        fn()
    api.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return best


def main():
    kernel = int(sys.argv[1])
    what = sys.argv[2] if len(sys.argv) > 2 else "all"
    api.set_stream(torch.cuda.current_stream().cuda_stream)
    api.set_gemm_mode(2)
    api.set_gemm_kernel(kernel)
    out = {"kernel": kernel}
    if what in ("check", "all"):
        shapes = [(256, 256, 32), (256, 256, 64), (256, 512, 256), (512, 256, 1024), (136, 264, 100),
                  (1024, 1024, 1024), (300, 700, 96), (2048, 2048, 512), (4096, 4096, 64)]
        for tB in (False, True):
            for (M, N, K) in shapes:
                nbad, where = check(kernel, M, N, K, tB)
                print(f"k{kernel} tB={int(tB)} M={M} N={N} K={K} bad={nbad} {where}", flush=True)
                out[f"chk_tB{int(tB)}_{M}x{N}x{K}"] = nbad
        nbad, where = check(kernel, 512, 512, 256, False, 3.0, 2.0)
        print(f"k{kernel} alpha/beta bad={nbad} {where}", flush=True)
        out["chk_alpha_beta"] = nbad
    if what in ("time", "all"):
        for n, iters, sets in ((1024, 50, 13), (2048, 20, 4), (4096, 10, 1), (8192, 5, 1)):
            for tB in (False, True):
                ms = timeit(n, iters, tB, sets)
                tf = 2.0 * n ** 3 / (ms * 1e-3) / 1e12
                print(f"k{kernel} tB={int(tB)} n={n} {ms * 1e3:.1f} us {tf:.1f} TFLOP/s", flush=True)
                out[f"time_tB{int(tB)}_{n}"] = {"us": ms * 1e3, "TFLOPs": tf}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/gemm_diag2_k{kernel}_{what}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
