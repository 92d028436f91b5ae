"""GPU diagnostic for the tcgen05 GEMM: error statistics per (BN, layout, shape) so that one gpurun
round-trip tells which descriptor / layout is wrong.  Writes gpurun_out/gemm_diag.txt."""
import os
import sys
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiramisu_b200 import api  # noqa: E402


def run_case(M, N, K, tB, bn, kind):
    os.environ["TB200_GEMM_BN"] = str(bn)
    rng = np.random.default_rng(0)
    if kind == "ints":
        A = rng.integers(-5, 5, (M, K)).astype(np.float32)
        B = rng.integers(-5, 5, (N, K) if tB else (K, N)).astype(np.float32)
    elif kind == "eye":
        A = np.zeros((M, K), np.float32)
        A[np.arange(min(M, K)), np.arange(min(M, K))] = 1
        B = (np.arange(K * N, dtype=np.float32).reshape((K, N)) % 997)
        if tB:
            B = np.ascontiguousarray(B.T)
    else:
        A = rng.uniform(-1, 1, (M, K)).astype(np.float32)
        B = rng.uniform(-1, 1, (N, K) if tB else (K, N)).astype(np.float32)
    want = A.astype(np.float64) @ (B.T if tB else B).astype(np.float64)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    dC = torch.full((M, N), float("nan"), device="cuda")
    api.cublas_sgemm(dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), M, N, K, 1.0, 0.0, tB=tB)
    api.synchronize()
    got = dC.cpu().numpy().astype(np.float64)
    err = np.abs(got - want)
    rel = err.max() / max(1e-30, np.abs(want).max())
    bad = np.argwhere(~(err <= 1e-2 * max(1.0, np.abs(want).max())))
    return rel, len(bad), (bad[:4].tolist() if len(bad) else [])


def main():
    api.set_gemm_mode(2)
    os.makedirs("gpurun_out", exist_ok=True)
    lines = []
    # MN-major B descriptor bring-up: which (leading, stride) byte offsets does the hardware want?
    for lbo, sbo in ((4096, 512), (512, 4096), (4096, 1024), (1024, 4096), (128, 512), (512, 128)):
        os.environ["TB200_DBG_B_LBO"], os.environ["TB200_DBG_B_SBO"] = str(lbo), str(sbo)
        for (M, N, K) in ((128, 256, 32), (256, 512, 256)):
            rel, nbad, where = run_case(M, N, K, False, 128, "ints")
            lines.append(f"MN-major probe lbo={lbo} sbo={sbo} M={M} N={N} K={K} rel={rel:.3e} bad={nbad} {where}")
    del os.environ["TB200_DBG_B_LBO"], os.environ["TB200_DBG_B_SBO"]
    for kind in ("ints", "rand"):
        for tB in (False, True):
            for bn in (64, 128, 256):
                for (M, N, K) in ((128, 256, 32), (128, 256, 64), (256, 512, 256), (136, 264, 100), (1024, 1024, 1024)):
                    try:
                        rel, nbad, where = run_case(M, N, K, tB, bn, kind)
                        lines.append(f"{kind:5s} tB={int(tB)} BN={bn:3d} M={M:5d} N={N:5d} K={K:5d} rel={rel:.3e} bad={nbad} {where}")
                    except Exception:
                        lines.append(f"{kind:5s} tB={int(tB)} BN={bn:3d} M={M} N={N} K={K} EXC {traceback.format_exc(limit=1)!r}")
                        open("gpurun_out/gemm_diag.txt", "w").write("\n".join(lines) + "\n")
                        raise
    open("gpurun_out/gemm_diag.txt", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
