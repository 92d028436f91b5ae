"""Run the lowered `*gpu` Halide benchmarks (tests/compat/_bin/libgen_halide_*.so) once each at the size of the
reference's utils/images/rgb.png (2112 x 3520 x 3; edge: its fixed 8192 x 8192 x 3; heat2dgpu 10000 x 10000) — the
workload for an `ncu --metrics gpu__time_duration.sum` launch list (kernel time vs. algorithmic bytes):

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/halide_gpu_run.py
    python tools/halide_gpu_run.py --report out.csv        -> GB/s per kernel against the algorithmic bytes
"""
import csv
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

N, M = 2112, 3520
# algorithmic bytes per call (inputs read once + outputs written once), keyed by a substring of the kernel name
ALG = {
    "cvtcolorgpu": 3 * M * N + M * N,
    "fusiongpu": 3 * M * N + 4 * 3 * M * N,
    "rgbyuv420gpu_tiramisu_k0": 3 * M * N + M * N,                       # y: reads all three planes
    "rgbyuv420gpu_tiramisu_k1": 3 * (M // 2) * (N // 2) + 2 * (M // 2) * (N // 2),   # u, v: the even pixels
    "gaussiangpu": 3 * M * N + 3 * (M - 5) * (N - 5),
    "convolutiongpu": 3 * M * N + 3 * (M - 2) * (N - 2),
    "edge_tiramisu_k0": 2 * 3 * 8192 * 8192,
    "edge_tiramisu_k1": 2 * 3 * 8192 * 8192,
    "heat2dgpu_tiramisu_k2": 2 * 10000 * 10000 * 4,
}


def run():
    import ctypes as C
    from tests import test_frontend_halide_gpu as t
    from oracle import oracle  # noqa: F401
    t.check_cvtcolor(t._lib_of("cvtcolorgpu"), M, N)
    t.check_fusion(t._lib_of("fusiongpu"), M, N)
    t.check_rgbyuv420(t._lib_of("rgbyuv420gpu"), M, N)
    t.check_gaussian(t._lib_of("gaussiangpu"), M, N)
    t.check_convolution(t._lib_of("convolutiongpu"), M, N)
    img = np.random.default_rng(6).integers(0, 256, (3, 8192, 8192), dtype=np.uint8)
    t._call(t._lib_of("edgegpu"), "edge_tiramisu", img)
    a = np.random.default_rng(8).random((10000, 10000), dtype=np.float32)
    out = np.empty_like(a)
    t._call(t._lib_of("heat2dgpu"), "heat2dgpu_tiramisu", np.array([10000, 10000], np.int32), a, out)
    print("ran all")


def report(path, peak_gbs):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5 and r[0].isdigit()]
    out = {}
    for r in rows:
        name, ns = r[4], float(r[-1])
        for key, b in ALG.items():
            if key in name:
                out[name] = {"us": ns / 1e3, "algorithmic_bytes": b, "GBps": b / ns, "frac_of_hbm_peak": b / ns / peak_gbs}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--report":
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        report(sys.argv[2], float(peaks["hbm_gbs"]))
    else:
        run()
