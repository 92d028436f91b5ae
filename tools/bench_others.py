"""Context numbers for the other BASELINE configs (sgemm, conv, LSTM), device-resident, CUDA events.
Imported by bench.py (never the headline); each entry carries its own roofline fraction."""
from __future__ import annotations

import numpy as np


def _time(torch, stream, fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def sgemm(torch, stream, peaks, sizes=(1024, 4096, 8192)):
    from tiramisu_b200 import api
    out = {}
    tf32_peak = float(peaks.get("tf32_tflops", peaks["bf16_tflops"] / 2.0))
    for n in sizes:
        with torch.cuda.stream(stream):
            g = torch.Generator(device="cuda").manual_seed(5)
            A = torch.rand((n, n), device="cuda", generator=g) * 2 - 1
            B = torch.rand((n, n), device="cuda", generator=g) * 2 - 1
            C = torch.zeros((n, n), device="cuda")
        stream.synchronize()
        # L2 hygiene: rotate through enough operand sets to exceed the 126 MB L2 for small n
        nsets = max(1, int(np.ceil(160e6 / (3 * n * n * 4)))) if n <= 2048 else 1
        sets = [(A, B, C)] + [(A.clone(), B.clone(), C.clone()) for _ in range(nsets - 1)]
        idx = [0]

        def fn():
            a, b, c = sets[idx[0] % len(sets)]
            idx[0] += 1
            api.cublas_sgemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), n, n, n, 3.0, 2.0)

        ms = _time(torch, stream, fn, iters=20 if n <= 4096 else 5)
        tfl = 2.0 * n ** 3 / (ms * 1e-3) / 1e12
        out[f"sgemm_{n}"] = {"ms": ms, "TFLOPs": tfl, "GFLOPs": tfl * 1e3, "frac_of_tf32_peak": tfl / tf32_peak,
                             "tf32_peak_TFLOPs": tf32_peak, "operand_sets": nsets}
    return out


def run(torch, stream, peaks):
    res = {}
    for name, fn in (("sgemm", sgemm),):
        try:
            res.update(fn(torch, stream, peaks))
        except Exception as e:
            res[name] = {"error": repr(e)}
    return res
