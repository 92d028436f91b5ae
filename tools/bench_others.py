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


def cublas_tf32(torch, n, iters):
    """Library reference point (context only, never on the product path): cuBLAS TF32 via torch.matmul."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.rand((n, n), device="cuda") * 2 - 1
        b = torch.rand((n, n), device="cuda") * 2 - 1
        for _ in range(3):
            a @ b
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            a @ b
        e1.record()
        torch.cuda.synchronize()
        return 2.0 * n ** 3 / (e0.elapsed_time(e1) / iters * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def sgemm(torch, stream, peaks, sizes=(1024, 2048, 4096, 8192, 16384)):
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

        ms = _time(torch, stream, fn, iters=20 if n <= 4096 else (5 if n <= 8192 else 3))
        tfl = 2.0 * n ** 3 / (ms * 1e-3) / 1e12
        out[f"sgemm_{n}"] = {"ms": ms, "TFLOPs": tfl, "GFLOPs": tfl * 1e3, "frac_of_tf32_peak": tfl / tf32_peak,
                             "tf32_peak_TFLOPs": tf32_peak, "operand_sets": nsets,
                             "hbm_bound_ms": 16.0 * n * n / (float(peaks["hbm_gbs"]) * 1e9) * 1e3}
        del sets
        try:
            out[f"sgemm_{n}"]["cublas_tf32_TFLOPs_same_run"] = cublas_tf32(torch, n, 20 if n <= 4096 else 3)
        except Exception as e:  # context only
            out[f"sgemm_{n}"]["cublas_tf32_TFLOPs_same_run"] = repr(e)
        torch.cuda.empty_cache()
    return out


def conv(torch, stream, peaks):
    """BASELINE config 3: conv+ReLU+maxpool N=32, 64->64, 3x3, 224x224 (device-resident)."""
    from tiramisu_b200 import _lib
    L = _lib.lib()
    B, N, FIn, FOut = 32, 224, 64, 64
    with torch.cuda.stream(stream):
        g = torch.Generator(device="cuda").manual_seed(3)
        x = torch.rand((B, N + 2, N + 2, FIn), device="cuda", generator=g) * 2 - 1
        w = torch.rand((FOut // 8, 3, 3, FIn, 8), device="cuda", generator=g) * 2 - 1
        b = torch.rand((FOut,), device="cuda", generator=g) * 2 - 1
        o = torch.empty((B, FOut // 8, N // 2, N // 2, 8), device="cuda")
    stream.synchronize()

    def fn():
        _lib.check(L.tiramisu_b200_conv_relu_maxpool(x.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), B, N,
                                                     FIn, FOut))

    ms = _time(torch, stream, fn, iters=10)
    flop = 2.0 * B * N * N * FOut * FIn * 9
    tf32_peak = float(peaks.get("tf32_tflops", peaks["bf16_tflops"] / 2.0))
    tfl = flop / (ms * 1e-3) / 1e12
    return {"conv_relu_maxpool_32x224x224_64to64": {"ms": ms, "TFLOPs": tfl, "GFLOPs": tfl * 1e3,
                                                    "frac_of_tf32_peak": tfl / tf32_peak,
                                                    "bytes_GBps": (x.numel() + o.numel()) * 4 / (ms * 1e-3) / 1e9}}


def lstm(torch, stream, peaks):
    """BASELINE config 4: LSTM hidden=1024, 4 layers, seq=100, batch=64 (device-resident)."""
    from tiramisu_b200 import _lib
    L = _lib.lib()
    F, B, Ly, T = 1024, 64, 4, 100
    with torch.cuda.stream(stream):
        g = torch.Generator(device="cuda").manual_seed(4)
        W = (torch.rand((Ly, 2, 4 * F, F), device="cuda", generator=g) * 2 - 1) / 32
        b = torch.rand((Ly, 4 * F), device="cuda", generator=g) * 2 - 1
        x = torch.rand((T, B, F), device="cuda", generator=g) * 2 - 1
        y = torch.empty_like(x)
    stream.synchronize()

    def fn():
        _lib.check(L.tiramisu_b200_lstm(W.data_ptr(), b.data_ptr(), x.data_ptr(), y.data_ptr(), F, B, Ly, T))

    ms = _time(torch, stream, fn, iters=5, warm=2)
    flop = T * Ly * 2.0 * (2 * B * 4 * F * F)
    tf32_peak = float(peaks.get("tf32_tflops", peaks["bf16_tflops"] / 2.0))
    tfl = flop / (ms * 1e-3) / 1e12
    return {"lstm_F1024_L4_T100_B64": {"ms": ms, "TFLOPs": tfl, "GFLOPs": tfl * 1e3,
                                       "frac_of_tf32_peak": tfl / tf32_peak, "us_per_cell": ms * 1e3 / (T * Ly)}}


def _smi_clock():
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
        mhz, cap = [x.strip() for x in out.split(",")]
        return {"sm_mhz_after": float(mhz), "sw_power_cap_after": cap}
    except Exception:
        return {}


def run(torch, stream, peaks):
    """Latency-bound items first, the power-hungry large GEMMs last, with a pause before each item: a
    long TF32 GEMM leaves the part power-capped (clocks.sm ~900 MHz for a second or so, measured:
    tools/clock_probe.py) and whatever runs next would be timed at half clock."""
    import time
    res = {}
    items = [("lstm", lambda: lstm(torch, stream, peaks)), ("conv", lambda: conv(torch, stream, peaks))]
    items += [(f"sgemm_{n}", (lambda n=n: sgemm(torch, stream, peaks, sizes=(n,)))) for n in (1024, 2048, 4096, 8192, 16384)]
    for name, fn in items:
        time.sleep(1.0)
        try:
            r = fn()
            for v in r.values():
                if isinstance(v, dict):
                    v.update(_smi_clock())
            res.update(r)
        except Exception as e:
            res[name] = {"error": repr(e)}
    return res
