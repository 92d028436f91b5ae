"""Does a long TF32 GEMM leave the GPU in a slow state for the kernels that follow?"""
import os, subprocess, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiramisu_b200 import api
from tools import bench_others

def smi():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader"],
                          capture_output=True, text=True).stdout.strip()

stream = torch.cuda.Stream(); api.set_stream(stream.cuda_stream)
peaks = {"bf16_tflops": 1636.3, "hbm_gbs": 6545.6}
print("idle", smi())
print("conv A", bench_others.conv(torch, stream, peaks)["conv_relu_maxpool_32x224x224_64to64"]["ms"], smi())
print("lstm A", bench_others.lstm(torch, stream, peaks)["lstm_F1024_L4_T100_B64"]["ms"], smi())
r = bench_others.sgemm(torch, stream, peaks, sizes=(16384,)); print("sgemm16384", r["sgemm_16384"]["TFLOPs"], smi())
print("conv B", bench_others.conv(torch, stream, peaks)["conv_relu_maxpool_32x224x224_64to64"]["ms"], smi())
print("lstm B", bench_others.lstm(torch, stream, peaks)["lstm_F1024_L4_T100_B64"]["ms"], smi())
time.sleep(2)
print("conv C", bench_others.conv(torch, stream, peaks)["conv_relu_maxpool_32x224x224_64to64"]["ms"], smi())
torch.cuda.empty_cache()
print("conv D", bench_others.conv(torch, stream, peaks)["conv_relu_maxpool_32x224x224_64to64"]["ms"], smi())
print("lstm D", bench_others.lstm(torch, stream, peaks)["lstm_F1024_L4_T100_B64"]["ms"], smi())
