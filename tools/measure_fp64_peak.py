"""Measures the FP64 / ComplexF64 GEMM peak of this GPU with cuBLAS (through torch.matmul): the roofline denominator
for every "% of FP64 tensor peak" quoted by bench.py (MEASURED_PEAKS.json has no FP64 figure).
Writes gpurun_out/fp64_peak.json."""
import json
import os
import time

import torch


def run(dtype, n, burst_iters=10, sustain_s=3.0):
    a = torch.randn(n, n, dtype=dtype, device="cuda")
    b = torch.randn(n, n, dtype=dtype, device="cuda")
    flops = (8.0 if dtype.is_complex else 2.0) * n ** 3
    for _ in range(3):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(burst_iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    it = 0
    while time.time() - t0 < sustain_s:
        for _ in range(5):
            torch.matmul(a, b)
        it += 5
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sustained = flops * it / (e0.elapsed_time(e1) * 1e-3)
    return flops / best / 1e12, sustained / 1e12


def main():
    out = {"gpu": torch.cuda.get_device_name(0), "how": "torch.matmul (cuBLAS) n=8192 (f64) / n=4096 (c128): best of 10 (burst) and a 3 s loop (sustained)"}
    b, s = run(torch.float64, 8192)
    out["fp64_tflops"], out["fp64_tflops_sustained"] = round(b, 2), round(s, 2)
    b, s = run(torch.complex128, 4096)
    out["c128_tflops"], out["c128_tflops_sustained"] = round(b, 2), round(s, 2)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/fp64_peak.json", "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
