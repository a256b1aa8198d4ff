#!/usr/bin/env python
"""Measured FP64 / FP32 / TF32 GEMM throughput of this GPU through cuBLAS (torch.matmul), best of 10 launches with
CUDA events: the denominators the Cholesky half-sweep's FLOP roofline is reported against (SURVEY.md 7 asked for the
measured number instead of the datasheet one).  Writes gpurun_out/fp_peaks.json."""
import json
import os
import sys

import torch


def best_tflops(dtype, n, allow_tf32=False, reps=10):
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    for _ in range(3):
        a @ b
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def main():
    out = dict(gpu=torch.cuda.get_device_name(0),
               fp64_tflops=best_tflops(torch.float64, 8192),
               fp32_tflops=best_tflops(torch.float32, 8192, False),
               tf32_tflops=best_tflops(torch.float32, 8192, True),
               how="torch.matmul 8192^3 (2*N^3 flop), best of 10, CUDA events")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/fp_peaks.json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    sys.exit(main())
