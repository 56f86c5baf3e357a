"""FP64 peak of the box ("FP64 peak is not in MEASURED_PEAKS.json; measure it (cuBLAS DGEMM 8192^3) and record next to hbm_gbs").
Library GEMM used as a yardstick only: nothing in the product calls cuBLAS.  Usage: python tools/fp64_peak.py"""
import json

import torch

n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2):
    torch.matmul(a, b)
torch.cuda.synchronize()
best = 1e9
for _ in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps({"fp64_dgemm_tflops": 2.0 * n ** 3 / (best * 1e-3) / 1e12, "n": n, "ms": best, "how": "torch.matmul float64 8192^3, best of 6, CUDA events",
                  "gpu": torch.cuda.get_device_name(0)}))
