#!/usr/bin/env python3
"""Measures the kind::tf32 dense peak that `roofline.frac_of_tf32` divides by (run on the GPU box):

  1. cuBLAS TF32 GEMM through torch (fp32 tensors, allow_tf32), 8192^3, best of 10 (burst) and back to back for 4 s
     (sustained) - the same procedure MEASURED_PEAKS.json uses for bf16;
  2. tests/cuda/tf32_peak: the raw tcgen05.mma.cta_group::2.kind::tf32 issue rate of all 74 SM pairs for 3 s, operands
     resident in shared memory (no loads), dense and in the 3xTF32 K-block pattern of the step program's GEMM engine.

Writes gpurun_out/r2_tf32_peak.json; the committed copy is profiles/r2_tf32_peak.json."""
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def clocks():
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=10).stdout.strip().split("\n")[0]
        a = [float(x) for x in out.split(",")]
        return dict(sm_mhz=a[0], sm_max_mhz=a[1], power_w=a[2])
    except Exception:
        return {}


def cublas_tf32():
    torch.backends.cuda.matmul.allow_tf32 = True
    n = 8192
    a = torch.randn(n, n, device="cuda")
    b = torch.randn(n, n, device="cuda")
    c = torch.empty(n, n, device="cuda")
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = 2 * n ** 3 / (best * 1e-3) / 1e12
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters, t0, mid = 0, time.time(), {}
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(20):
            torch.matmul(a, b, out=c)
        iters += 20
        if not mid and time.time() - t0 > 2.0:
            mid = clocks()
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    sustained = 2 * n ** 3 * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12
    return dict(burst_tflops=burst, sustained_tflops=sustained, clocks_under_load=mid)


def main():
    res = {"gpu": torch.cuda.get_device_name(0), "cublas_tf32_8192": cublas_tf32()}
    exe = os.path.join(ROOT, "tests", "cuda", "tf32_peak")
    if os.path.exists(exe):
        out = subprocess.run(["timeout", "120", exe, "3"], capture_output=True, text=True).stdout
        res["tcgen05_issue_rate"] = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
        res["tcgen05_raw_output"] = out
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r2_tf32_peak.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    sys.exit(main())
