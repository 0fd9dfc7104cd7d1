"""Measure the dense TF32 matmul peak of this GPU the way MEASURED_PEAKS.json measured bf16:
torch.matmul 8192^3 with allow_tf32 (cuBLAS), best of 10 (burst) and back to back for a few
seconds (sustained).  Prints one JSON line.  ("measure with a TF32 8192^3 matmul".)"""
import json
import sys
import time

import torch


def measure(seconds=3.0, n=8192):
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.randn(n, n, device="cuda")
    b = torch.randn(n, n, device="cuda")
    c = torch.empty(n, n, device="cuda")
    fl = 2.0 * n ** 3
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = fl / (best * 1e-3) / 1e12
    # sustained: back-to-back launches for `seconds`
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 0
    t0 = time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(20):
            torch.matmul(a, b, out=c)
        iters += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sustained = fl * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12
    return {"tf32_tflops": burst, "tf32_tflops_sustained": sustained, "n": n, "seconds": seconds,
            "how": "torch.matmul fp32 inputs, allow_tf32=True (cuBLAS), 8192^3: best of 10 (burst), back to back (sustained)"}


if __name__ == "__main__":
    print(json.dumps(measure(float(sys.argv[1]) if len(sys.argv) > 1 else 3.0)))
