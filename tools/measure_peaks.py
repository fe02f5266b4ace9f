"""Measures the INT8 tensor peak of this B200 the same way MEASURED_PEAKS.json measures bf16
(cuBLASLt GEMM 8192^3 through torch, best of 10 = burst, back-to-back for 4 s = sustained), plus
the bf16 figure again for reference.  BASELINE.md §2: "FP64 tensor / INT8 tensor: not measured —
must be measured before being used as a denominator".  Writes JSON to stdout."""
import json
import time

import torch


def bench(fn, flop):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flop / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return best, flop * n / (e0.elapsed_time(e1) * 1e-3) / 1e12


def main():
    n = 8192
    out = {"gpu_name": torch.cuda.get_device_name(0), "torch": torch.__version__, "n": n,
           "how": "torch._int_mm int8 8192^3 (2*N^3 ops) and torch.matmul bf16 8192^3: best of 10 (burst), back to back 4 s (sustained), CUDA events"}
    a = torch.randn(n, n, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(n, n, device="cuda", dtype=torch.bfloat16)
    out["bf16_tflops"], out["bf16_tflops_sustained"] = bench(lambda: torch.matmul(a, b), 2.0 * n ** 3)
    ai = torch.randint(-128, 127, (n, n), device="cuda", dtype=torch.int8)
    bi = torch.randint(-128, 127, (n, n), device="cuda", dtype=torch.int8).t().contiguous().t()
    try:
        out["int8_tops"], out["int8_tops_sustained"] = bench(lambda: torch._int_mm(ai, bi), 2.0 * n ** 3)
    except Exception as e:  # noqa
        out["int8_error"] = str(e)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
