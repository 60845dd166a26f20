"""Dense TF32 tensor-pipe peak of this B200 as cuBLAS reaches it (fp32 inputs, TF32 allowed, 8192^3): burst (best of
10, like MEASURED_PEAKS.json's bf16 figure) and sustained (back to back for 4 s under the power cap).  Written to
gpurun_out/tf32_peak.json; the committed copy, profiles/tf32_peak.json, is the denominator of the FLAT roofline in
bench.py.   Run:  gpurun -- python profiles/measure_tf32_peak.py"""
import json
import os
import time

import torch

torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda")
b = torch.randn(n, n, device="cuda")
for _ in range(3):
    a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    a @ b
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2.0 * n ** 3 / (best / 1e3) / 1e12
t0 = time.time()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
it = 0
while time.time() - t0 < 4.0:
    for _ in range(20):
        a @ b
    it += 20
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
sustained = 2.0 * n ** 3 * it / (e0.elapsed_time(e1) / 1e3) / 1e12
# the same in bf16, to tie the figure to MEASURED_PEAKS.json on this very box
ah, bh = a.bfloat16(), b.bfloat16()
for _ in range(3):
    ah @ bh
torch.cuda.synchronize()
bb = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ah @ bh
    e1.record()
    torch.cuda.synchronize()
    bb = min(bb, e0.elapsed_time(e1))
out = {"tf32_tflops": round(burst, 1), "tf32_tflops_sustained": round(sustained, 1),
       "bf16_tflops_same_box": round(2.0 * n ** 3 / (bb / 1e3) / 1e12, 1), "gpu": torch.cuda.get_device_name(0),
       "how": "torch.matmul fp32 8192^3 with allow_tf32 (cuBLAS TF32): best of 10 (burst), back to back for 4 s (sustained)"}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/tf32_peak.json", "w"), indent=1)
print(json.dumps(out))
