"""TF32 tensor throughput with the protocol of MEASURED_PEAKS.json (torch.matmul 8192^3, best of 10 = burst;
back to back for 4 s = sustained), BASELINE.md section 2 asks for it before a TF32 utilisation is quoted."""
import json
import sys
import time

import torch

torch.backends.cuda.matmul.allow_tf32 = True
dev = torch.device("cuda:0")
n = 8192
a = torch.randn(n, n, device=dev)
b = torch.randn(n, n, device=dev)
for _ in range(3):
    a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2 * n ** 3 / (best * 1e-3) / 1e12
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); reps = 0
e0.record()
while time.perf_counter() - t0 < 4.0:
    for _ in range(10):
        a @ b
    reps += 10
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sustained = 2 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
out = {"tf32_tflops": burst, "tf32_tflops_sustained": sustained, "how": "torch.matmul fp32 inputs, allow_tf32=True, 8192^3: best of 10 (burst), back to back 4 s (sustained)", "gpu": torch.cuda.get_device_name(0)}
print(json.dumps(out))
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(json.dumps(out, indent=1) + "\n")
