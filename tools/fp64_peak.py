"""Measure the cuBLAS DGEMM ceiling on this B200 (FP64 roofline denominator).
Run on the GPU box: python tools/fp64_peak.py  -> prints JSON and writes gpurun_out/fp64_peak.json"""
import json, time, torch
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
res = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    for _ in range(3):
        c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"dgemm_{n}_burst_tflops"] = 2 * n**3 / best * 1e-9
    # sustained ~3 s
    t0 = time.time(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); cnt = 0
    while time.time() - t0 < 3.0:
        for _ in range(4):
            c = a @ b; cnt += 1
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    res[f"dgemm_{n}_sustained_tflops"] = 2 * n**3 * cnt / e0.elapsed_time(e1) * 1e-9
# tall-skinny shape like the weighted Gram product (5050 x 5000 x 20000 is ~1e12 flop)
M, K, N = 5050, 5000, 20000
a = torch.randn(M, K, dtype=torch.float64, device=dev); b = torch.randn(K, N, dtype=torch.float64, device=dev)
for _ in range(2): c = a @ b
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
res["dgemm_gram_shape_tflops"] = 2 * M * K * N / e0.elapsed_time(e1) * 1e-9
print(json.dumps(res))
import os
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/fp64_peak.json", "w"), indent=1)
