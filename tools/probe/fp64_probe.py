# Probe: cuBLAS DGEMM (the FP64 yard-stick) and a copy bandwidth check.
import torch, json, time
dev = torch.device("cuda:0")
res = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    for _ in range(3): c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"dgemm_{n}_tflops"] = 2 * n**3 / best * 1e-9
    print(n, best, "ms", res[f"dgemm_{n}_tflops"], "TFLOP/s")
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device=dev)
a = a @ a.T + n * torch.eye(n, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
for _ in range(2): L = torch.linalg.cholesky(a)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); L = torch.linalg.cholesky(a); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
res["cusolver_potrf_8192_tflops"] = n**3 / 3 / ms * 1e-9
print("potrf 8192", ms, "ms", res["cusolver_potrf_8192_tflops"])
x = torch.empty(1 << 28, dtype=torch.float64, device=dev); y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize()
e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize()
res["copy_gbs"] = 2 * x.numel() * 8 / e0.elapsed_time(e1) * 1e-6
print("copy GB/s", res["copy_gbs"])
json.dump(res, open("gpurun_out/fp64_probe.json", "w"), indent=1)
