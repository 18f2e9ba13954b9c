# cuBLAS DGEMM yard-stick (torch.matmul f64) + HBM copy, printed as JSON lines.
import json, torch
assert torch.cuda.is_available()
def t(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for n in (1024, 2048, 4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    ms = t(lambda: torch.matmul(a, b))
    print(json.dumps({"dgemm_n": n, "ms": ms, "tflops": 2 * n**3 / ms * 1e-9}))
# the matvec stage shapes at chi=256, w=3
for (m, k, n) in ((768, 256, 1024), (1024, 768, 256), (1536, 512, 2048), (2048, 1536, 512)):
    a = torch.randn(m, k, dtype=torch.float64, device="cuda"); b = torch.randn(k, n, dtype=torch.float64, device="cuda")
    ms = t(lambda: torch.matmul(a, b), reps=20)
    print(json.dumps({"dgemm_mkn": [m, k, n], "ms": ms, "tflops": 2 * m * k * n / ms * 1e-9}))
a = torch.randn(512, 512, dtype=torch.float64, device="cuda")
ms = t(lambda: torch.linalg.svd(a), reps=3)
print(json.dumps({"cusolver_svd_512_ms": ms}))
a = torch.randn(1024, 1024, dtype=torch.float64, device="cuda")
ms = t(lambda: torch.linalg.svd(a), reps=3)
print(json.dumps({"cusolver_svd_1024_ms": ms}))
