"""Write-only HBM bandwidth on this GPU (torch fill_ / zero_ of 16 GB), the practical ceiling of the
observation kernel, next to the copy figure of MEASURED_PEAKS.json."""
import torch
n = 4 * 1024 ** 3  # 4 Gi floats = 16 GiB
x = torch.empty(n, dtype=torch.float32, device="cuda")
for name, fn in (("zero_", lambda: x.zero_()), ("fill_", lambda: x.fill_(1.5))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("%s: %.2f ms for %.1f GB -> %.0f GB/s" % (name, best, n * 4 / 1e9, n * 4 / best / 1e6))
y = torch.empty(n // 2, dtype=torch.float32, device="cuda")
z = torch.empty(n // 2, dtype=torch.float32, device="cuda")
for _ in range(2):
    z.copy_(y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(5):
    e0.record(); z.copy_(y); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("copy_: %.2f ms, read+write %.0f GB/s" % (best, 2 * (n // 2) * 4 / best / 1e6))
