import torch, time
n = 5_300_000_000
x = torch.empty(n, dtype=torch.uint8, device="cuda")
y = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
z = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
def t(f, k=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(k):
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.zero_())
print("write-only (cudaMemset-style zero_) %.1f GB in %.3f ms = %.0f GB/s" % (n/1e9, ms, n/ms/1e6))
xi = x.view(torch.int32)
ms = t(lambda: xi.fill_(7))
print("write-only (fill_ int32 kernel)      %.1f GB in %.3f ms = %.0f GB/s" % (n/1e9, ms, n/ms/1e6))
ms = t(lambda: z.copy_(y))
print("copy (read+write)                     %.1f GB moved in %.3f ms = %.0f GB/s" % (n/1e9, ms, n/ms/1e6))
