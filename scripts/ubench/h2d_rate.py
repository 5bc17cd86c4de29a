import torch, time
for mb in (12.8, 25.6, 128, 256, 1024):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    reps = max(4, int(2e9 / n))
    for _ in range(reps): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("H2D %7.1f MB: %.1f GB/s" % (mb, reps * n / dt / 1e9))
# two streams concurrently
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
n = int(12.8e6)
hs = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(2)]; ds = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(2)]
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(100):
    with torch.cuda.stream(s1): ds[0].copy_(hs[0], non_blocking=True)
    with torch.cuda.stream(s2): ds[1].copy_(hs[1], non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("H2D 2 streams x 12.8 MB: %.1f GB/s" % (200 * n / dt / 1e9))
