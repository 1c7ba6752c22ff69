import torch, time
n = 1560_000_000 // 8
d = torch.empty(n, dtype=torch.float64, device='cuda'); h = torch.empty(n, dtype=torch.float64).pin_memory()
d2 = torch.empty(816_000_000 // 8, dtype=torch.float64, device='cuda'); h2 = torch.empty(816_000_000 // 8, dtype=torch.float64).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for _ in range(2): h.copy_(d, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): h.copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print("D2H alone GB/s", 1.56 / dt)
t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print("D2H 1.56 GB + H2D 0.816 GB concurrently: ms", dt * 1e3, "D2H GB/s", 1.56 / dt)
