import torch, time
x = torch.empty(256 << 20, dtype=torch.float32).pin_memory()   # 1 GiB
d = torch.empty_like(x, device="cuda")
for _ in range(2): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print("H2D pinned: %.1f GB/s" % (x.numel() * 4 / dt / 1e9))
t = time.perf_counter()
for _ in range(5): x.copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print("D2H pinned: %.1f GB/s" % (x.numel() * 4 / dt / 1e9))
# both directions at once (what finetune() does: features in, predictions out)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
y = torch.empty(128 << 20, dtype=torch.float32).pin_memory()
e = torch.empty_like(y, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2):
        y.copy_(e, non_blocking=True); y.copy_(e, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print("H2D 1 GiB + D2H 1 GiB concurrently: %.1f GB/s each way" % (x.numel() * 4 / dt / 1e9))
