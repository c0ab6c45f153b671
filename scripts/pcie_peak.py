import torch, time
x = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for _ in range(2): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print("H2D GB/s", (1 << 30) / dt / 1e9)
t0 = time.perf_counter()
for _ in range(5): x.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print("D2H GB/s", (1 << 30) / dt / 1e9)
