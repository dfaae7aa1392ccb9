import torch, time
x = torch.empty(1_360_000_000 // 8, dtype=torch.float64).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(2): d.copy_(x, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(f"H2D pinned 1.36 GB: {dt*1e3:.1f} ms = {x.numel()*8/dt/1e9:.1f} GB/s")
h = torch.empty_like(x).pin_memory()
t0 = time.perf_counter()
for _ in range(5): h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(f"D2H pinned 1.36 GB: {dt*1e3:.1f} ms = {x.numel()*8/dt/1e9:.1f} GB/s")
