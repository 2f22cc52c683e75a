import torch, time
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for _ in range(2): h.copy_(d, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
print("D2H 1-D pinned GB/s", 5 * n / (time.perf_counter() - t0) / 1e9)
# 2-D: 4096 rows x 128 KiB out of a pitched host buffer
hp = torch.empty(4096, 1 << 20, dtype=torch.uint8).pin_memory()
dd = torch.empty(4096, 1 << 17, dtype=torch.uint8, device="cuda")
for _ in range(2): hp[:, :1 << 17].copy_(dd, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(8): hp[:, k << 17:(k + 1) << 17].copy_(dd, non_blocking=True)
torch.cuda.synchronize()
print("D2H 2-D (4096 x 128 KiB rows) GB/s", 8 * dd.numel() / (time.perf_counter() - t0) / 1e9)
