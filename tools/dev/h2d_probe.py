"""Dev probe: pinned host -> device copy bandwidth for one batch of points (8 x 120k x 16 B), the e2e leg's input."""
import torch
n = 8 * 120000 * 4
h = torch.empty(n, dtype=torch.float32).pin_memory(); d = torch.empty(n, dtype=torch.float32, device="cuda")
for _ in range(5): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
print("H2D %.1f MB: %.4f ms -> %.1f GB/s" % (n * 4 / 1e6, ms, n * 4 / ms / 1e6))
