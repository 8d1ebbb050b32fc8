"""PCIe probe: pinned host -> device copy rate for one bench step's luma (103.7 MB), the ceiling of bench.py's e2e."""
import torch

n = 50 * 1920 * 1080
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    d.copy_(h, non_blocking=True)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print("H2D %.1f MB in %.3f ms = %.1f GB/s -> e2e ceiling %.2fe6 CTU/s" % (n / 1e6, ms, n / ms / 1e6, 25500 / ms / 1e3))
