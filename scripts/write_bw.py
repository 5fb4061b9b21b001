"""Calibration: HBM write-only bandwidth (torch fill) next to the copy bandwidth, same method as MEASURED_PEAKS."""
import torch
dev = torch.device("cuda", 0)
n = 2_400_000_000 // 8
x = torch.empty(n, dtype=torch.float64, device=dev)
y = torch.empty(n, dtype=torch.float64, device=dev)
def t(f, reps=10):
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
ms = t(lambda: x.zero_())
print(f"fill  {n*8/1e9:.2f} GB  {ms:.3f} ms  {n*8/ms/1e6:.1f} GB/s (write only)")
ms = t(lambda: y.copy_(x))
print(f"copy  {2*n*8/1e9:.2f} GB  {ms:.3f} ms  {2*n*8/ms/1e6:.1f} GB/s (read + write)")
ms = t(lambda: x.sum())
print(f"read  {n*8/1e9:.2f} GB  {ms:.3f} ms  {n*8/ms/1e6:.1f} GB/s (read only)")
