"""Host <-> device copy bandwidth of this box from pinned memory (the ceiling of the end-to-end leg)."""
import time

import torch

for mb in (8, 64, 256):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        print(f"{name} {mb} MB: {10 * n / (time.perf_counter() - t) / 1e9:.1f} GB/s")
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10):
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print(f"both directions {mb} MB each: {10 * n / dt / 1e9:.1f} GB/s per direction")
