"""Diagnostic (GPU): do pinned H2D and D2H copies overlap on this box?  (torch is used for streams only.)"""
import time
import torch
a_h = torch.empty(5292000 // 4, dtype=torch.float32).pin_memory()
b_h = torch.empty(8248464 // 4, dtype=torch.float32).pin_memory()
a_d = torch.empty_like(a_h, device="cuda")
b_d = torch.empty_like(b_h, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, n=20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        if h2d:
            with torch.cuda.stream(s1): a_d.copy_(a_h, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): b_h.copy_(b_d, non_blocking=True)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6
for _ in range(2):
    print(f"h2d 5.29 MB alone {run(True, False):7.1f} us | d2h 8.25 MB alone {run(False, True):7.1f} us | both {run(True, True):7.1f} us")
