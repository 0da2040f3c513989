#!/usr/bin/env python
"""Host<->device copy bandwidth on this box (pinned memory), alone and both directions at once: the bound of the e2e path."""
import torch, time
n = 400_000_000
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both():
    h2d(); d2h()
for name, fn, b in (("h2d", h2d, n), ("d2h", d2h, n), ("both", both, 2 * n)):
    dt = t(fn)
    print(f"{name}: {b / dt / 1e9:.1f} GB/s ({dt * 1e3:.2f} ms)")
