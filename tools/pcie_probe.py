"""Host <-> device copy rates of this box for the size bench.py's e2e path moves (136 MB each way), alone and concurrently.
The e2e number of bench.py is bounded by the concurrent rate: 2 x 136 MB per apply over PCIe."""
import json
import torch

n = 16974593
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    s1.synchronize()
    s2.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    d_in.copy_(h_in, non_blocking=True)


def d2h():
    h_out.copy_(d_out, non_blocking=True)


def both():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


gb = n * 8 / 1e9
t1, t2, t3 = timed(h2d), timed(d2h), timed(both)
print(json.dumps({"bytes_each_way": n * 8, "h2d_ms": t1, "h2d_gbs": gb / t1 * 1e3, "d2h_ms": t2, "d2h_gbs": gb / t2 * 1e3,
                  "concurrent_ms": t3, "concurrent_gbs_each_way": gb / t3 * 1e3, "e2e_bound_gdofs": n / t3 / 1e6}))
