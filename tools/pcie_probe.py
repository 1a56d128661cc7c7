"""Developer probe: pinned host <-> device copy bandwidth (alone and both directions at once), the ceiling of bench.py's `e2e`."""
import torch

n_up, n_down = 2_104_965_136, 4_296_278_016   # bytes per step of config C2 (bench.py e2e: h2d / d2h)
h_up = torch.empty(n_up, dtype=torch.uint8).pin_memory()
h_down = torch.empty(n_down, dtype=torch.uint8).pin_memory()
d_up = torch.empty(n_up, dtype=torch.uint8, device="cuda")
d_down = torch.empty(n_down, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


t = timed(lambda: d_up.copy_(h_up, non_blocking=True))
print(f"H2D alone   {n_up / t / 1e6:6.1f} GB/s ({t:.1f} ms)")
t = timed(lambda: h_down.copy_(d_down, non_blocking=True))
print(f"D2H alone   {n_down / t / 1e6:6.1f} GB/s ({t:.1f} ms)")


def both():
    with torch.cuda.stream(s1):
        d_up.copy_(h_up, non_blocking=True)
    with torch.cuda.stream(s2):
        h_down.copy_(d_down, non_blocking=True)


t = timed(both)
print(f"both at once: {t:.1f} ms -> a step cannot take less; decoded {n_down / t / 1e6:6.1f} GB/s is the e2e ceiling of C2")
