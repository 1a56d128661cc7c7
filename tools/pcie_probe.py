"""Developer probe: pinned host <-> device copy bandwidth on 1..N GPUs AT ONCE — the ceiling of bench.py's `e2e`.

For every device: H2D alone, D2H alone, both directions at once (a C2 step moves 2.10 GB up and 4.30 GB down, so a step
cannot take less than the `both` time).  With --gpus N the same copies run on N devices concurrently (one host thread and
two CUDA streams per device), which gives the box's N-link ceiling: the sum over devices of decoded bytes / step time."""
import argparse
import threading
import time

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--scale", type=float, default=1.0, help="fraction of the C2 step's bytes per device")
args = ap.parse_args()
G = min(args.gpus, torch.cuda.device_count())
n_up, n_down = int(2_104_965_136 * args.scale), int(4_296_278_016 * args.scale)   # bytes per step of config C2 (bench.py e2e: h2d / d2h)

bufs = []
for g in range(G):
    torch.cuda.set_device(g)
    bufs.append(dict(h_up=torch.empty(n_up, dtype=torch.uint8).pin_memory(), h_down=torch.empty(n_down, dtype=torch.uint8).pin_memory(),
                     d_up=torch.empty(n_up, dtype=torch.uint8, device=f"cuda:{g}"), d_down=torch.empty(n_down, dtype=torch.uint8, device=f"cuda:{g}"),
                     s1=torch.cuda.Stream(device=g), s2=torch.cuda.Stream(device=g)))


def run(devs, up, down, reps=3):
    """All `devs` start together; returns the best wall time (s) of the slowest device."""
    best = 1e9
    for _ in range(reps + 1):
        for g in devs:
            torch.cuda.synchronize(g)
        bar = threading.Barrier(len(devs) + 1)
        done = []

        def work(g):
            b = bufs[g]
            torch.cuda.set_device(g)
            bar.wait()
            if up:
                with torch.cuda.stream(b["s1"]):
                    b["d_up"].copy_(b["h_up"], non_blocking=True)
            if down:
                with torch.cuda.stream(b["s2"]):
                    b["h_down"].copy_(b["d_down"], non_blocking=True)
            b["s1"].synchronize()
            b["s2"].synchronize()
            done.append(time.perf_counter())

        th = [threading.Thread(target=work, args=(g,)) for g in devs]
        for t in th:
            t.start()
        bar.wait()
        t0 = time.perf_counter()
        for t in th:
            t.join()
        best = min(best, max(done) - t0)
    return best


for k in sorted({1, 2, 4, 8, G}):
    if k > G:
        continue
    devs = list(range(k))
    tu, td, tb = run(devs, True, False), run(devs, False, True), run(devs, True, True)
    print(f"{k} GPU(s) at once: H2D {k * n_up / tu / 1e9:6.1f} GB/s  D2H {k * n_down / td / 1e9:6.1f} GB/s  both {tb * 1e3:7.1f} ms "
          f"-> e2e ceiling of C2 = {k * n_down / tb / 1e9:6.1f} GB/s decoded ({n_down / tb / 1e9:5.1f} per GPU)")
