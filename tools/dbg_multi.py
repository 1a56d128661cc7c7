"""Developer tool (2+ GPUs): timeline of device-resident decode launches on several devices of one context."""
import sys, time, threading, torch
sys.path.insert(0, '/root/repo')
import bench
from auroralib.compression_b200 import BatchCodec, _abi as A, corpus
G = min(int(sys.argv[1]) if len(sys.argv) > 1 else 2, torch.cuda.device_count())
n, size, steps = 16384, 65536, 4
codec = BatchCodec(device_mask=(1 << G) - 1)
fmt = A.FMT_LZ10
sh = []
for g in range(G):
    dev = torch.device("cuda", g); torch.cuda.set_device(dev); ts = torch.cuda.Stream(device=dev)
    raw, _ = corpus.generate_mix(n, size, seed=1 + g, device=dev)
    r_off = torch.arange(n, dtype=torch.int64, device=dev) * size; r_len = torch.full((n,), size, dtype=torch.int64, device=dev)
    packed, p_off, c_len, tot, _ = bench.gpu_encode(codec, fmt, raw.view(-1), r_off, r_len, A.make_opts(quality=8), dev, ts, device_index=g)
    d_dst = torch.zeros(n * size + 16, dtype=torch.uint8, device=dev)
    sh.append(dict(dev=dev, ts=ts, a=(packed, p_off, c_len, d_dst, r_off, r_len, torch.zeros(n, dtype=torch.int64, device=dev), torch.zeros(n, dtype=torch.int64, device=dev), torch.zeros(n, dtype=torch.int32, device=dev))))
def run(devs, threaded, label):
    for g in devs:
        s = sh[g]; torch.cuda.set_device(s["dev"])
        for _ in range(2): codec.decode_device(fmt, *s["a"], device=g, stream=s["ts"].cuda_stream)
        s["ts"].synchronize()
    ev = {}
    for g in devs:
        torch.cuda.set_device(sh[g]["dev"])
        ev[g] = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    def drive(g):
        s = sh[g]; torch.cuda.set_device(s["dev"])
        for k in range(steps):
            ev[g][k][0].record(s["ts"]); codec.decode_device(fmt, *s["a"], device=g, stream=s["ts"].cuda_stream); ev[g][k][1].record(s["ts"])
    if threaded:
        th = [threading.Thread(target=drive, args=(g,)) for g in devs]; [t.start() for t in th]; [t.join() for t in th]
    else:
        for k in range(steps):
            for g in devs:
                s = sh[g]; torch.cuda.set_device(s["dev"])
                ev[g][k][0].record(s["ts"]); codec.decode_device(fmt, *s["a"], device=g, stream=s["ts"].cuda_stream); ev[g][k][1].record(s["ts"])
    for g in devs: sh[g]["ts"].synchronize()
    for g in devs:
        e = ev[g]
        print(label, "dev", g, "timeline (start,end ms):", [(round(e[0][0].elapsed_time(a), 2), round(e[0][0].elapsed_time(b), 2)) for a, b in e])
run([0], False, "only dev0     ")
run([1], False, "only dev1     ")
run(list(range(G)), False, "all, 1 thread ")
run(list(range(G)), True, "all, threaded ")
