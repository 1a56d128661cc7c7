"""Developer probe: GPU encoder throughput on the C2 corpus (device resident).
Usage: python tools/enc_probe.py [streams] [formats, e.g. lz10,yaz0,mio0,yay0] [qualities, e.g. 0,8] [finders, e.g. default,parallel,serial]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from auroralib.compression_b200 import BatchCodec, _abi as A, corpus
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
codec = BatchCodec(1)
dev = torch.device("cuda:0")
raw, _ = corpus.generate_mix(n, 65536, device=dev)
FMT = {"lz10": A.FMT_LZ10, "yaz0": A.FMT_YAZ0, "lzss": A.FMT_LZSS, "mio0": A.FMT_MIO0, "yay0": A.FMT_YAY0, "lz11": A.FMT_LZ11, "blz": A.FMT_BLZ}
FINDER = {"default": 0, "parallel": A.STRATEGY_PARALLEL_FINDER, "serial": A.STRATEGY_SERIAL_FINDER}
names = (sys.argv[2] if len(sys.argv) > 2 else "lz10,yaz0").split(",")
quals = [int(q) for q in (sys.argv[3] if len(sys.argv) > 3 else "0,8").split(",")]
finders = (sys.argv[4] if len(sys.argv) > 4 else "default").split(",")
for name, q, finder in ((a, b, c) for a in names for b in quals for c in finders):
    fmt = FMT[name]
    if True:
        bound = (codec.encode_bound(fmt, 65536) + 15) & ~15
        i64 = dict(dtype=torch.int64, device=dev)
        r_off = torch.arange(n, **i64) * 65536
        r_len = torch.full((n,), 65536, **i64)
        c_buf = torch.empty(n * bound + 16, dtype=torch.uint8, device=dev)
        c_off = torch.arange(n, **i64) * bound
        c_cap = torch.full((n,), bound, **i64)
        c_len = torch.zeros(n, **i64)
        st = torch.zeros(n, dtype=torch.int32, device=dev)
        ts = torch.cuda.Stream()
        best = 1e9
        for it in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            codec.encode_device(fmt, raw.view(-1), r_off, r_len, c_buf, c_off, c_cap, c_len, st, A.make_opts(quality=q, strategy=FINDER[finder]), stream=ts.cuda_stream)
            ts.synchronize()
            best = min(best, time.perf_counter() - t0)
        assert int(st.abs().sum()) == 0
        print(f"{name} q{q} {finder}: {n} x 64 KiB in {best*1e3:.1f} ms = {n*65536/best/1e9:.2f} GB/s raw in, ratio {int(c_len.sum())/(n*65536):.4f}", flush=True)
