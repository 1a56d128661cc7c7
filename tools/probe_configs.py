"""Developer probe (NOT the benchmark): configs C3 / C4 of BASELINE.json at reduced stream counts, device resident.
  C3: Yaz0 / Yay0 / MIO0, decoded sizes log-uniform in [256 KiB, 4 MiB], classes T/M/X/B, BE and LE headers
  C4: LZ4 / LZO / Snappy blocks, decoded sizes log-uniform in [4 KiB, 64 KiB], classes T/M/X
Inputs are encoded by the CPU oracle (this is a dev tool; bench.py never does that)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auroralib.compression_b200 import BatchCodec, _abi as A, corpus  # noqa: E402
from oracle import oracle as O  # noqa: E402


def build(config, n, seed):
    rng = np.random.default_rng(seed)
    if config == "c3":
        lo, hi, classes = 256 << 10, 4 << 20, "TMXB"
    else:
        lo, hi, classes = 4 << 10, 64 << 10, "TMX"
    sizes = np.exp(rng.uniform(np.log(lo), np.log(hi), size=n)).astype(np.int64)
    raws = []
    for ci, c in enumerate(classes):
        idx = [i for i in range(n) if i % len(classes) == ci]
        if not idx:
            continue
        mx = int(max(sizes[i] for i in idx))
        # class generators are periodic in 64 KiB tiles for T; generate the longest and cut prefixes
        x = corpus.generate(c, len(idx), mx, seed=seed + ci, device="cuda").cpu().numpy()
        for k, i in enumerate(idx):
            raws.append((i, x[k, :sizes[i]].tobytes()))
    raws.sort()
    return [r for _, r in raws]


def run(codec, fmt, raws, opts_enc, balance, iters=3):
    n = len(raws)
    comps, st = O.encode_batch(fmt, raws, opts_enc)
    assert (st == 0).all()
    from auroralib.compression_b200.batch import layout, pack
    base, off, ln = pack(comps)
    caps, doff, total = layout([len(r) for r in raws])
    dev = torch.device("cuda:0")
    d_src = torch.from_numpy(base).to(dev)
    d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
    d_len = torch.from_numpy(ln.astype(np.int64)).to(dev)
    d_dst = torch.zeros(total + 16, dtype=torch.uint8, device=dev)
    d_doff = torch.from_numpy(doff.astype(np.int64)).to(dev)
    d_cap = torch.from_numpy(caps.astype(np.int64)).to(dev)
    d_ol = torch.zeros(n, dtype=torch.int64, device=dev)
    d_co = torch.zeros(n, dtype=torch.int64, device=dev)
    d_st = torch.zeros(n, dtype=torch.int32, device=dev)
    ts = torch.cuda.Stream()
    torch.cuda.synchronize()
    times = []
    opts = A.make_opts(balance=balance, byte_order=opts_enc.byte_order)
    for it in range(iters + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ts):
            e0.record(ts)
            codec.decode_device(fmt, d_src, d_off, d_len, d_dst, d_doff, d_cap, d_ol, d_co, d_st, opts=opts, stream=ts.cuda_stream)
            e1.record(ts)
        torch.cuda.synchronize()
        if it >= 2:
            times.append(e0.elapsed_time(e1))
    ok = (d_st == 0).cpu().numpy()
    out = d_dst.cpu().numpy()
    good = all(out[int(doff[i]):int(doff[i]) + len(raws[i])].tobytes() == raws[i] for i in range(n) if ok[i])
    ms = float(np.median(times))
    ob, ib = sum(len(r) for r in raws), sum(len(c) for c in comps)
    return ms, ob, ib, int(ok.sum()), good


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--streams", type=int, default=0)
    args = ap.parse_args()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6551.7
    codec = BatchCodec(1)
    if args.config == "c3":
        n = args.streams or 1024
        fmts = [("yaz0", A.FMT_YAZ0), ("yay0", A.FMT_YAY0), ("mio0", A.FMT_MIO0)]
    else:
        n = args.streams or 131072
        fmts = [("lz4b", A.FMT_LZ4_BLOCK), ("lzo", A.FMT_LZO), ("snappyb", A.FMT_SNAPPY_BLOCK)]
    t0 = time.time()
    raws = build(args.config, n, 0xA0130000 if args.config == "c3" else 0xA0140000)
    print(f"{args.config}: {n} streams, {sum(map(len, raws)) / 2**30:.2f} GiB decoded, built in {time.time() - t0:.1f}s", flush=True)
    for name, fmt in fmts:
        for order in ((A.ENDIAN_BIG, A.ENDIAN_LITTLE) if args.config == "c3" else (A.ENDIAN_DEFAULT,)):
            for balance in ((2, 1) if order != A.ENDIAN_LITTLE else (1,)):
                ms, ob, ib, nok, good = run(codec, fmt, raws, A.make_opts(quality=8, byte_order=order), balance)
                print(f"{name:8s} order={order} balance={'on ' if balance == 1 else 'off'} {ms:9.3f} ms  out {ob / ms / 1e6:8.1f} GB/s  "
                      f"in+out {(ob + ib) / ms / 1e6:8.1f} GB/s  roofline {100 * (ob + ib) / ms / 1e6 / peak:5.1f}%  ratio {ib / ob:.3f}  ok {nok}/{n} bytes_equal={good}", flush=True)


if __name__ == "__main__":
    main()
