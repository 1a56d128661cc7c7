"""Developer probe (NOT the benchmark): device-resident decode timing per corpus class with oracle-encoded
inputs.  Usage: python tools/perf_probe.py [--streams N] [--size B] [--formats lz10,yaz0,lz4b] [--iters K]"""
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
from auroralib.compression_b200.batch import layout, pack  # noqa: E402
from oracle import oracle as O  # noqa: E402

FMT = {"lz10": A.FMT_LZ10, "yaz0": A.FMT_YAZ0, "lz4b": A.FMT_LZ4_BLOCK, "mio0": A.FMT_MIO0, "yay0": A.FMT_YAY0,
       "lz11": A.FMT_LZ11, "lzss": A.FMT_LZSS, "lzo": A.FMT_LZO, "snappyb": A.FMT_SNAPPY_BLOCK, "prs": A.FMT_PRS,
       "lz4": A.FMT_LZ4, "snappy": A.FMT_SNAPPY, "lzhudson": A.FMT_LZHUDSON, "lz40": A.FMT_LZ40, "lz60": A.FMT_LZ60, "smsr00": A.FMT_SMSR00, "blz": A.FMT_BLZ}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=8192)
    ap.add_argument("--size", type=int, default=65536)
    ap.add_argument("--formats", default="lz10,yaz0,lz4b")
    ap.add_argument("--classes", default="T,M,X,mix")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--quality", type=int, default=8)
    ap.add_argument("--no-verify", action="store_true", help="skip the byte comparison (experimental kernel variants)")
    args = ap.parse_args()
    peak = 6551.7
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    codec = BatchCodec(1)
    dev = torch.device("cuda:0")
    for cname in args.classes.split(","):
        if cname == "mix":
            raw, _ = corpus.generate_mix(args.streams, args.size, device=dev)
        else:
            raw = corpus.generate(cname, args.streams, args.size, device=dev)
        raw_h = raw.cpu().numpy()
        n = raw_h.shape[0]
        roff = np.arange(n, dtype=np.uint64) * np.uint64(args.size)
        rlen = np.full(n, args.size, dtype=np.uint64)
        for fname in args.formats.split(","):
            fmt = FMT[fname]
            caps, coff, ctotal = layout([codec.encode_bound(fmt, args.size)] * n)
            comp = np.zeros(ctotal + 16, dtype=np.uint8)
            t0 = time.time()
            clen, st = O.encode_packed(fmt, raw_h.reshape(-1), roff, rlen, comp, coff, caps, A.make_opts(quality=args.quality))
            t_enc = time.time() - t0
            assert (st == 0).all()
            # repack tightly (16-byte aligned) like a real batch
            blobs_off = np.zeros(n, dtype=np.uint64)
            padded = (clen + np.uint64(15)) & ~np.uint64(15)
            blobs_off[1:] = np.cumsum(padded[:-1])
            tight = np.zeros(int(padded.sum()) + 16, dtype=np.uint8)
            for i in range(n):
                tight[int(blobs_off[i]):int(blobs_off[i]) + int(clen[i])] = comp[int(coff[i]):int(coff[i]) + int(clen[i])]
            d_src = torch.from_numpy(tight).to(dev)
            d_off = torch.from_numpy(blobs_off.astype(np.int64)).to(dev)
            d_len = torch.from_numpy(clen.astype(np.int64)).to(dev)
            d_dst = torch.zeros(n * args.size + 16, dtype=torch.uint8, device=dev)
            d_doff = torch.from_numpy(roff.astype(np.int64)).to(dev)
            d_cap = torch.from_numpy(rlen.astype(np.int64)).to(dev)
            d_olen = torch.zeros(n, dtype=torch.int64, device=dev)
            d_cons = torch.zeros(n, dtype=torch.int64, device=dev)
            d_st = torch.zeros(n, dtype=torch.int32, device=dev)
            ts = torch.cuda.Stream()   # a real (non-default) stream: handle 0 means "the context's own stream" in the ABI
            torch.cuda.synchronize()
            times = []
            for it in range(args.iters + 2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(ts):
                    e0.record(ts)
                    codec.decode_device(fmt, d_src, d_off, d_len, d_dst, d_doff, d_cap, d_olen, d_cons, d_st, stream=ts.cuda_stream)
                    e1.record(ts)
                torch.cuda.synchronize()
                if it >= 2:
                    times.append(e0.elapsed_time(e1))
            ok = d_st == 0
            if fname not in ("lzo", "prs"):   # the reference's LZO encoder / PRS order heuristic have known self-inconsistencies
                assert bool(ok.all()), "decode status"
            # PRS / LZO: the reference's order heuristic / its encoder's dropped-first-match quirk can "succeed" with other bytes
            # (the parity tests compare those with the oracle instead of the raw input)
            if fname not in ("prs", "lzo") and not args.no_verify:
                assert torch.equal(d_dst[:n * args.size].view(n, args.size)[ok], raw[ok]), "decode mismatch"
                if not bool(ok.all()):
                    print(f"  ({int((~ok).sum())} of {n} streams not OK by design of the reference's encoder / order heuristic)")
            ms = float(np.median(times))
            out_b, in_b = n * args.size, int(clen.sum())
            print(f"{fname:8s} class {cname:3s} n={n} ratio {in_b / out_b:.3f}  {ms:8.3f} ms  out {out_b / ms / 1e6:8.1f} GB/s  "
                  f"in+out {(in_b + out_b) / ms / 1e6:8.1f} GB/s  roofline {100 * (in_b + out_b) / ms / 1e6 / peak:5.1f}%  (oracle enc {out_b / t_enc / 1e6:.0f} MB/s)",
                  flush=True)


if __name__ == "__main__":
    main()
