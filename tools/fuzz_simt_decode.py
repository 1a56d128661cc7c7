"""Developer tool (no GPU): differential fuzz of the DECODE kernels' device source on the CPU lane emulation (tests/simt) against
the oracle: flag-LZ kernel (parser / resolver warp pair), byte-LZ kernel, BLZ.  Valid and corrupted streams, exact / short /
larger destinations; status, out_len, consumed and every decoded byte.  Usage: python tools/fuzz_simt_decode.py [cases] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auroralib.compression_b200 import _abi as A  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests import test_simt_kernels as T  # noqa: E402
from tests.util import corrupt, synth  # noqa: E402


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    import ctypes as C
    flag = T._build("decode_flaglz", "flaglz_dec_harness.cpp", "DEC_DEVICE_INC").simt_decode_flaglz
    byte = T._build("decode_bytelz", "bytelz_dec_harness.cpp", "DEC_DEVICE_INC").simt_decode_bytelz
    blz = T._build("decode_blz", "blz_harness.cpp", "BLZ_DEVICE_INC").simt_decode_blz
    for f in (flag, byte, blz):
        f.restype = C.c_int
    O.build()
    rng = np.random.default_rng(seed)
    bmp = open(os.path.join(ROOT, "tests", "golden", "Test.bmp"), "rb").read()
    pool = T.FLAG_DEC_FORMATS + T.BYTE_FORMATS + [A.FMT_BLZ]
    t0 = time.time()
    streams_total = ok_total = 0
    for case in range(cases):
        fmt = pool[int(rng.integers(0, len(pool)))]
        q = int(rng.choice([0, 5, 8, 13]))
        raws = []
        for i in range(int(rng.integers(1, 8))):
            n = int(rng.choice([1, 5, 6, 31, 32, 33, 64, 65, 100, 1000, 4095, 4096, 4097, 5000, 9000, 20000, 70000], p=[0.06] * 16 + [0.04]))
            kind = int(rng.integers(0, 6))
            if kind == 5:
                o = int(rng.integers(0, len(bmp) - n - 1))
                raws.append(bmp[o:o + n])
            else:
                raws.append(synth(rng, n, kind))
        eopts = dict(quality=q)
        order = A.ENDIAN_DEFAULT
        if fmt in (A.FMT_YAZ0, A.FMT_YAY0, A.FMT_MIO0, A.FMT_PRS) and rng.integers(0, 2):
            eopts["byte_order"] = int(rng.choice([A.ENDIAN_BIG, A.ENDIAN_LITTLE]))
            order = int(rng.choice([A.ENDIAN_DEFAULT, eopts["byte_order"]]))   # decode with the detection, or told
        comps, st = O.encode_batch(fmt, raws, A.make_opts(**eopts))
        streams, caps = [], []
        for c, r, s in zip(comps, raws, st):
            if s != 0:
                continue
            mode = int(rng.integers(0, 8))
            streams.append(corrupt(rng, c, mode) if mode < 5 else c)
            caps.append(int(rng.choice([len(r), len(r), max(len(r) - 1, 0), max(len(r) - 100, 0), len(r) + 57])))
        if not streams:
            continue
        dopts = A.make_opts(byte_order=order)
        ref, rlen, rcons, rst = O.decode_batch(fmt, streams, caps, dopts)
        if fmt == A.FMT_BLZ:
            got, out_len, consumed, status = T.simt_decode_blz(blz, streams, caps)
        else:
            got, out_len, consumed, status = T.simt_decode_bytelz(flag if fmt in T.FLAG_DEC_FORMATS else byte, fmt, streams, caps,
                                                                  byte_order=order, flag_lz=fmt in T.FLAG_DEC_FORMATS)
        bad = [i for i in range(len(streams)) if status[i] != rst[i] or out_len[i] != rlen[i] or consumed[i] != rcons[i]
               or (rst[i] == 0 and got[i] != ref[i])]
        if bad:
            i = bad[0]
            print(f"MISMATCH case {case}: {A.FORMAT_NAMES[fmt]} q{q} order {order} stream {i} (len {len(streams[i])}, cap {caps[i]}): "
                  f"status {status[i]} / {rst[i]}, out_len {out_len[i]} / {rlen[i]}, consumed {consumed[i]} / {rcons[i]}", flush=True)
            np.save(f"/tmp/fuzz_simt_decode_case{case}.npy", np.frombuffer(streams[i], dtype=np.uint8))
            sys.exit(1)
        streams_total += len(streams)
        ok_total += int((rst == 0).sum())
        if case % 50 == 49:
            print(f"{case + 1} cases, {streams_total} streams ({ok_total} decode OK per oracle), {time.time() - t0:.0f} s: identical", flush=True)
    print(f"done: {cases} cases, {streams_total} streams ({ok_total} decode OK per oracle) identical to the oracle decoder ({time.time() - t0:.0f} s)")


if __name__ == "__main__":
    main()
