"""Developer tool (no GPU): differential fuzz of the lane-per-position encoder's device source on the CPU lane emulation
(tests/simt) against the oracle.  Usage: python tools/fuzz_simt.py [cases] [seed] [big]
Every case runs the lane-per-position search and, for a third of the cases, the sequential replay (any encoder format)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auroralib.compression_b200 import _abi as A  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests import test_simt_kernels as T  # noqa: E402
from tests.util import synth  # noqa: E402


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    big = len(sys.argv) > 3 and sys.argv[3] == "big"   # a third of the buffers over 64 KiB (16-bit table positions wrap)
    lib = T.SimtLibs()   # builds the three emulation libraries from the current kernel sources
    O.build()
    rng = np.random.default_rng(seed)
    bmp = open(os.path.join(ROOT, "tests", "golden", "Test.bmp"), "rb").read()
    t0 = time.time()
    streams = 0
    for case in range(cases):
        seq = case % 3 == 2
        pool = T.SEQ_FLAG_FORMATS + T.BYTE_FORMATS if seq else T.PAR_FORMATS
        fmt = pool[int(rng.integers(0, len(pool)))]
        q = int(rng.integers(0, 16))
        strategy = int(rng.integers(0, 2))
        kw = dict(strategy=strategy, skew=int(rng.integers(0, 16)), seq=seq)
        if fmt in (A.FMT_LZ10, A.FMT_LZ11) and rng.integers(0, 2):
            kw["vram_mode"] = int(rng.integers(0, 2))
        if fmt in (A.FMT_YAZ0, A.FMT_YAY0, A.FMT_MIO0) and rng.integers(0, 2):
            kw["byte_order"] = int(rng.choice([A.ENDIAN_BIG, A.ENDIAN_LITTLE]))
        if fmt == A.FMT_LZSS and rng.integers(0, 2):
            kw["lzss"] = A.lz_props_bits(int(rng.integers(8, 13)), int(rng.integers(3, 9)), int(rng.integers(1, 4)))
        raws = []
        for i in range(int(rng.integers(1, 10))):
            n = int(rng.choice([0, 1, 3, 4, 5, 31, 32, 33, 63, 64, 65, 100, 1000, 4095, 4096, 4097, 5000, 9000, 20000]))
            if big and rng.integers(0, 3) == 0:
                n = int(rng.integers(65536, 150000))
            kind = int(rng.integers(0, 6))
            if kind == 5:
                o = int(rng.integers(0, len(bmp) - n - 1))
                raws.append(bmp[o:o + n])
            else:
                raws.append(synth(rng, n, kind))
        try:
            T._check(lib, O, fmt, raws, q, **kw)
        except AssertionError as e:
            print(f"MISMATCH case {case}: fmt {A.FORMAT_NAMES[fmt]} q{q} {kw} sizes {[len(r) for r in raws]}: {e}", flush=True)
            np.save(f"/tmp/fuzz_simt_case{case}.npy", np.array([np.frombuffer(r, dtype=np.uint8) for r in raws], dtype=object), allow_pickle=True)
            sys.exit(1)
        streams += len(raws)
        if case % 50 == 49:
            print(f"{case + 1} cases, {streams} buffers, {time.time() - t0:.0f} s: byte-identical", flush=True)
    print(f"done: {cases} cases, {streams} buffers byte-identical to the oracle encoder ({time.time() - t0:.0f} s)")


if __name__ == "__main__":
    main()
