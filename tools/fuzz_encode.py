"""Developer tool (GPU box): differential fuzz of the GPU encoders against the oracle encoder (byte identity), both match finders.
Usage: python tools/fuzz_encode.py [seed] [buffers per case]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from auroralib.compression_b200 import BatchCodec, _abi as A
from oracle import oracle as O
from tests.util import fmt_id, synth

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 91
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
O.build()
codec = BatchCodec(1)
bad_total = cases = 0
for fmt in (A.FMT_LZ10, A.FMT_BLZ, A.FMT_YAZ0, A.FMT_YAZ1, A.FMT_LZSS, A.FMT_MIO0, A.FMT_YAY0, A.FMT_LZ11, A.FMT_LZ40, A.FMT_LZ60, A.FMT_LZ4_BLOCK, A.FMT_LZO, A.FMT_PRS):
    for q in (0, 1, 2, 4, 6, 7, 9, 10, 15):
        rng = np.random.default_rng(seed * 100000 + fmt * 100 + q)
        raws = [synth(rng, int(rng.choice([rng.integers(0, 70), rng.integers(70, 5000), rng.integers(5000, 150000)], p=[0.2, 0.5, 0.3])), int(rng.integers(0, 5))) for _ in range(n)]
        for strat in (0, A.STRATEGY_PARALLEL_FINDER, A.STRATEGY_SERIAL_FINDER | A.STRATEGY_COMPATIBILITY, A.STRATEGY_PARALLEL_FINDER | A.STRATEGY_COMPATIBILITY):
            opts = A.make_opts(quality=q, strategy=strat, vram_mode=int(rng.integers(-1, 2)))
            got, st = codec.encode_batch(fmt, raws, opts)
            ref, rst = O.encode_batch(fmt, raws, opts)
            bad = [i for i in range(n) if st[i] != rst[i] or got[i] != ref[i]]
            bad_total += len(bad)
            cases += 1
            if bad:
                print(f"{fmt_id(fmt)} q{q} strategy {strat:#x}: {len(bad)} of {n} differ, first #{bad[0]} len {len(raws[bad[0]])}", flush=True)
    print(f"{fmt_id(fmt):10s} done", flush=True)
print("cases", cases, "TOTAL mismatches", bad_total)
sys.exit(1 if bad_total else 0)
