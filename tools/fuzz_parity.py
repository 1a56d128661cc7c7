"""Developer tool (GPU box): a larger differential fuzz of the decode kernels against the oracle than the test-suite runs
(other seeds, longer streams): valid, truncated, bit-flipped, padded and empty inputs; status, out_len, consumed and bytes.
Usage: python tools/fuzz_parity.py [seed] [streams per format]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from auroralib.compression_b200 import BatchCodec, _abi as A
from oracle import oracle as O
from tests.util import ALL_FORMATS, corrupt, fmt_id, synth

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 77
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
O.build()
codec = BatchCodec(1)
total_bad = 0
for fmt in ALL_FORMATS:
    rng = np.random.default_rng(seed * 1000 + fmt)
    raws = [synth(rng, int(rng.choice([rng.integers(5, 2000), rng.integers(2000, 20000), rng.integers(20000, 90000)], p=[0.5, 0.35, 0.15])), i % 5) for i in range(n)]
    comps, caps = [], []
    for q in (0, 5, 8, 13):
        part = raws[q % 4::4] if q != 13 else raws[3::4]
        c, st = O.encode_batch(fmt, part, A.make_opts(quality=q))
        for r, cc, s in zip(part, c, st):
            if s != 0:
                continue
            mode = int(rng.integers(0, 8))   # 0..4: corrupt() modes, 5..7: leave valid
            comps.append(corrupt(rng, cc, mode) if mode < 5 else cc)
            caps.append(max(0, len(r) + int(rng.choice([0, 0, 0, 0, 64, 5000, -1, -100]))))
    outs, ol, co, gs = codec.decode_batch(fmt, comps, caps)
    ref, rl, rc, rs = O.decode_batch(fmt, comps, caps)
    bad = [i for i in range(len(comps)) if outs[i] != ref[i] or gs[i] != rs[i] or ol[i] != rl[i] or co[i] != rc[i]]
    total_bad += len(bad)
    print(f"{fmt_id(fmt):12s} streams {len(comps):5d} ok-status {int((rs == 0).sum()):5d} mismatches {len(bad)}" + (f" first #{bad[0]} gpu {gs[bad[0]]}/{ol[bad[0]]}/{co[bad[0]]} ref {rs[bad[0]]}/{rl[bad[0]]}/{rc[bad[0]]}" if bad else ""), flush=True)
print("TOTAL mismatches", total_bad)
sys.exit(1 if total_bad else 0)
