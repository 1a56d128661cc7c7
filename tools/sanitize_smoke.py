"""Small decode/encode batch for compute-sanitizer (memcheck / racecheck): every kernel family, ragged + corrupt inputs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from auroralib.compression_b200 import BatchCodec, _abi as A
from oracle import oracle as O
from tests.util import ALL_FORMATS, corrupt, synth

codec = BatchCodec(1)
rng = np.random.default_rng(3)
raws = [synth(rng, int(n), i % 5) for i, n in enumerate([0, 1, 5, 17, 300, 4097, 9000, 20000, 70000, 33, 1000, 2500])]
bad = 0
for fmt in ALL_FORMATS + [A.FMT_ECD, A.FMT_LZ00, A.FMT_LZ77]:   # + the keystream pass, host-written prefixes, chunked sub-streams
    comps, st = O.encode_batch(fmt, raws, A.make_opts(quality=8))
    blobs, caps = [], []
    for r, c, s in zip(raws, comps, st):
        if s == 0:
            blobs += [c, corrupt(rng, c, 0), corrupt(rng, c, 1)]
            caps += [len(r), len(r), len(r) + 64]
    outs, ol, co, gs = codec.decode_batch(fmt, blobs, caps)
    ref, rl, rc, rs = O.decode_batch(fmt, blobs, caps)
    bad += sum(1 for i in range(len(blobs)) if outs[i] != ref[i] or gs[i] != rs[i])
for fmt in (A.FMT_LZ10, A.FMT_YAZ0, A.FMT_MIO0, A.FMT_LZSS, A.FMT_LZHUDSON, A.FMT_LZ40, A.FMT_SMSR00, A.FMT_LZ00, A.FMT_ECD, A.FMT_PRS):
    got, st = codec.encode_batch(fmt, raws[:9], A.make_opts(quality=8))
    ref, rst = O.encode_batch(fmt, raws[:9], A.make_opts(quality=8))
    bad += sum(1 for a, b in zip(got, ref) if a != b)
m = codec.is_match_batch(A.FMT_LZ10, [b"\x10\x05\x00\x00\x00abcde", b"xx"])
print("mismatches", bad)
sys.exit(1 if bad else 0)
