"""Developer tool (no GPU): the decode kernels' device source on the CPU lane emulation under AddressSanitizer.
  python tools/simt_asan_decode.py cases      # stage 1, plain python: streams (valid + corrupt) and the oracle's answers -> /tmp/simt_asan_cases.pkl
  ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:verify_asan_link_order=0 \\
  LD_PRELOAD=$(gcc -print-file-name=libasan.so) python tools/simt_asan_decode.py run    # stage 2: emulation libraries built with -fsanitize=address
(two stages because the sanitizer's __cxa_throw interceptor does not get along with the exceptions inside liboracle.so)."""
import ctypes as C
import os
import pickle
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auroralib.compression_b200 import _abi as A  # noqa: E402
from tests import test_simt_kernels as T  # noqa: E402

CASES = "/tmp/simt_asan_cases.pkl"


def stage_cases():
    from oracle import oracle as O
    from tests.util import corrupt, synth
    O.build()
    bmp = open(os.path.join(ROOT, "tests", "golden", "Test.bmp"), "rb").read()
    rng = np.random.default_rng(1)
    out = []
    for fmt in T.FLAG_DEC_FORMATS + T.BYTE_FORMATS:
        raws = [bmp[:9000]] + [synth(rng, int(k), i % 5) for i, k in enumerate([6, 33, 1000, 4096, 6000, 12000, 70000])]
        comps, st = O.encode_batch(fmt, raws, A.make_opts(quality=8))
        streams, caps = [], []
        for c, r, s in zip(comps, raws, st):
            if s:
                continue
            streams.append(c)
            caps.append(len(r))
            for mode in range(5):
                streams.append(corrupt(rng, c, mode))
                caps.append(len(r))
            streams.append(c)
            caps.append(max(len(r) - 1, 0))
        ref, rl, rc, rs = O.decode_batch(fmt, streams, caps, A.make_opts())
        out.append((fmt, streams, caps, rl, rc, rs))
    pickle.dump(out, open(CASES, "wb"))
    print("cases", sum(len(x[1]) for x in out))


def stage_run():
    libs = {}
    for name, harness in (("decode_flaglz", "flaglz_dec_harness.cpp"), ("decode_bytelz", "bytelz_dec_harness.cpp")):
        src = open(os.path.join(T.CSRC, name + ".cu")).read()
        inc = f"/tmp/simt_asan_{name}.inc"
        open(inc, "w").write(src[:src.index("// ---- kernel\n")])
        so = f"/tmp/libsimt_asan_{name}.so"
        subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=address", "-fno-omit-frame-pointer", "-std=c++17", "-shared", "-fPIC", "-I", T.SIMT,
                               "-I", T.CSRC, f'-DAURORA_REAL_COMMON="{os.path.join(T.CSRC, "common.cuh")}"',
                               f'-DAURORA_REAL_STAGE="{os.path.join(T.CSRC, "stage.cuh")}"', f'-DDEC_DEVICE_INC="{inc}"',
                               os.path.join(T.SIMT, harness), "-o", so])
        libs[name] = C.CDLL(so)
    flag, byte = libs["decode_flaglz"].simt_decode_flaglz, libs["decode_bytelz"].simt_decode_bytelz
    flag.restype = byte.restype = C.c_int
    n = 0
    for fmt, streams, caps, rl, rc, rs in pickle.load(open(CASES, "rb")):
        is_flag = fmt in T.FLAG_DEC_FORMATS
        got, ol, co, st = T.simt_decode_bytelz(flag if is_flag else byte, fmt, streams, caps, flag_lz=is_flag)
        assert (st == rs).all() and (ol == rl).all() and (co == rc).all()
        n += len(streams)
        print(A.FORMAT_NAMES[fmt], "ok", len(streams), flush=True)
    print("no AddressSanitizer report over", n, "streams")


if __name__ == "__main__":
    stage_cases() if sys.argv[1:] == ["cases"] else stage_run()
