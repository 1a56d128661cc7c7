"""Developer tool (no GPU): the encoder kernels' device source (lane-per-position search and sequential replay) on the CPU lane
emulation under AddressSanitizer; memory errors only, no comparison (tests/test_simt_kernels.py compares with the oracle).
  ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:verify_asan_link_order=0 \\
  LD_PRELOAD=$(gcc -print-file-name=libasan.so) python tools/simt_asan_encode.py"""
import sys, os, subprocess, ctypes as C, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tests.test_simt_kernels as T
from tests.util import synth
from auroralib.compression_b200 import _abi as A
def build(name,harness,macro,extra=()):
    src=open(os.path.join(T.CSRC,name+'.cu')).read(); inc=f'/tmp/asan_{name}.inc'; open(inc,'w').write(src[:src.index('// ---- kernel\n')])
    so=f'/tmp/libasan_{name}.so'
    subprocess.check_call(['g++','-O1','-g','-fsanitize=address','-fno-omit-frame-pointer','-std=c++17','-shared','-fPIC','-I',T.SIMT,'-I',T.CSRC,f'-DAURORA_REAL_COMMON="{os.path.join(T.CSRC,"common.cuh")}"',f'-DAURORA_REAL_STAGE="{os.path.join(T.CSRC,"stage.cuh")}"',f'-D{macro}="{inc}"',*extra,os.path.join(T.SIMT,harness),'-o',so])
    return C.CDLL(so)
class L: pass
lib=L()
lib.par=build('encode_lz_par','par_harness.cpp','PAR_DEVICE_INC').simt_encode_lz_par
lib.seq_flag=build('encode_lz','seq_harness.cpp','SEQ_DEVICE_INC').simt_encode_seq
lib.seq_byte=build('encode_bytelz','seq_harness.cpp','SEQ_DEVICE_INC',['-DSEQ_BYTELZ']).simt_encode_seq
for f in (lib.par,lib.seq_flag,lib.seq_byte): f.restype=C.c_int
bmp=open(os.path.join(T.ROOT, 'tests', 'golden', 'Test.bmp'), 'rb').read()
rng=np.random.default_rng(3)
raws=[bmp[:9000],bmp[:5],b'']+[synth(rng,int(k),i%5) for i,k in enumerate([1,4,33,1000,4096,6000,12000,70000])]+[bytes(40000)]
n=0
for fmt in T.PAR_FORMATS:
    for q in (0,8,12):
        for strat in (0,1):
            T.simt_encode(lib,fmt,raws if fmt!=A.FMT_BLZ else [r[::-1] for r in raws],q,strategy=strat,skew=int(rng.integers(0,16))); n+=len(raws)
    T.simt_encode(lib,fmt,[bmp[:20000]],3,caps=[100]); n+=1
    print('par',A.FORMAT_NAMES[fmt],'ok',flush=True)
for fmt in T.SEQ_FLAG_FORMATS+T.BYTE_FORMATS:
    for q in (0,12):
        T.simt_encode(lib,fmt,raws[:8],q,seq=True,skew=3); n+=8
    T.simt_encode(lib,fmt,[bmp[:20000]],3,caps=[100],seq=True); n+=1
    print('seq',A.FORMAT_NAMES[fmt],'ok',flush=True)
print('no AddressSanitizer report over',n,'buffers')
