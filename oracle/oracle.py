"""ctypes binding of the CPU ORACLE (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package (auroralib.compression_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from auroralib.compression_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        vp = C.c_void_p
        L.ora_decode_batch.argtypes = [C.c_int, C.POINTER(_abi.CodecOpts), C.c_size_t, vp, vp, vp, vp, vp, vp, vp,
                                       vp, vp, C.c_int]
        L.ora_encode_batch.argtypes = [C.c_int, C.POINTER(_abi.CodecOpts), C.c_size_t, vp, vp, vp, vp, vp, vp, vp,
                                       vp, C.c_int]
        L.ora_decoded_size_batch.argtypes = [C.c_int, C.POINTER(_abi.CodecOpts), C.c_size_t, vp, vp, vp, C.c_int,
                                             vp, vp]
        L.ora_is_match_batch.argtypes = [C.c_int, C.POINTER(_abi.CodecOpts), C.c_size_t, vp, vp, vp, vp]
        L.ora_xxh64.restype = C.c_uint64
        L.ora_xxh64.argtypes = [vp, C.c_size_t, C.c_uint64]
        L.ora_xxh32.restype = C.c_uint32
        L.ora_xxh32.argtypes = [vp, C.c_size_t, C.c_uint32]
        L.ora_crc32c.restype = C.c_uint32
        L.ora_crc32c.argtypes = [vp, C.c_size_t]
        L.ora_hardware_threads.restype = C.c_int
        _LIB = L
    return _LIB


def hardware_threads():
    return int(lib().ora_hardware_threads())


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _as_u8(b):
    if isinstance(b, np.ndarray):
        return b.view(np.uint8).reshape(-1)
    return np.frombuffer(bytes(b), dtype=np.uint8)


def pack(blobs, align=16):
    """Concatenate byte blobs at aligned offsets -> (base u8 array, off u64, len u64)."""
    n = len(blobs)
    ln = np.array([len(b) for b in blobs], dtype=np.uint64)
    padded = (ln + np.uint64(align - 1)) & ~np.uint64(align - 1)
    off = np.zeros(n, dtype=np.uint64)
    if n > 1:
        off[1:] = np.cumsum(padded[:-1])
    total = int(padded.sum()) if n else 0
    base = np.zeros(max(total, 16), dtype=np.uint8)
    for i, b in enumerate(blobs):
        if len(b):
            base[int(off[i]):int(off[i]) + len(b)] = _as_u8(b)
    return base, off, ln


def layout(caps, align=16):
    caps = np.asarray(caps, dtype=np.uint64)
    padded = (caps + np.uint64(align - 1)) & ~np.uint64(align - 1)
    off = np.zeros(len(caps), dtype=np.uint64)
    if len(caps) > 1:
        off[1:] = np.cumsum(padded[:-1])
    return caps, off, int(padded.sum()) if len(caps) else 0


def decode_packed(fmt, base, off, ln, dst, doff, caps, opts=None, threads=0):
    """Raw batch call on already packed numpy arrays.  Returns (out_len, consumed, status)."""
    opts = opts or _abi.make_opts()
    n = len(off)
    out_len = np.zeros(n, dtype=np.uint64)
    consumed = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    lib().ora_decode_batch(fmt, C.byref(opts), n, _ptr(base), _ptr(off), _ptr(ln), _ptr(dst), _ptr(doff), _ptr(caps),
                           _ptr(out_len), _ptr(consumed), _ptr(status), threads)
    return out_len, consumed, status


def decode_batch(fmt, blobs, caps, opts=None, threads=0):
    """Returns (list of output bytes (truncated to cap), out_len, consumed, status)."""
    base, off, ln = pack(blobs)
    caps, doff, total = layout(caps)
    dst = np.zeros(max(total, 16), dtype=np.uint8)
    out_len, consumed, status = decode_packed(fmt, base, off, ln, dst, doff, caps, opts, threads)
    outs = [dst[int(doff[i]):int(doff[i]) + min(int(out_len[i]), int(caps[i]))].tobytes() for i in range(len(blobs))]
    return outs, out_len, consumed, status


def decode(fmt, blob, cap, opts=None):
    outs, out_len, consumed, status = decode_batch(fmt, [blob], [cap], opts, threads=1)
    return outs[0], int(out_len[0]), int(consumed[0]), int(status[0])


def encode_packed(fmt, base, off, ln, dst, doff, caps, opts=None, threads=0):
    opts = opts or _abi.make_opts()
    n = len(off)
    out_len = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    lib().ora_encode_batch(fmt, C.byref(opts), n, _ptr(base), _ptr(off), _ptr(ln), _ptr(dst), _ptr(doff), _ptr(caps),
                           _ptr(out_len), _ptr(status), threads)
    return out_len, status


def encode_batch(fmt, blobs, opts=None, threads=0):
    base, off, ln = pack(blobs)
    caps, doff, total = layout([2 * len(b) + 1024 for b in blobs])
    dst = np.zeros(max(total, 16), dtype=np.uint8)
    out_len, status = encode_packed(fmt, base, off, ln, dst, doff, caps, opts, threads)
    outs = [dst[int(doff[i]):int(doff[i]) + int(out_len[i])].tobytes() if status[i] == 0 else b""
            for i in range(len(blobs))]
    return outs, status


def encode(fmt, blob, opts=None):
    outs, status = encode_batch(fmt, [blob], opts, threads=1)
    return outs[0], int(status[0])


def decoded_size(fmt, blob, opts=None, size_scan=0):
    opts = opts or _abi.make_opts()
    base, off, ln = pack([blob])
    out = np.zeros(1, dtype=np.uint64)
    st = np.zeros(1, dtype=np.int32)
    lib().ora_decoded_size_batch(fmt, C.byref(opts), 1, _ptr(base), _ptr(off), _ptr(ln), size_scan, _ptr(out), _ptr(st))
    return int(out[0]), int(st[0])


def is_match(fmt, blob, opts=None):
    opts = opts or _abi.make_opts()
    base, off, ln = pack([blob])
    m = np.zeros(1, dtype=np.uint8)
    lib().ora_is_match_batch(fmt, C.byref(opts), 1, _ptr(base), _ptr(off), _ptr(ln), _ptr(m))
    return bool(m[0])


def xxh64(data, seed=0):
    a = _as_u8(data)
    return int(lib().ora_xxh64(_ptr(a) if len(a) else None, len(a), seed))


def xxh32(data, seed=0):
    a = _as_u8(data)
    return int(lib().ora_xxh32(_ptr(a) if len(a) else None, len(a), seed))


def crc32c(data):
    a = _as_u8(data)
    return int(lib().ora_crc32c(_ptr(a) if len(a) else None, len(a)))
