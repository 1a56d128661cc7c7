"""Batch entry points over the C ABI: many independent blobs at once (the north-star's new API).

`BatchCodec` owns one aurora_ctx (one per process; one process per GPU under torchrun, or all visible
GPUs of the box when device_mask == 0).  Host-buffer calls go through aurora_*_batch (H2D + kernel + D2H
inside the call); device-resident calls take torch CUDA tensors and only launch.
"""
import ctypes as C

import numpy as np

from . import _abi, _lib


class AuroraError(RuntimeError):
    pass


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def pack(blobs, align=16):
    """Concatenate byte blobs at `align`-byte aligned offsets -> (base u8 array, off u64, len u64)."""
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
            base[int(off[i]):int(off[i]) + len(b)] = np.frombuffer(bytes(b), dtype=np.uint8) if not isinstance(b, np.ndarray) else b
    return base, off, ln


def layout(caps, align=16):
    caps = np.asarray(caps, dtype=np.uint64)
    padded = (caps + np.uint64(align - 1)) & ~np.uint64(align - 1)
    off = np.zeros(len(caps), dtype=np.uint64)
    if len(caps) > 1:
        off[1:] = np.cumsum(padded[:-1])
    return caps, off, (int(padded.sum()) if len(caps) else 0)


class BatchCodec:
    def __init__(self, device_mask=0):
        self._L = _lib.load()
        self._ctx = self._L.aurora_init(device_mask)
        if not self._ctx:
            raise AuroraError("aurora_init failed: " + self._L.aurora_last_error_string(None).decode() +
                              " (this engine has no CPU fallback; a B200 is required)")

    def close(self):
        if self._ctx:
            self._L.aurora_shutdown(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_count(self):
        return self._L.aurora_ctx_device_count(self._ctx)

    @property
    def kernel_launches(self):
        return int(self._L.aurora_kernel_launch_count(self._ctx))

    def _check(self, rc, what):
        if rc != _abi.OK:
            raise AuroraError(f"{what}: {_abi.STATUS_NAMES[rc]}: {self._L.aurora_last_error_string(self._ctx).decode()}")

    # ------------------------------------------------------------------ host buffers (packed numpy arrays)
    def decode_packed(self, fmt, base, off, ln, dst, doff, caps, opts=None):
        opts = opts or _abi.make_opts()
        n = len(off)
        out_len = np.zeros(n, dtype=np.uint64)
        consumed = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        rc = self._L.aurora_decode_batch(self._ctx, fmt, C.byref(opts), n, _ptr(base), _ptr(off), _ptr(ln), _ptr(dst),
                                         _ptr(doff), _ptr(caps), _ptr(out_len), _ptr(consumed), _ptr(status))
        self._check(rc, "aurora_decode_batch")
        return out_len, consumed, status

    def decode_batch(self, fmt, blobs, caps, opts=None):
        """-> (list of decoded bytes truncated to cap, out_len, consumed, status)"""
        base, off, ln = pack(blobs)
        caps, doff, total = layout(caps)
        dst = np.zeros(max(total, 16), dtype=np.uint8)
        out_len, consumed, status = self.decode_packed(fmt, base, off, ln, dst, doff, caps, opts)
        outs = [dst[int(doff[i]):int(doff[i]) + min(int(out_len[i]), int(caps[i]))].tobytes() for i in range(len(blobs))]
        return outs, out_len, consumed, status

    def decoded_size_batch(self, fmt, blobs, opts=None, size_scan=False):
        opts = opts or _abi.make_opts()
        base, off, ln = pack(blobs)
        n = len(blobs)
        out = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        rc = self._L.aurora_decoded_size_batch(self._ctx, fmt, C.byref(opts), n, _ptr(base), _ptr(off), _ptr(ln),
                                               1 if size_scan else 0, _ptr(out), _ptr(status))
        self._check(rc, "aurora_decoded_size_batch")
        return out, status

    def is_match_batch(self, fmt, blobs, opts=None):
        opts = opts or _abi.make_opts()
        base, off, ln = pack(blobs)
        n = len(blobs)
        m = np.zeros(n, dtype=np.uint8)
        rc = self._L.aurora_is_match_batch(self._ctx, fmt, C.byref(opts), n, _ptr(base), _ptr(off), _ptr(ln), _ptr(m))
        self._check(rc, "aurora_is_match_batch")
        return m.astype(bool)

    def scan_offsets(self, fmt, image, opts=None):
        """IsMatch at every byte offset of `image` -> bool array (the data-parallel `-scan`)."""
        opts = opts or _abi.make_opts()
        a = np.frombuffer(bytes(image), dtype=np.uint8) if not isinstance(image, np.ndarray) else image
        m = np.zeros(max(len(a), 1), dtype=np.uint8)
        if len(a):
            rc = self._L.aurora_scan_offsets(self._ctx, fmt, C.byref(opts), _ptr(a), len(a), _ptr(m))
            self._check(rc, "aurora_scan_offsets")
        return m[:len(a)].astype(bool)

    def encode_bound(self, fmt, raw_len):
        return int(self._L.aurora_encode_bound(fmt, raw_len))

    def encode_packed(self, fmt, base, off, ln, dst, doff, caps, opts=None):
        opts = opts or _abi.make_opts()
        n = len(off)
        out_len = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        rc = self._L.aurora_encode_batch(self._ctx, fmt, C.byref(opts), n, _ptr(base), _ptr(off), _ptr(ln), _ptr(dst),
                                         _ptr(doff), _ptr(caps), _ptr(out_len), _ptr(status))
        self._check(rc, "aurora_encode_batch")
        return out_len, status

    def encode_batch(self, fmt, blobs, opts=None):
        base, off, ln = pack(blobs)
        caps, doff, total = layout([self.encode_bound(fmt, len(b)) for b in blobs])
        dst = np.zeros(max(total, 16), dtype=np.uint8)
        out_len, status = self.encode_packed(fmt, base, off, ln, dst, doff, caps, opts)
        outs = [dst[int(doff[i]):int(doff[i]) + int(out_len[i])].tobytes() if status[i] == 0 else b"" for i in range(len(blobs))]
        return outs, status

    # ------------------------------------------------------------------ device-resident (torch CUDA tensors)
    def decode_device(self, fmt, d_src, d_off, d_len, d_dst, d_doff, d_cap, d_out_len, d_consumed, d_status,
                      opts=None, device=0, stream=None):
        """All arguments are torch CUDA tensors (uint8 / int64 / int32).  Asynchronous on `stream`."""
        opts = opts or _abi.make_opts()
        rc = self._L.aurora_decode_batch_device(
            self._ctx, device, fmt, C.byref(opts), d_off.numel(), d_src.data_ptr(), d_src.numel(), d_off.data_ptr(),
            d_len.data_ptr(), d_dst.data_ptr(), d_doff.data_ptr(), d_cap.data_ptr(), d_out_len.data_ptr(),
            d_consumed.data_ptr(), d_status.data_ptr(), stream)
        self._check(rc, "aurora_decode_batch_device")

    def encode_device(self, fmt, d_src, d_off, d_len, d_dst, d_doff, d_cap, d_out_len, d_status, opts=None, device=0,
                      stream=None):
        opts = opts or _abi.make_opts()
        rc = self._L.aurora_encode_batch_device(
            self._ctx, device, fmt, C.byref(opts), d_off.numel(), d_src.data_ptr(), d_src.numel(), d_off.data_ptr(),
            d_len.data_ptr(), d_dst.data_ptr(), d_doff.data_ptr(), d_cap.data_ptr(), d_out_len.data_ptr(),
            d_status.data_ptr(), stream)
        self._check(rc, "aurora_encode_batch_device")


_DEFAULT = None


def default_codec():
    """Process-wide BatchCodec (all visible devices)."""
    global _DEFAULT
    if _DEFAULT is None:
        _DEFAULT = BatchCodec(0)
    return _DEFAULT
