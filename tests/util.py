"""Shared helpers for the parity tests: seeded synthetic inputs and corruption operators."""
import numpy as np

from auroralib.compression_b200 import _abi as A

ALL_FORMATS = [A.FMT_YAZ0, A.FMT_YAZ1, A.FMT_YAY0, A.FMT_MIO0, A.FMT_LZ10, A.FMT_LZ11, A.FMT_LZSS, A.FMT_LZ4,
               A.FMT_LZ4_LEGACY, A.FMT_LZ4_BLOCK, A.FMT_LZO, A.FMT_SNAPPY, A.FMT_SNAPPY_BLOCK, A.FMT_PRS, A.FMT_LZHUDSON, A.FMT_LZ40, A.FMT_LZ60, A.FMT_SMSR00, A.FMT_BLZ]
SIZED_FORMATS = [A.FMT_YAZ0, A.FMT_YAZ1, A.FMT_YAY0, A.FMT_MIO0, A.FMT_LZ10, A.FMT_LZ11, A.FMT_LZSS, A.FMT_LZHUDSON, A.FMT_LZ40, A.FMT_LZ60, A.FMT_SMSR00, A.FMT_BLZ]


def fmt_id(f):
    return A.FORMAT_NAMES[f]


def end_position(fmt, comp):
    """source.Position after a successful Decompress: the end of the stream, except BLZ, which reads its codes from the middle
    of the stream and leaves the position in front of the padding and the footer (BLZ.cs:56-62)."""
    return len(comp) - comp[-5] if fmt == A.FMT_BLZ else len(comp)


def synth(rng, n, kind):
    """Small asset-like buffers: 0 tiles, 1 u16 tilemap runs, 2 mixed entropy, 3 zeros/runs, 4 random."""
    if n == 0:
        return b""
    if kind == 0:
        tiles = rng.integers(0, 256, size=(16, 32), dtype=np.uint8)
        picks = rng.integers(0, 16, size=(n + 31) // 32)
        return tiles[picks].reshape(-1)[:n].tobytes()
    if kind == 1:
        out = np.zeros((n + 1) // 2, dtype=np.uint16)
        i = 0
        while i < len(out):
            run = int(rng.geometric(1 / 24))
            base = int(rng.integers(0, 1024))
            inc = int(rng.integers(0, 2))
            k = min(run, len(out) - i)
            out[i:i + k] = (base + inc * np.arange(k)) & 0x3FF | (int(rng.integers(0, 4)) << 10)
            i += k
        return out.tobytes()[:n]
    if kind == 2:
        parts = []
        total = 0
        while total < n:
            seg = int(rng.integers(16, 600))
            t = int(rng.integers(0, 3))
            if t == 0:
                p = rng.integers(0, 256, size=seg, dtype=np.uint8).tobytes()
            elif t == 1:
                p = bytes(seg)
            else:
                m = rng.integers(0, 256, size=int(rng.integers(1, 40)), dtype=np.uint8).tobytes()
                p = (m * (seg // len(m) + 1))[:seg]
            parts.append(p)
            total += seg
        return b"".join(parts)[:n]
    if kind == 3:
        out = bytearray(n)
        i = 0
        while i < n:
            run = int(rng.integers(1, 700))
            v = int(rng.integers(0, 4))
            out[i:i + run] = bytes([v]) * min(run, n - i)
            i += run
        return bytes(out[:n])
    return rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()


def corrupt(rng, blob, mode):
    """0 truncate, 1 flip a byte, 2 append garbage, 3 empty, 4 cut inside the header."""
    b = bytearray(blob)
    if mode == 0 and len(b) > 1:
        return bytes(b[:int(rng.integers(1, len(b)))])
    if mode == 1 and len(b) > 0:
        i = int(rng.integers(0, len(b)))
        b[i] ^= int(rng.integers(1, 256))
        return bytes(b)
    if mode == 2:
        return bytes(b) + rng.integers(0, 256, size=int(rng.integers(1, 40)), dtype=np.uint8).tobytes()
    if mode == 3:
        return b""
    if mode == 4:
        return bytes(b[:int(rng.integers(0, min(17, len(b) + 1)))])
    return bytes(b)
