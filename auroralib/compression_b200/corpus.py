"""Synthetic game-asset-like corpora for the benchmark and the tests.

Classes (all produce `size`-byte streams; batched torch ops so the full configs can be generated on the GPU):
  T  4-bpp tile sheets: 256 random 32-byte tiles followed by tile picks with Zipf(1.2) reuse
  M  u16 tilemaps: runs (geometric, mean 24) of constant or +1-incrementing 10-bit tile ids with 2 flag bits
  X  mixed entropy: alternating segments (256-4096 B) of random bytes, zeros and repeated 3-64-byte motifs
  B  32-bpp bitmaps: smooth gradients + flat regions + 10 % noise pixels (statistically like Test.bmp)
The mix of config C2 is 40 % T, 30 % M, 30 % X; stream i of a batch has class CLASS_ORDER[i % 10].
Seeds: a torch.Generator seeded with (seed + class ordinal); the data is synthetic and is never shipped.
"""
import math

import torch

CLASS_ORDER = "TTTTMMMXXX"   # 40 / 30 / 30
_CHUNK = 1024


def _gen_T(S, size, g, dev):
    tiles = torch.randint(0, 256, (S, 256, 32), dtype=torch.uint8, generator=g, device=dev)
    head = tiles.reshape(S, 8192)
    if size <= 8192:
        return head[:, :size].contiguous()
    npick = (size - 8192 + 31) // 32
    w = 1.0 / torch.arange(1, 257, dtype=torch.float64, device=dev) ** 1.2
    cdf = (torch.cumsum(w, 0) / w.sum()).to(torch.float32)
    u = torch.rand((S, npick), generator=g, device=dev)
    rank = torch.searchsorted(cdf, u).clamp_(max=255)
    perm = torch.argsort(torch.rand((S, 256), generator=g, device=dev), dim=1)
    idx = torch.gather(perm, 1, rank)
    picks = torch.gather(tiles, 1, idx.unsqueeze(-1).expand(-1, -1, 32)).reshape(S, npick * 32)
    return torch.cat([head, picks], 1)[:, :size].contiguous()


def _gen_M(S, size, g, dev):
    n16 = (size + 1) // 2
    K = int(n16 / 24 * 1.6) + 32
    u = torch.rand((S, K), generator=g, device=dev).clamp_(min=1e-7)
    lens = (torch.log(u) / math.log(1 - 1 / 24)).floor().to(torch.int64) + 1
    ends = torch.cumsum(lens, 1)
    starts = ends - lens
    pos = torch.arange(n16, device=dev).unsqueeze(0).expand(S, -1).contiguous()
    run = torch.searchsorted(ends, pos, right=True).clamp_(max=K - 1)
    base = torch.randint(0, 1024, (S, K), generator=g, device=dev)
    inc = torch.randint(0, 2, (S, K), generator=g, device=dev)
    flags = torch.randint(0, 4, (S, K), generator=g, device=dev)
    v = (torch.gather(base, 1, run) + torch.gather(inc, 1, run) * (pos - torch.gather(starts, 1, run))) & 0x3FF
    v = v | (torch.gather(flags, 1, run) << 10)
    lo = (v & 0xFF).to(torch.uint8)
    hi = (v >> 8).to(torch.uint8)
    return torch.stack([lo, hi], 2).reshape(S, n16 * 2)[:, :size].contiguous()


def _gen_X(S, size, g, dev):
    K = size // 256 + 2
    lens = torch.randint(256, 4097, (S, K), generator=g, device=dev)
    ends = torch.cumsum(lens, 1)
    starts = ends - lens
    pos = torch.arange(size, device=dev).unsqueeze(0).expand(S, -1).contiguous()
    seg = torch.searchsorted(ends, pos, right=True).clamp_(max=K - 1)
    kind = torch.gather(torch.randint(0, 3, (S, K), generator=g, device=dev), 1, seg)
    rnd = torch.randint(0, 256, (S, size), dtype=torch.uint8, generator=g, device=dev)
    motif = torch.randint(0, 256, (S, K * 64), dtype=torch.uint8, generator=g, device=dev)
    mlen = torch.randint(3, 65, (S, K), generator=g, device=dev)
    off = (pos - torch.gather(starts, 1, seg)) % torch.gather(mlen, 1, seg)
    mot = torch.gather(motif, 1, seg * 64 + off)
    out = torch.where(kind == 0, rnd, torch.where(kind == 1, torch.zeros_like(rnd), mot))
    return out.contiguous()


def _gen_B(S, size, g, dev):
    npx = (size + 3) // 4
    width = 512
    pos = torch.arange(npx, device=dev).unsqueeze(0).expand(S, -1)
    x, y = pos % width, pos // width
    # per-stream gradient parameters and a coarse flat-region map (64x8-pixel cells)
    gx = torch.randint(0, 4, (S, 4, 1), generator=g, device=dev)
    gy = torch.randint(0, 4, (S, 4, 1), generator=g, device=dev)
    c0 = torch.randint(0, 256, (S, 4, 1), generator=g, device=dev)
    chan = (c0 + (gx * x.unsqueeze(1)) // 8 + (gy * y.unsqueeze(1)) // 4) & 0xFF
    chan[:, 3, :] = 0xFF
    cells = (npx // (64 * 8)) + 2
    flat = torch.randint(0, 3, (S, cells), generator=g, device=dev) == 0
    cell = ((y // 8) * (width // 64) + x // 64).clamp(max=cells - 1)
    is_flat = torch.gather(flat, 1, cell)
    flat_col = torch.randint(0, 256, (S, 4, cells), generator=g, device=dev)
    fc = torch.gather(flat_col, 2, cell.unsqueeze(1).expand(-1, 4, -1))
    chan = torch.where(is_flat.unsqueeze(1), fc, chan)
    noise = torch.rand((S, npx), generator=g, device=dev) < 0.10
    nz = torch.randint(0, 256, (S, 4, npx), generator=g, device=dev)
    chan = torch.where(noise.unsqueeze(1), nz, chan)
    return chan.permute(0, 2, 1).reshape(S, npx * 4)[:, :size].to(torch.uint8).contiguous()


_GEN = {"T": _gen_T, "M": _gen_M, "X": _gen_X, "B": _gen_B}


def generate(cls, n_streams, size, seed=0xA0120000, device="cpu"):
    """(n_streams, size) uint8 tensor of class `cls` on `device`."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed + "TMXB".index(cls))
    parts = []
    chunk = max(1, min(_CHUNK, (1 << 26) // max(size, 1)))   # bound the int64 index temporaries (~8 x chunk x size bytes)
    for s in range(0, n_streams, chunk):
        parts.append(_GEN[cls](min(chunk, n_streams - s), size, g, dev))
    if not parts:
        return torch.zeros((0, size), dtype=torch.uint8, device=dev)
    return torch.cat(parts, 0)


def generate_mix(n_streams, size, seed=0xA0120000, device="cpu", order=CLASS_ORDER):
    """C2 mix: stream i has class order[i % len(order)].  Returns ((n_streams, size) uint8, list of class letters)."""
    classes = [order[i % len(order)] for i in range(n_streams)]
    out = torch.empty((n_streams, size), dtype=torch.uint8, device=device)
    for c in sorted(set(classes)):
        idx = torch.tensor([i for i, k in enumerate(classes) if k == c], device=device)
        out[idx] = generate(c, len(idx), size, seed, device)
    return out, classes
