// decode_bytelz.cu — batched decoder for the byte-tagged LZ family with 64 KiB (8 KiB for PRS) windows:
// LZ4 (raw block, legacy frame, v1 frame, skippable frames), Snappy (raw block and framing format), LZO1X
// and SEGA PRS.  One compressed stream per warp.
//
// Reference semantics restated on the device (paths under /root/reference/src):
//   LZ4 block      AuroraLib.Compression/Formats/Common/LZ4.cs:176-200, ReadExtension :241-252
//   LZ4 containers LZ4.cs:50-111, LZ4.Frame.cs:107-174, LZ4.FrameDescriptor.cs:18-26
//   Snappy         Formats/Common/Snappy.cs:39-68 (framing), :109-122 (varint), :205-250 (block)
//   LZO            Formats/Common/LZO.cs:49-139, ReadExtendedInt :252-262
//   PRS            AuroraLib.Compression.Sega/Sega/PRS.cs:42-102, :161-218
//   window         AuroraLib.Compression/IO/LzWindows.cs:72-100 (BackCopy), :124-135 (CopyFrom)
//
// Design: the token walk of these formats is inherently sequential (every token's position depends on
// the previous token's extension bytes), so the warp walks tokens uniformly — every lane evaluates the
// same tag from the TMA-staged shared-memory copy of the input — and spends its 32 lanes on the two
// things that are parallel: literal runs (input ring -> HBM) and match copies (HBM -> HBM through L1/L2,
// out[dst+i] = out[dst-d + (i mod d)]).  The window is up to 64 KiB, so unlike the flag-LZ family the
// output is written straight to global memory and back-references are served by L1/L2.
#include "common.cuh"
#include "stage.cuh"

namespace aurora {

namespace {

enum Kind { B_LZ4 = 0, B_LZ4_BLOCK = 1, B_SNAPPY = 2, B_SNAPPY_BLOCK = 3, B_LZO = 4, B_PRS = 5 };

constexpr int kORing = 2048;            // per-warp output ring: the most recent decoded bytes stay in shared memory
//           // per-warp output ring: the most recent decoded bytes stay in shared memory
constexpr int kORingMask = kORing - 1;
constexpr int kOKeep = 1024;            // back-references reaching at most this far are served from the ring
constexpr int kOPiece = 512;            // bytes per copy step / drain step
constexpr int kSmemPerWarp = kORing + kInStage + 64;
constexpr int kWarpsPerBlock = 23;      // x2 blocks per SM = 46 warps (latency bound: occupancy beats registers; measured)
                                        // leaves ~70 KiB of L1 for the far back-references (measured best: profiles/)

// Output cursor.  Decoded bytes are produced into an 8 KiB shared-memory ring (8 KiB aligned, so the wrapped address
// is one LOP3) and drained to HBM in 512-byte steps with 16-byte vector stores; a back-reference is served from the
// ring when it reaches at most kOKeep bytes back and from global memory (already drained) otherwise, so the common
// short-distance copies never wait for an HBM/L2 round trip.  `win_base` is the output position where the current
// LzWindows instance was created: references before it read the (zero) pre-history of a fresh ring.
struct GOut {
    uint8_t* dst;
    uint64_t cap;
    uint32_t rb;         // shared address of the ring
    uint32_t written;
    uint32_t flushed;    // multiple of kOPiece
    uint32_t win_base;
    uint32_t ring_len;   // the reference's window size (BackCopy(0, n) semantics)
    bool size_only;
    bool aligned;

    __device__ __forceinline__ void drain() {
        const uint32_t lane = lane_id();
        if (written - flushed < uint32_t(kOPiece)) return;
        while (written - flushed >= uint32_t(kOPiece)) {
            const uint64_t end = uint64_t(flushed) + kOPiece;
            if (end <= cap && aligned) {
                const uint4 v = lds_u128(((flushed + lane * 16) & kORingMask) | rb);
                *reinterpret_cast<uint4*>(dst + flushed + lane * 16) = v;
            } else {
                for (uint32_t i = lane; i < uint32_t(kOPiece); i += 32)
                    if (uint64_t(flushed) + i < cap) dst[flushed + i] = uint8_t(lds_u8(((flushed + i) & kORingMask) | rb));
            }
            flushed += kOPiece;
        }
        // a far back-reference of the NEXT copy may read these bytes back from global memory on another lane: order the
        // stores before it (lock-step execution hid the missing order on the GPU; the lane emulation of tests/simt did not)
        __syncwarp();
    }
    __device__ __forceinline__ void finish() {
        const uint32_t lane = lane_id();
        drain();
        for (uint32_t p = flushed + lane; p < written; p += 32)
            if (uint64_t(p) < cap) dst[p] = uint8_t(lds_u8((p & kORingMask) | rb));
        flushed = written;
        __syncwarp();   // the next stream (or PRS's second attempt) rewrites these ring slots (racecheck: write-after-read)
    }
    // make everything decoded so far visible in global memory without disturbing the 512-byte drain cadence
    __device__ __forceinline__ void sync_tail() {
        const uint32_t lane = lane_id();
        drain();
        for (uint32_t p = flushed + lane; p < written; p += 32)
            if (uint64_t(p) < cap) dst[p] = uint8_t(lds_u8((p & kORingMask) | rb));
        __syncwarp();
    }
    // LzWindows.Write / CopyFrom: `len` bytes from the staged input at relative position ipos
    __device__ __forceinline__ void lit_copy(InStream& in, uint32_t ipos, uint32_t len) {
        const uint32_t lane = lane_id();
        uint32_t done = 0;
        while (done < len) {
            const uint32_t piece = min(len - done, uint32_t(kOPiece));
            in.ensure(ipos + done, piece);
            const uint32_t wa = smem_u32(in.window(ipos + done));
            for (uint32_t i = lane; i < piece; i += 32) sts_u8(((written + i) & kORingMask) | rb, lds_u8(wa + i));
            written += piece;
            done += piece;
            __syncwarp();
            drain();
        }
    }
    // LzWindows.BackCopy(distance, length): out[o] = out[o - d], realised as out[dst+i] = out[dst-d + (i mod d)] per piece
    __device__ __forceinline__ void match_copy(uint32_t d, uint32_t len) {
        const uint32_t lane = lane_id();
        if (d == 0) d = ring_len;   // BackCopy(0, n) re-reads the ring slot it writes: one window back
        if (len <= 32) {
            // short match (the common case): one warp step, no piece loop
            const uint32_t base = written;
            const uint32_t r = d < len ? c_rcp.v[d & 511] : 0u;
            if (lane < len) {
                const int32_t s = int32_t(base) - int32_t(d) + int32_t(lane - ((lane * r) >> 20) * d);
                uint32_t v = 0;
                if (s >= int32_t(win_base)) {
                    if (d <= uint32_t(kOKeep)) v = lds_u8((uint32_t(s) & kORingMask) | rb);
                    else if (uint64_t(s) < cap) v = dst[s];
                }
                sts_u8(((base + lane) & kORingMask) | rb, v);
            }
            written = base + len;
            __syncwarp();
            drain();
            return;
        }
        uint32_t done = 0;
        while (done < len) {
            const uint32_t piece = min(len - done, uint32_t(kOPiece));
            const uint32_t base = written;
            const int32_t srcp = int32_t(base) - int32_t(d);
            const bool wrap = d < piece;                      // then d < 512: reciprocal table applies
            const uint32_t r = wrap ? c_rcp.v[d & 511] : 0u;
            const bool near = d <= uint32_t(kOKeep);
            for (uint32_t i = lane; i < piece; i += 32) {
                const uint32_t off = i - ((i * r) >> 20) * d;   // i mod d when wrapping, i otherwise (r == 0)
                const int32_t s = srcp + int32_t(off);
                uint32_t v = 0;
                if (s >= int32_t(win_base)) {
                    if (near) v = lds_u8((uint32_t(s) & kORingMask) | rb);
                    else if (uint64_t(s) < cap) v = dst[s];
                }
                sts_u8(((base + i) & kORingMask) | rb, v);
            }
            written += piece;
            done += piece;
            __syncwarp();
            drain();
        }
    }
    // One short LZ4 sequence (lit + mlen <= 32) as a single warp step: lanes [0, lit) copy the literals from the staged
    // input, lanes [lit, lit + mlen) copy the match, when the match source lies entirely before this sequence's
    // literals (otherwise literals first, then the match).  The caller has made the literal bytes readable.
    __device__ __forceinline__ void seq_copy_small(InStream& in, uint32_t ipos, uint32_t lit, uint32_t d, uint32_t mlen) {
        const uint32_t lane = lane_id();
        if (d == 0) d = ring_len;
        const uint32_t base = written, mbase = base + lit;
        const int32_t srcp = int32_t(mbase) - int32_t(d);
        const uint32_t wa = smem_u32(in.window(ipos));
        const bool near = d <= uint32_t(kOKeep);
        const uint32_t r = d < mlen ? c_rcp.v[d & 511] : 0u;
        const bool indep = srcp + int32_t(min(mlen, d)) <= int32_t(base);
        if (!indep) {
            if (lane < lit) sts_u8(((base + lane) & kORingMask) | rb, lds_u8(wa + lane));
            __syncwarp();
        }
        const uint32_t first = indep ? 0u : lit;
        if (lane >= first && lane < lit + mlen) {
            uint32_t v = 0;
            if (lane < lit) {
                v = lds_u8(wa + lane);
            } else {
                const uint32_t i = lane - lit;
                const int32_t s = srcp + int32_t(i - ((i * r) >> 20) * d);
                if (s >= int32_t(win_base)) {
                    if (near) v = lds_u8((uint32_t(s) & kORingMask) | rb);
                    else if (uint64_t(s) < cap) v = dst[s];
                }
            }
            sts_u8(((base + lane) & kORingMask) | rb, v);
        }
        written += lit + mlen;
        __syncwarp();
        drain();
    }
    __device__ __forceinline__ void new_window() { win_base = written; }
    // LzWindows.CopyFrom (IO/LzWindows.cs:124-135) reads ring-sized pieces with ReadExactly: on a truncated
    // input only the pieces that could be read completely are committed.  Returns false on truncation.
    __device__ __forceinline__ bool copy_from(InStream& in, uint32_t ipos, uint32_t len, uint32_t avail) {
        if (len <= avail) {
            lit_copy(in, ipos, len);
            return true;
        }
        uint32_t done = 0;
        for (;;) {
            const uint32_t rp = (written - win_base) & (ring_len - 1);
            const uint32_t piece = min(len - done, ring_len - rp);
            if (piece > avail - done) break;
            lit_copy(in, ipos + done, piece);
            done += piece;
        }
        return false;
    }
};

struct Res {
    int status;
    uint32_t consumed;
};

// ------------------------------------------------------------------------------------------------ LZ4
// LZ4.cs:241-252 on the staged input; returns false when the block is exhausted (IndexOutOfRange)
__device__ __forceinline__ bool lz4_ext(InStream& in, uint32_t& sp, uint32_t end, uint32_t& length) {
    if (length == 15) {
        uint32_t b;
        do {
            if (sp >= end) return false;
            in.ensure(sp, 8);
            b = in.at(sp++);
            length += b;
        } while (b == 255);
    }
    return true;
}

// Commit of an element-per-lane batch (shared by the LZ4 and Snappy front ends): lane k < n holds one element = `lit`
// literal bytes at shared address `la` followed by a match of `mlen` bytes at distance `d` (either part may be empty); the
// batch decodes `cum` <= kBatchOut bytes in total.  Order of the parts is stream order: a warp scan places every element,
// literals and far matches (source drained to HBM before the batch: independent of everything in it) are stored first,
// then the remaining matches replay in stream order from the output ring.
constexpr uint32_t kBatchOut = 1024;
template <bool kLitAfter = false>   // LZO: an element is a match FOLLOWED by its 0..3 trailing literals
__device__ __forceinline__ void batch_commit(GOut& out, const uint32_t n, const uint32_t cum, const uint32_t lit, const uint32_t la,
                                             const uint32_t mlen, const uint32_t d) {
    const uint32_t lane = lane_id();
    const uint32_t incl = warp_incl_scan(lit + mlen);
    const uint32_t base = out.written;
    const uint32_t epos = base + incl - (lit + mlen);
    const uint32_t lpos = kLitAfter ? epos + mlen : epos, mpos = kLitAfter ? epos : epos + lit;
    const uint32_t dd = d ? d : out.ring_len;
    const uint32_t rb = out.rb;
    // ---- 3. literals: short runs one lane per sequence, long runs (15..269 bytes) warp-cooperatively
    {
        const bool shortl = lit <= 14;
        const uint32_t maxlit = __reduce_max_sync(kFull, shortl ? lit : 0u);
        for (uint32_t j = 0; j < maxlit; j++)
            if (shortl && j < lit) sts_u8(((lpos + j) & kORingMask) | rb, lds_u8(la + j));
        uint32_t lm = __ballot_sync(kFull, !shortl);
        while (lm) {
            const int k = __ffs(lm) - 1;
            lm &= lm - 1;
            const uint32_t kp = __shfl_sync(kFull, lpos, k), kl = __shfl_sync(kFull, lit, k), ka = __shfl_sync(kFull, la, k);
            for (uint32_t i = lane; i < kl; i += 32) sts_u8(((kp + i) & kORingMask) | rb, lds_u8(ka + i));
        }
    }
    // ---- 4. far matches: the source is in HBM already
    const bool has = lane < n && mlen > 0;
    const bool far = has && int32_t(mpos - dd + mlen) <= int32_t(out.flushed) && dd >= mlen;
    uint32_t fm = __ballot_sync(kFull, far);
    while (fm) {
        uint32_t v[4], dp[4], ln[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            v[t] = 0;
            ln[t] = 0;
            dp[t] = 0;
            if (fm) {
                const int k = __ffs(fm) - 1;
                fm &= fm - 1;
                const uint32_t kp = __shfl_sync(kFull, mpos, k), kl = __shfl_sync(kFull, mlen, k), kd = __shfl_sync(kFull, dd, k);
                dp[t] = kp;
                ln[t] = kl;
                const int32_t sidx = int32_t(kp) - int32_t(kd) + int32_t(lane);
                if (lane < kl && sidx >= int32_t(out.win_base) && uint64_t(sidx) < out.cap) v[t] = out.dst[sidx];
                if (kl > 32) {   // the rest of a long match right away (rare)
                    for (uint32_t i = lane + 32; i < kl; i += 32) {
                        const int32_t s2 = int32_t(kp) - int32_t(kd) + int32_t(i);
                        uint32_t w = 0;
                        if (s2 >= int32_t(out.win_base) && uint64_t(s2) < out.cap) w = out.dst[s2];
                        sts_u8(((kp + i) & kORingMask) | rb, w);
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t < 4; t++)
            if (lane < ln[t]) sts_u8(((dp[t] + lane) & kORingMask) | rb, v[t]);
    }
    // ---- 5. the other matches, in stream order, from the ring
    uint32_t nm = __ballot_sync(kFull, has && !far);
    while (nm) {
        const int k = __ffs(nm) - 1;
        nm &= nm - 1;
        const uint32_t kp = __shfl_sync(kFull, mpos, k), kl = __shfl_sync(kFull, mlen, k), kd = __shfl_sync(kFull, dd, k);
        __syncwarp();
        const uint32_t rr = kd < kl ? c_rcp.v[kd & 511] : 0u;   // kd < kl <= 273: the reciprocal table applies
        for (uint32_t i = lane; i < kl; i += 32) {
            const int32_t sidx = int32_t(kp) - int32_t(kd) + int32_t(i - ((i * rr) >> 20) * kd);
            uint32_t w = 0;
            if (sidx >= int32_t(out.win_base)) w = lds_u8((uint32_t(sidx) & kORingMask) | rb);
            sts_u8(((kp + i) & kORingMask) | rb, w);
        }
    }
    out.written = base + cum;
    __syncwarp();
    out.drain();
}

// Sequence-per-lane batch (LZ4.cs:176-200 for up to 32 consecutive "regular" sequences: at most one literal-length and one
// match-length extension byte, the match present inside the block).  The token walk of LZ4 is uniform work —
// 32 lanes computing the same scalars — and the kernel is issue bound, so everything that can be is moved to one lane per
// sequence:
//   1. chain (uniform, ~16 instructions per sequence): only the sequence starts — token, literal count, extension byte;
//   2. lane k decodes sequence k (literal source, distance, match length); one warp scan places all of them;
//   3. literals: every lane copies its own (<= 14) literal bytes, all sequences at once; longer runs warp-cooperatively;
//   4. far matches — the source was drained to HBM before the batch began, so it depends on nothing in the batch —
//      run four at a time with all four global loads issued before the stores (four L2 round trips in flight);
//   5. the other matches replay in stream order from the output ring (their sources are at most ~800 bytes behind the
//      batch: the ring still holds them), periodic when distance < length.
// A batch decodes at most kBatchOut = 1024 bytes: the last byte it writes maps to the ring slot 2048 below it, which is
// older than anything a match of the batch may still read (>= batch start - 1024).
// Returns the number of sequences consumed (0: the sequence at sp is not regular, the caller decodes it alone).
__device__ __forceinline__ uint32_t lz4_batch32(InStream& in, GOut& out, uint32_t& sp, const uint32_t end) {
    const uint32_t lane = lane_id();
    in.ensure(sp, kInMirror - 16);
    const uint32_t wa = smem_u32(in.window(sp));
    const uint32_t blk = end - sp;   // bytes of the block from sp
    // ---- 1. chain of sequence starts
    constexpr uint32_t kWin = kInMirror - 16;   // contiguous staged bytes from sp
    uint32_t r = 0, cum = 0, n = 0, myr = 0;
#pragma unroll 1
    for (uint32_t k = 0; k < 32; k++) {
        if (r + 2 > blk || r + 2 > kWin) break;
        const uint32_t token = lds_u8(wa + r);
        uint32_t lit = token >> 4, ml = token & 15;
        uint32_t q = r + 1;
        if (lit == 15) {   // one literal-length extension byte (runs of 15..269 literals)
            const uint32_t e = lds_u8(wa + q);
            if (e == 255) break;
            lit += e;
            q++;
        }
        q += lit + 2;                        // behind the distance
        if (q > blk || q + 1 > kWin) break;  // no match inside the block, or the sequence leaves the staged window
        if (ml == 15) {
            if (q >= blk) break;
            const uint32_t e = lds_u8(wa + q);
            if (e == 255) break;             // a second extension byte follows
            ml += e;
            q++;
        }
        const uint32_t o = lit + ml + 4;
        if (cum + o > kBatchOut) break;
        if (lane == k) myr = r;
        cum += o;
        r = q;
        n = k + 1;
    }
    if (n < 3) return 0;   // not worth a batch: the caller's one-sequence paths take it
    // ---- 2. my sequence
    uint32_t lit = 0, mlen = 0, d = 0;
    uint32_t la = wa + myr + 1;   // shared address of my literals
    if (lane < n) {
        const uint32_t token = lds_u8(wa + myr);
        lit = token >> 4;
        uint32_t ml = token & 15;
        if (lit == 15) {
            lit += lds_u8(la);
            la++;
        }
        d = lds_u8(la + lit) | (lds_u8(la + lit + 1) << 8);
        if (ml == 15) ml += lds_u8(la + lit + 2);
        mlen = ml + 4;
    }
    batch_commit(out, n, cum, lit, la, mlen, d);
    sp += r;
    return n;
}

// LZ4.cs:176-200.  The block occupies relative input bytes [sp, end).
__device__ int lz4_block(InStream& in, GOut& out, uint32_t sp, uint32_t end) {
    // sequences to decode one at a time after a batch attempt that found too few regular ones (doubling back-off:
    // long literal runs and long matches never batch, and a failed attempt costs a chain walk)
    uint32_t hold = 0, backoff = 4;
    while (sp < end) {
#ifndef AURORA_NO_LZ4_BATCH
        if (hold == 0) {
            if (lz4_batch32(in, out, sp, end)) {
                backoff = 4;
                continue;
            }
            hold = backoff;
            backoff = min(backoff * 2, 64u);
        }
        hold--;
#endif
        in.ensure(sp, 64);
        const uint32_t token = in.at(sp);
        if ((token >> 4) != 15 && sp + 4 + (token >> 4) <= end) {
            // fast path: no literal extension, at most one match-length extension byte, literals + match fit one warp step
            // (the vast majority of sequences; a 32-byte tile copy is token 0x0F + one extension byte)
            const uint32_t lit = token >> 4, off = sp + 1 + lit;
            uint32_t ml = token & 15, nx = off + 2;
            if (ml == 15) {
                ml += in.at(nx);   // 255 (more extension bytes follow) fails the size test below
                nx++;
            }
            if (lit + ml + 4 <= 32) {
                const uint32_t d = in.at(off) | (in.at(off + 1) << 8);
                // (measured and rejected: an L1 prefetch of the NEXT sequence's far source here — 8 % slower, the kernel is issue bound)
                out.seq_copy_small(in, sp + 1, lit, d, ml + 4);
                sp = nx;
                continue;
            }
        }
        sp++;
        uint32_t plain = token >> 4;
        if (!lz4_ext(in, sp, end, plain)) return AURORA_END_OF_STREAM;
        if (uint64_t(sp) + plain > end) return AURORA_END_OF_STREAM;
        out.lit_copy(in, sp, plain);
        sp += plain;
        if (sp >= end) break;
        if (sp + 2 > end) return AURORA_END_OF_STREAM;
        in.ensure(sp, 16);
        const uint32_t d = in.at(sp) | (in.at(sp + 1) << 8);
        sp += 2;
        uint32_t ml = token & 15;
        if (!lz4_ext(in, sp, end, ml)) return AURORA_END_OF_STREAM;
        out.match_copy(d, ml + 4);
    }
    return AURORA_OK;
}

__device__ __forceinline__ uint32_t rd32(InStream& in, uint32_t p) {
    in.ensure(p, 8);
    return in.at(p) | (in.at(p + 1) << 8) | (in.at(p + 2) << 16) | (in.at(p + 3) << 24);
}
__device__ __forceinline__ bool lz4_magic(uint32_t v) {
    return v == 0x184C2102u || v == 0x184D2204u || (v >= 0x184D2A50u && v <= 0x184D2A5Fu);
}

// XXH32 (seed 0) of n bytes at p in global memory, the hash the reference injects as LZ4.HashAlgorithm
// (LZ4.Frame.cs:24-40).  The four stripe accumulators are independent: lanes 0..3 own one each; lane 0 folds them and
// walks the tail.  The result is broadcast to the warp.
__device__ uint32_t xxh32_warp(const uint8_t* p, uint32_t n) {
    const uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
    const uint32_t lane = lane_id();
    auto rd32 = [&](uint32_t o) { return uint32_t(p[o]) | (uint32_t(p[o + 1]) << 8) | (uint32_t(p[o + 2]) << 16) | (uint32_t(p[o + 3]) << 24); };
    auto rotl = [](uint32_t x, int r) { return (x << r) | (x >> (32 - r)); };
    uint32_t h;
    const uint32_t stripes = n / 16;
    if (stripes > 0) {
        uint32_t v = lane == 0 ? P1 + P2 : lane == 1 ? P2 : lane == 2 ? 0u : 0u - P1;
        if (lane < 4)
            for (uint32_t s = 0; s < stripes; s++) v = rotl(v + rd32(16 * s + 4 * lane) * P2, 13) * P1;
        const uint32_t v1 = __shfl_sync(kFull, v, 0), v2 = __shfl_sync(kFull, v, 1), v3 = __shfl_sync(kFull, v, 2), v4 = __shfl_sync(kFull, v, 3);
        h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
    } else {
        h = P5;
    }
    h += n;
    uint32_t o = stripes * 16;
    while (o + 4 <= n) {
        h = rotl(h + rd32(o) * P3, 17) * P4;
        o += 4;
    }
    while (o < n) {
        h = rotl(h + uint32_t(p[o]) * P5, 11) * P1;
        o++;
    }
    h ^= h >> 15;
    h *= P2;
    h ^= h >> 13;
    h *= P3;
    h ^= h >> 16;
    return h;
}

// LZ4.cs:50-111 + LZ4.Frame.cs:107-174.  `src` is the blob in global memory (checksum verification reads it directly).
__device__ Res lz4_container(InStream& in, GOut& out, const uint8_t* src, uint32_t slen, int verify) {
    uint32_t sp = 0;
    while (sp < slen) {
        if (sp + 4 > slen) return Res{AURORA_END_OF_STREAM, slen};
        uint32_t magic = rd32(in, sp);
        sp += 4;
        for (;;) {   // SwitchStart
            if (magic == 0x184C2102u) {   // legacy: ReadLZ4L
                if (sp + 4 > slen) return Res{AURORA_END_OF_STREAM, slen};
                uint32_t bs = rd32(in, sp);
                sp += 4;
                uint32_t next_magic = 0;
                for (;;) {
                    if (uint64_t(sp) + bs > slen) return Res{AURORA_END_OF_STREAM, slen};
                    out.new_window();   // a fresh LzWindows per legacy block (LZ4.cs:164)
                    const int st = lz4_block(in, out, sp, sp + bs);
                    if (st != AURORA_OK) return Res{st, sp + bs};
                    sp += bs;
                    if (sp >= slen) return Res{AURORA_OK, sp};               // ReadByte() == -1
                    in.ensure(sp, 8);
                    if (in.at(sp) == 0xFF) return Res{AURORA_OK, sp + 1};   // (sbyte)0xFF == -1: the encoder's end flag
                    if (sp + 4 > slen) return Res{AURORA_END_OF_STREAM, slen};
                    bs = rd32(in, sp);
                    sp += 4;
                    if (lz4_magic(bs)) { next_magic = bs; break; }
                }
                magic = next_magic;
                continue;
            } else if (magic == 0x184D2204u) {   // v1 frame
                const uint32_t dest_start = out.written;
                if (sp + 2 > slen) return Res{AURORA_END_OF_STREAM, slen};
                in.ensure(sp, 32);
                const uint32_t FLG = in.at(sp), BD = in.at(sp + 1);
                sp += 2;
                uint32_t bmax;
                switch ((BD & 0x70) >> 4) {
                    case 4: bmax = 0x10000; break;
                    case 5: bmax = 0x40000; break;
                    case 6: bmax = 0x100000; break;
                    case 7: bmax = 0x400000; break;
                    default: return Res{AURORA_INVALID_DATA, sp};
                }
                uint64_t content = 0;
                if (FLG & 8) {
                    if (sp + 8 > slen) return Res{AURORA_END_OF_STREAM, slen};
                    content = uint64_t(rd32(in, sp)) | (uint64_t(rd32(in, sp + 4)) << 32);
                    sp += 8;
                }
                if (FLG & 1) {
                    if (sp + 4 > slen) return Res{AURORA_END_OF_STREAM, slen};
                    sp += 4;
                }
                if (sp + 1 > slen) return Res{AURORA_END_OF_STREAM, slen};
                sp += 1;   // header checksum byte: read, never verified (LZ4.FrameDescriptor.cs:25)
                if (FLG & 1) return Res{AURORA_NOT_SUPPORTED, sp};
                out.new_window();   // one window for all blocks of the frame (linked blocks)
                for (;;) {
                    if (sp + 4 > slen) return Res{AURORA_END_OF_STREAM, slen};
                    const uint32_t bsz = rd32(in, sp);
                    sp += 4;
                    if (bsz == 0) break;
                    const bool stored = (bsz & 0x80000000u) != 0;
                    const uint32_t sz = bsz & 0x7FFFFFFFu;
                    if (sz > bmax) return Res{AURORA_INVALID_DATA, sp};
                    if (uint64_t(sp) + sz > slen) return Res{AURORA_END_OF_STREAM, slen};
                    const uint32_t blk = sp;
                    sp += sz;
                    if (FLG & 16) {   // block checksum (CheckChecksum, LZ4.Frame.cs:26-41)
                        if (sp + 4 > slen) return Res{AURORA_END_OF_STREAM, slen};
                        // read from global memory: the staged input only moves forward and the block comes first
                        const uint32_t want = uint32_t(src[sp]) | (uint32_t(src[sp + 1]) << 8) | (uint32_t(src[sp + 2]) << 16) | (uint32_t(src[sp + 3]) << 24);
                        sp += 4;
                        if (verify && !out.size_only && xxh32_warp(src + blk, sz) != want) return Res{AURORA_INVALID_DATA, sp};
                    }
                    if (stored) {
                        out.lit_copy(in, blk, sz);
                    } else {
                        const int st = lz4_block(in, out, blk, blk + sz);
                        if (st != AURORA_OK) return Res{st, sp};
                    }
                }
                if ((FLG & 8) && uint64_t(out.written) != uint64_t(dest_start) + content) return Res{AURORA_SIZE_MISMATCH, sp};
                if (FLG & 4) {   // content checksum over the decoded frame
                    if (sp + 4 > slen) return Res{AURORA_END_OF_STREAM, slen};
                    const uint32_t want = uint32_t(src[sp]) | (uint32_t(src[sp + 1]) << 8) | (uint32_t(src[sp + 2]) << 16) | (uint32_t(src[sp + 3]) << 24);
                    sp += 4;
                    if (verify && !out.size_only) {
                        out.sync_tail();
                        if (uint64_t(out.written) > out.cap) return Res{AURORA_DST_TOO_SMALL, sp};
                        if (xxh32_warp(out.dst + dest_start, out.written - dest_start) != want) return Res{AURORA_INVALID_DATA, sp};
                    }
                }
                break;
            } else if (magic >= 0x184D2A50u && magic <= 0x184D2A5Fu) {   // skippable
                if (sp + 4 > slen) return Res{AURORA_END_OF_STREAM, slen};
                const uint32_t bs = rd32(in, sp);
                sp += 4;
                const uint64_t np = uint64_t(sp) + bs;
                sp = np > slen ? slen : uint32_t(np);   // Position may run past the end; the loop then stops
                break;
            } else {
                return Res{AURORA_OK, sp - 4};
            }
        }
    }
    return Res{AURORA_OK, sp};
}

// ------------------------------------------------------------------------------------------------ Snappy
// Element-per-lane batch of Snappy elements (Snappy.cs:216-247): literals of 1..60 bytes and the 2- and 3-byte copies;
// literals with length bytes and 5-byte copies end the batch and go through the element-at-a-time path.  `room` = bytes
// the block may still decode: an element that would reach or pass it is left to that path, too (it ends the loop).
__device__ __forceinline__ uint32_t snappy_batch32(InStream& in, GOut& out, uint32_t& sp, const uint32_t slen, const uint64_t room) {
    const uint32_t lane = lane_id();
    in.ensure(sp, kInMirror - 16);
    const uint32_t wa = smem_u32(in.window(sp));
    constexpr uint32_t kWin = kInMirror - 16;
    const uint32_t avail = min(slen - sp, kWin);
    const uint32_t cap_out = uint32_t(min(uint64_t(kBatchOut), room));
    uint32_t r = 0, cum = 0, n = 0, myr = 0;
#pragma unroll 1
    for (uint32_t k = 0; k < 32; k++) {
        if (r >= avail) break;
        const uint32_t tag = lds_u8(wa + r);
        const uint32_t type = tag & 3, L = tag >> 2;
        uint32_t size, o;
        if (type == 0) {
            if (L >= 60) break;
            o = L + 1;
            size = 1 + o;
        } else if (type == 1) {
            o = (L & 7) + 4;
            size = 2;
        } else if (type == 2) {
            o = L + 1;
            size = 3;
        } else {
            break;
        }
        if (r + size > avail || cum + o >= cap_out) break;
        if (lane == k) myr = r;
        cum += o;
        r += size;
        n = k + 1;
    }
    if (n < 3) return 0;
    uint32_t lit = 0, mlen = 0, d = 0;
    const uint32_t a = wa + myr;
    if (lane < n) {
        const uint32_t tag = lds_u8(a);
        const uint32_t type = tag & 3, L = tag >> 2;
        if (type == 0) {
            lit = L + 1;
        } else if (type == 1) {
            mlen = (L & 7) + 4;
            d = ((tag >> 5) << 8) | lds_u8(a + 1);
        } else {
            mlen = L + 1;
            d = lds_u8(a + 1) | (lds_u8(a + 2) << 8);
        }
    }
    batch_commit(out, n, cum, lit, a + 1, mlen, d);
    sp += r;
    return n;
}

// Snappy.cs:205-250 on relative input starting at sp; returns the position after the block's last token
__device__ Res snappy_block(InStream& in, GOut& out, uint32_t sp, uint32_t slen) {
    // ReadDecompressedSize (:109-122)
    uint32_t size = 0, shift = 0, b;
    do {
        if (sp >= slen) return Res{AURORA_END_OF_STREAM, slen};
        in.ensure(sp, 8);
        b = in.at(sp++);
        if (shift < 32) size |= (b & 0x7F) << shift;
        shift += 7;
    } while (b & 0x80);
    const uint64_t end_position = uint64_t(out.written) + size;
    if (!out.size_only && end_position > out.cap) return Res{AURORA_DST_TOO_SMALL, sp};   // SetLength on a fixed destination
    out.new_window();
    uint32_t hold = 0, backoff = 4;   // elements to decode one at a time after a failed batch attempt (doubling back-off)
    while (out.written < end_position) {
        if (sp >= slen) return Res{AURORA_END_OF_STREAM, slen};
#ifndef AURORA_NO_SNAPPY_BATCH
        if (hold == 0) {
            if (snappy_batch32(in, out, sp, slen, end_position - out.written)) {
                backoff = 4;
                continue;
            }
            hold = backoff;
            backoff = min(backoff * 2, 64u);
        }
        hold--;
#endif
        in.ensure(sp, 16);
        const uint32_t tag = in.at(sp++);
        const uint32_t type = tag & 3;
        uint32_t length = tag >> 2;
        uint32_t distance;
        if (type == 0) {
            if (length >= 60) {
                const uint32_t nb = length - 59;
                if (sp + nb > slen) return Res{AURORA_END_OF_STREAM, slen};
                length = 0;
                for (uint32_t i = 0; i < nb; i++) length |= in.at(sp + i) << (8 * i);
                sp += nb;
            }
            const int32_t l1 = int32_t(length) + 1;
            if (l1 < 0) return Res{AURORA_INVALID_DATA, sp};
            if (!out.copy_from(in, sp, uint32_t(l1), slen - sp)) return Res{AURORA_END_OF_STREAM, slen};
            sp += uint32_t(l1);
            continue;
        } else if (type == 1) {
            if (sp + 1 > slen) return Res{AURORA_END_OF_STREAM, slen};
            length = (length & 7) + 3;
            distance = ((tag >> 5) << 8) | in.at(sp);
            sp += 1;
        } else if (type == 2) {
            if (sp + 2 > slen) return Res{AURORA_END_OF_STREAM, slen};
            distance = in.at(sp) | (in.at(sp + 1) << 8);
            sp += 2;
        } else {
            if (sp + 4 > slen) return Res{AURORA_END_OF_STREAM, slen};
            distance = in.at(sp) | (in.at(sp + 1) << 8) | (in.at(sp + 2) << 16) | (in.at(sp + 3) << 24);
            sp += 4;
            if (distance > 0x10000u) return Res{AURORA_INVALID_DATA, sp};   // aliases in the reference's 64 KiB ring
        }
        out.match_copy(distance, length + 1);
    }
    return Res{AURORA_OK, sp};
}

// Snappy.cs:39-68
__device__ Res snappy_framed(InStream& in, GOut& out, uint32_t slen) {
    if (slen < 10) return Res{AURORA_END_OF_STREAM, slen};
    in.ensure(0, 16);
    const uint32_t id[10] = {0xff, 0x06, 0x00, 0x00, 0x73, 0x4e, 0x61, 0x50, 0x70, 0x59};
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 10; i++) ok = ok && in.at(i) == id[i];
    if (!ok) return Res{AURORA_INVALID_IDENTIFIER, 10};
    uint32_t sp = 10;
    while (sp < slen) {
        in.ensure(sp, 16);
        const uint32_t type = in.at(sp);
        if (sp + 4 > slen) return Res{AURORA_END_OF_STREAM, slen};
        const uint32_t clen = in.at(sp + 1) | (in.at(sp + 2) << 8) | (in.at(sp + 3) << 16);
        sp += 4;
        if (type == 0) {
            if (sp + 4 > slen) return Res{AURORA_END_OF_STREAM, slen};
            sp += 4;   // CRC skipped (:50)
            const Res r = snappy_block(in, out, sp, slen);
            if (r.status != AURORA_OK) return r;
            sp = r.consumed;
        } else if (type == 1) {
            if (sp + 4 > slen) return Res{AURORA_END_OF_STREAM, slen};
            sp += 4;
            if (clen < 4) return Res{AURORA_INVALID_DATA, sp};
            const uint32_t n = min(clen - 4, slen - sp);   // SubStream.CopyTo copies what is there
            out.lit_copy(in, sp, n);
            sp += n;
        } else {
            if (type >= 0x02 && type <= 0x7F) return Res{AURORA_INVALID_DATA, sp};
            const uint64_t np = uint64_t(sp) + clen;
            sp = np > slen ? slen : uint32_t(np);
        }
    }
    return Res{AURORA_OK, sp};
}

// ------------------------------------------------------------------------------------------------ LZO
// LZO.cs:252-262
__device__ __forceinline__ bool lzo_ext(InStream& in, uint32_t& sp, uint32_t slen, uint32_t& length) {
    uint32_t b;
    for (;;) {
        if (sp >= slen) return false;
        in.ensure(sp, 8);
        b = in.at(sp++);
        if (b != 0) break;
        length += 255;
    }
    length += b;
    return true;
}

// Element-per-lane batch of LZO instructions (LZO.cs:62-137).  The interpretation of an instruction depends on the
// trailing-literal count of the one before (`plain`), which the uniform chain carries along: per instruction it only
// derives the size, the output bytes and the next `plain`; lane k then decodes instruction k completely.  An element is a
// literal run, or a match followed by its 0..3 trailing literals.  Length extensions of more than one byte, the end
// marker and instructions that leave the staged window end the batch.  `at` = position of the first flag byte.
// Measured (round 1), batch + 20 warps vs element-at-a-time + 23 warps: C2 corpus 284 vs 266 GB/s, but C4-like (131 072 streams
// of 4-64 KiB, the BASELINE config of LZO) 301 vs 351 GB/s — LZO's chain is twice as heavy as LZ4's and short literal-heavy
// streams rarely batch.  Re-measured at 16 warps (64 registers, no spills; the 20 / 23-warp builds spill): C2 corpus 313 vs 266
// (class T 371 vs 247), C4-like 328 vs 351 GB/s: the batches win on long streams, the resident warps win on the short streams of
// LZO's BASELINE config.  Compiled only with -DAURORA_LZO_BATCH (-DAURORA_LZO_WARPS=16) until one build wins both.
__device__ __forceinline__ uint32_t lzo_batch32(InStream& in, GOut& out, uint32_t& at, uint32_t& plain_io, const uint32_t slen) {
    const uint32_t lane = lane_id();
    in.ensure(at, kInMirror - 16);
    const uint32_t wa = smem_u32(in.window(at));
    constexpr uint32_t kWin = kInMirror - 16;
    const uint32_t avail = min(slen - at, kWin);
    uint32_t r = 0, cum = 0, n = 0, myr = 0, myplain = 0, plain = plain_io;
#pragma unroll 1
    for (uint32_t k = 0; k < 32; k++) {
        if (r + 4 > avail) break;   // flag + up to three header bytes must be readable (the tail of the stream goes one by one)
        const uint32_t flag = lds_u8(wa + r);
        const uint32_t code = flag >> 4;
        uint32_t size, o, np;
        if (code == 0 && plain == 0) {
            uint32_t len = 3 + flag, hdr = 1;
            if (len == 3) {
                const uint32_t e = lds_u8(wa + r + 1);
                if (e == 0) break;          // more extension bytes
                len = 18 + e;
                hdr = 2;
            }
            size = hdr + len;
            o = len;
            np = 4;
        } else {
            uint32_t ml, hdr, pb = flag;    // pb: the byte whose low two bits give the trailing literals
            if (code == 0) {
                ml = plain <= 3 ? 2 : 3;
                hdr = 2;
            } else if (code <= 3) {
                const uint32_t mask = code == 1 ? 7u : 31u;
                ml = 2 + (flag & mask);
                hdr = 3;
                if (ml == 2) {
                    const uint32_t e = lds_u8(wa + r + 1);
                    if (e == 0) break;
                    ml = (code == 1 ? 9u : 33u) + e;
                    hdr = 4;
                }
                pb = lds_u8(wa + r + hdr - 2);
                if (code == 1) {
                    const uint32_t dist = (16384u + ((flag & 8) << 11)) | (lds_u8(wa + r + hdr - 1) << 6) | (pb >> 2);
                    if (dist == 16384u) break;   // end marker
                }
            } else if (code <= 7) {
                ml = 3 + ((flag >> 5) & 1);
                hdr = 2;
            } else {
                ml = 5 + ((flag >> 5) & 3);
                hdr = 2;
            }
            np = pb & 3;
            size = hdr + np;
            o = ml + np;
        }
        if (r + size > avail || cum + o > kBatchOut) break;
        if (lane == k) {
            myr = r;
            myplain = plain;
        }
        cum += o;
        r += size;
        plain = np;
        n = k + 1;
    }
    if (n < 3) return 0;
    // ---- lane k decodes instruction k
    uint32_t lit = 0, mlen = 0, d = 1, la = 0;
    if (lane < n) {
        const uint32_t a = wa + myr;
        const uint32_t flag = lds_u8(a);
        const uint32_t code = flag >> 4;
        if (code == 0 && myplain == 0) {
            lit = 3 + flag;
            la = a + 1;
            if (lit == 3) {
                lit = 18 + lds_u8(a + 1);
                la = a + 2;
            }
        } else {
            uint32_t hdr, pb = flag;
            if (code == 0) {
                mlen = myplain <= 3 ? 2 : 3;
                d = (lds_u8(a + 1) << 2) + (flag >> 2) + (myplain <= 3 ? 1u : 2049u);
                hdr = 2;
            } else if (code <= 3) {
                const uint32_t mask = code == 1 ? 7u : 31u;
                mlen = 2 + (flag & mask);
                hdr = 3;
                if (mlen == 2) {
                    mlen = (code == 1 ? 9u : 33u) + lds_u8(a + 1);
                    hdr = 4;
                }
                pb = lds_u8(a + hdr - 2);
                const uint32_t b1 = lds_u8(a + hdr - 1);
                if (code == 1) d = (16384u + ((flag & 8) << 11)) | (b1 << 6) | (pb >> 2);
                else d = ((b1 << 6) | (pb >> 2)) + 1;
            } else if (code <= 7) {
                mlen = 3 + ((flag >> 5) & 1);
                d = (lds_u8(a + 1) << 3) + ((flag >> 2) & 7) + 1;
                hdr = 2;
            } else {
                mlen = 5 + ((flag >> 5) & 3);
                d = (lds_u8(a + 1) << 3) + ((flag & 0x1c) >> 2) + 1;
                hdr = 2;
            }
            lit = pb & 3;
            la = a + hdr;
        }
    }
    batch_commit<true>(out, n, cum, lit, la, mlen, d);
    at += r;
    plain_io = plain;
    return n;
}

// LZO.cs:49-139
__device__ Res lzo_decode(InStream& in, GOut& out, uint32_t slen) {
    uint32_t sp = 0, plain = 0, length, distance;
    const Res eos{AURORA_END_OF_STREAM, slen};
    if (sp >= slen) return eos;
    in.ensure(sp, 16);
    uint32_t flag = in.at(sp++);
    if (flag > 17) {
        length = flag - 17;
        if (!out.copy_from(in, sp, length, slen - sp)) return eos;
        sp += length;
        if (sp >= slen) return eos;
        in.ensure(sp, 16);
        flag = in.at(sp++);
    }
    uint32_t hold = 0, backoff = 4;   // instructions to decode one at a time after a failed batch attempt (doubling back-off)
    for (;;) {
#ifdef AURORA_LZO_BATCH   // off by default: see lzo_batch32
        if (hold == 0) {
            uint32_t at = sp - 1;   // the flag byte just read
            if (lzo_batch32(in, out, at, plain, slen)) {
                backoff = 4;
                sp = at;
                if (sp >= slen) return eos;   // while ((flag = ReadByte()) != -1) ... throw EndOfStream
                in.ensure(sp, 16);
                flag = in.at(sp++);
                continue;
            }
            hold = backoff;
            backoff = min(backoff * 2, 64u);
        }
        hold--;
#endif
        in.ensure(sp, 16);
        const uint32_t code = flag >> 4;
        bool literal_run = false;
        if (code == 0) {
            if (plain == 0) {
                length = 3 + flag;
                if (length == 3) {
                    length = 18;
                    if (!lzo_ext(in, sp, slen, length)) return eos;
                }
                plain = 4;
                if (!out.copy_from(in, sp, length, slen - sp)) return eos;
                sp += length;
                literal_run = true;
            } else if (plain <= 3) {
                if (sp >= slen) return eos;
                distance = (in.at(sp++) << 2) + (flag >> 2) + 1;
                length = 2;
            } else {
                if (sp >= slen) return eos;
                distance = (in.at(sp++) << 2) + (flag >> 2) + (2048 + 1);
                length = 3;
            }
        } else if (code == 1) {
            length = 2 + (flag & 7);
            if (length == 2) {
                length = 9;
                if (!lzo_ext(in, sp, slen, length)) return eos;
            }
            distance = 16384 + ((flag & 8) << 11);
            if (sp + 2 > slen) return eos;
            in.ensure(sp, 8);
            flag = in.at(sp++);
            distance |= (in.at(sp++) << 6) | (flag >> 2);
            if (distance == 16384) return Res{AURORA_OK, sp};
        } else if (code <= 3) {
            length = 2 + (flag & 0x1f);
            if (length == 2) {
                length = 33;
                if (!lzo_ext(in, sp, slen, length)) return eos;
            }
            if (sp + 2 > slen) return eos;
            in.ensure(sp, 8);
            flag = in.at(sp++);
            distance = ((in.at(sp++) << 6) | (flag >> 2)) + 1;
        } else if (code <= 7) {
            length = 3 + ((flag >> 5) & 1);
            if (sp >= slen) return eos;
            distance = (in.at(sp++) << 3) + ((flag >> 2) & 7) + 1;
        } else {
            length = 5 + ((flag >> 5) & 3);
            if (sp >= slen) return eos;
            distance = (in.at(sp++) << 3) + ((flag & 0x1c) >> 2) + 1;
        }
        if (!literal_run) {
            plain = flag & 3;
            out.match_copy(distance, length);
            if (!out.copy_from(in, sp, plain, slen - sp)) return eos;
            sp += plain;
        }
        if (sp >= slen) return eos;   // while ((flag = ReadByte()) != -1) ... throw EndOfStream
        in.ensure(sp, 16);
        flag = in.at(sp++);
    }
}

// ------------------------------------------------------------------------------------------------ PRS
struct PrsBits {
    uint32_t cur, left;   // the unread bits of the current flag byte in READ order, next bit = bit 0 (an MSB-first byte is reversed on fetch)
    bool msb;
    // FlagReader.Readbit (IO/FlagReader.cs:53-65): the flag byte is fetched lazily from the same stream
    __device__ __forceinline__ int bit(InStream& in, uint32_t& sp, uint32_t slen) {
        if (left == 0) {
            if (sp >= slen) return -1;
            in.ensure(sp, 8);
            const uint32_t b = in.at(sp++);
            cur = msb ? __brev(b) >> 24 : b;
            left = 8;
        }
        const int v = int(cur & 1u);
        cur >>= 1;
        left--;
        return v;
    }
};

// Element-per-lane batch of PRS tokens (PRS.cs:59-102).  An element = the literal bits that follow each other inside ONE
// flag byte (their data bytes are contiguous in the input) plus the match token behind them, if the flag byte has a bit
// left.  The chain is the bit walk itself (flag bytes are fetched lazily between the data bytes, so every position
// depends on every bit before it) and is uniform work; lane k keeps element k as the chain passes it, and the shared
// batch commit places and copies all elements.  The end token, a token that needs input bytes past the staged window /
// the end of the input, and an element that would pass kBatchOut decoded bytes end the batch BEFORE that element (the
// state is rolled back to the element's start); the token-at-a-time path in prs_walk decodes it with the reference's
// exact error behaviour.  Returns the number of elements consumed.
__device__ __forceinline__ uint32_t prs_batch32(InStream& in, GOut& out, PrsBits& fr, uint32_t& sp, const uint32_t slen) {
    const uint32_t lane = lane_id();
    in.ensure(sp, kInMirror - 16);
    const uint32_t wa = smem_u32(in.window(sp));
    constexpr uint32_t kWin = kInMirror - 16;
    const uint32_t avail = min(slen - sp, kWin);
    const bool big = fr.msb;
    uint32_t bits = fr.cur, left = fr.left, r = 0;          // running state
    uint32_t cbits = bits, cleft = left, cr = 0;            // state after the last committed element
    uint32_t cum = 0, n = 0;
    uint32_t my_la = wa, my_lit = 0, my_len = 0, my_d = 0;
#define PRS_FETCH()                                   \
    {                                                 \
        if (r >= avail) break;                        \
        const uint32_t fb = lds_u8(wa + r);           \
        r++;                                          \
        bits = big ? __brev(fb) >> 24 : fb;           \
        left = 8;                                     \
    }
#pragma unroll 1
    for (uint32_t k = 0; k < 32; k++) {
        if (left == 0) PRS_FETCH();
        const uint32_t ones = uint32_t(__ffs(int(~bits))) - 1u;   // literal bits; bits < 2^left, so ones <= left
        const uint32_t e_la = wa + r;
        r += ones;
        bits >>= ones;
        left -= ones;
        if (r > avail) break;
        uint32_t mlen = 0, d = 0;
        if (left != 0) {
            bits >>= 1;   // the 0 bit that opens a match token
            left--;
            if (left == 0) PRS_FETCH();
            const uint32_t b = bits & 1u;
            bits >>= 1;
            left--;
            if (b) {
                if (r + 3 > avail) break;   // u16 + a possible length byte
                const uint32_t x0 = lds_u8(wa + r), x1 = lds_u8(wa + r + 1);
                const uint32_t v = big ? (x0 << 8) | x1 : x0 | (x1 << 8);
                r += 2;
                if (v == 0) break;          // end token
                mlen = v & 7u;
                d = 0x2000u - (v >> 3);
                if (mlen == 0) {
                    mlen = lds_u8(wa + r) + 1u;
                    r++;
                } else {
                    mlen += 2;
                }
            } else {
                if (left == 0) PRS_FETCH();
                const uint32_t b1 = bits & 1u;
                bits >>= 1;
                left--;
                if (left == 0) PRS_FETCH();
                const uint32_t b0 = bits & 1u;
                bits >>= 1;
                left--;
                if (r >= avail) break;
                mlen = b1 * 2 + b0 + 2;
                d = 0x100u - lds_u8(wa + r);
                r++;
            }
        }
        const uint32_t o = ones + mlen;
        if (cum + o > kBatchOut) break;
        if (lane == k) {
            my_la = e_la;
            my_lit = ones;
            my_len = mlen;
            my_d = d;
        }
        cum += o;
        n = k + 1;
        cbits = bits;
        cleft = left;
        cr = r;
    }
#undef PRS_FETCH
    if (n < 3) return 0;
    batch_commit(out, n, cum, my_lit, my_la, my_len, my_d);
    fr.cur = cbits;
    fr.left = cleft;
    sp += cr;
    return n;
}

// PRS.cs:59-102 (copy == true) and ValidateByteOrder :172-218 (copy == false; status kPrsValid / kPrsInvalid)
constexpr int kPrsValid = 100, kPrsInvalid = 101;
template <bool kCopy>
__device__ Res prs_walk(InStream& in, GOut& out, uint32_t slen, bool big) {
    PrsBits fr{0, 0, big};
    uint32_t sp = 0;
    uint32_t produced = 0;
    int budget = 3;
    const Res eos{AURORA_END_OF_STREAM, slen};
    uint32_t hold = 0, backoff = 4;   // tokens to decode one at a time after a failed batch attempt (doubling back-off)
    while (sp < slen) {
#ifndef AURORA_NO_PRS_BATCH
        if (kCopy) {
            if (hold == 0) {
                if (prs_batch32(in, out, fr, sp, slen)) {
                    backoff = 4;
                    continue;
                }
                hold = backoff;
                backoff = min(backoff * 2, 64u);
            }
            hold--;
        }
#endif
        int b = fr.bit(in, sp, slen);
        if (b < 0) return eos;
        if (b) {
            if (sp >= slen) return kCopy ? eos : Res{kPrsInvalid, sp};   // validate: Position++ runs past the end, loop ends, false
            if (kCopy) {
                // the literal bits that follow in the SAME flag byte govern the next input bytes: copy the run at once
                uint32_t n = 1;
                while (fr.left > 0 && (fr.cur & 1u)) {
                    n++;
                    fr.cur >>= 1;
                    fr.left--;
                }
                const uint32_t avail = slen - sp;
                out.lit_copy(in, sp, min(n, avail));
                if (n > avail) return eos;   // ReadUInt8 at the end of the input
                sp += n;
                continue;
            }
            sp++;
            produced++;
        } else {
            uint32_t distance, length;
            b = fr.bit(in, sp, slen);
            if (b < 0) return eos;
            if (b) {
                if (sp + 2 > slen) return eos;
                in.ensure(sp, 8);
                const uint32_t v = big ? (in.at(sp) << 8) | in.at(sp + 1) : in.at(sp) | (in.at(sp + 1) << 8);
                sp += 2;
                if (v == 0) return kCopy ? Res{AURORA_OK, sp} : Res{kPrsValid, sp};
                length = v & 7;
                distance = 0x2000 - (v >> 3);
                if (length == 0) {
                    if (sp >= slen) return eos;
                    length = in.at(sp++) + 1;
                } else {
                    length += 2;
                }
            } else {
                const int b1 = fr.bit(in, sp, slen);
                if (b1 < 0) return eos;
                const int b0 = fr.bit(in, sp, slen);
                if (b0 < 0) return eos;
                length = uint32_t(b1 * 2 + b0) + 2;
                if (sp >= slen) return eos;
                in.ensure(sp, 8);
                distance = 0x100 - in.at(sp++);
            }
            if (kCopy) {
                out.match_copy(distance, length);
            } else {
                if (distance > produced) return Res{kPrsInvalid, sp};
                if (budget == 0) return Res{kPrsValid, sp};
                budget--;
                produced += length;
            }
        }
    }
    return kCopy ? eos : Res{kPrsInvalid, sp};
}

// PRS.cs:42-57 + GetByteOrder :161-170
__device__ Res prs_decode(InStream& in, GOut& out, uint32_t slen, const uint8_t* src, const DecodeParams& P) {
    if (slen == 0) return Res{AURORA_END_OF_STREAM, 0};
    in.ensure(0, 8);
    const uint32_t flag = in.at(0);
    int detected = -1;   // 0 little, 1 big
    if (flag > 12 && (flag & 1)) {
        const Res v = prs_walk<false>(in, out, slen, false);
        if (v.status == AURORA_END_OF_STREAM) return Res{AURORA_END_OF_STREAM, 0};   // the exception leaves GetByteOrder; `finally` restored Position
        if (v.status == kPrsValid) detected = 0;
        in.begin(P.src_base, P.src_limit, src);
    }
    if (detected < 0 && (flag & 128)) {
        const Res v = prs_walk<false>(in, out, slen, true);
        if (v.status == AURORA_END_OF_STREAM) return Res{AURORA_END_OF_STREAM, 0};
        if (v.status == kPrsValid) detected = 1;
        in.begin(P.src_base, P.src_limit, src);
    }
    const bool first_big = detected == 1;
    Res r{AURORA_OK, 0};
#pragma unroll 1   // one inlined copy of the walk: the retry with the other order (PRS.cs:49-56) is the second trip
    for (int attempt = 0; attempt < 2; attempt++) {
        if (attempt) {
            __syncwarp();   // the first attempt's ring reads are done before the second one rewrites the ring
            in.begin(P.src_base, P.src_limit, src);
            out.written = 0;
            out.flushed = 0;
        }
        out.new_window();
        r = prs_walk<true>(in, out, slen, attempt ? !first_big : first_big);
        if (r.status == AURORA_OK) break;
    }
    return r;
}

template <int K>
__device__ void decode_stream(const DecodeParams& P, uint32_t idx, InStream& in, uint32_t rb) {
    const uint8_t* src = P.src_base + P.src_off[idx];
    const uint64_t slen64 = P.src_len[idx];
    const uint32_t slen = slen64 > 0xFFFFFFF0ull ? 0xFFFFFFF0u : uint32_t(slen64);
    GOut out;
    out.dst = P.dst_base + P.dst_off[idx];
    out.cap = P.size_only ? 0 : P.dst_cap[idx];
    out.rb = rb;
    out.written = 0;
    out.flushed = 0;
    out.win_base = 0;
    out.aligned = (reinterpret_cast<uintptr_t>(out.dst) & 15) == 0;
    out.ring_len = (K == B_PRS) ? 0x2000u : 0x10000u;
    out.size_only = P.size_only != 0;
    in.begin(P.src_base, P.src_limit, src);
    Res r;
    if (K == B_LZ4) r = lz4_container(in, out, src, slen, P.lz4_verify);
    else if (K == B_LZ4_BLOCK) {
        const int st = lz4_block(in, out, 0, slen);
        r = Res{st, slen};
    } else if (K == B_SNAPPY) r = snappy_framed(in, out, slen);
    else if (K == B_SNAPPY_BLOCK) r = snappy_block(in, out, 0, slen);
    else if (K == B_LZO) r = lzo_decode(in, out, slen);
    else r = prs_decode(in, out, slen, src, P);
    out.finish();
    if (r.status == AURORA_OK && uint64_t(out.written) > out.cap) r.status = AURORA_DST_TOO_SMALL;
    if (lane_id() == 0) {
        P.out_len[idx] = out.written;
        P.consumed[idx] = r.consumed;
        P.status[idx] = r.status;
    }
}

#ifndef AURORA_LZ4_WARPS
#define AURORA_LZ4_WARPS 20
#endif
// resident warps per block (x2 blocks per SM): the LZ4 kernels trade warps for the registers of the sequence-per-lane batch
template <int K>
struct BlockShape {
    // block kernels with an element-per-lane batch: 20 warps (48 registers); the framed kernels inline two block decoders:
    // 16 warps (64 registers); PRS: 16 warps (64 registers: the element-per-lane batch spills below that; measured 12 / 14 / 16 / 20 / 23 warps: 88 / 94 / 97 / 74 / 57 GB/s)
#ifndef AURORA_LZO_WARPS
#define AURORA_LZO_WARPS 23
#endif
#ifndef AURORA_PRS_WARPS
#define AURORA_PRS_WARPS 16
#endif
    static constexpr int kWarps = (K == B_LZ4_BLOCK || K == B_SNAPPY_BLOCK) ? AURORA_LZ4_WARPS : K == B_LZO ? AURORA_LZO_WARPS : (K == B_LZ4 || K == B_SNAPPY) ? 16 : K == B_PRS ? AURORA_PRS_WARPS : kWarpsPerBlock;
};

// ---- kernel
template <int K>
__global__ void __launch_bounds__(BlockShape<K>::kWarps * 32, 2) decode_bytelz_kernel(const DecodeParams P) {
    constexpr int kWarpsPerBlock = BlockShape<K>::kWarps;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5;
    const uint32_t s0 = smem_u32(smem);
    uint8_t* aligned = smem + (((s0 + kORing - 1) & ~uint32_t(kORing - 1)) - s0);
    const uint32_t rb = smem_u32(aligned + size_t(warp) * kORing);
    uint8_t* wbase = aligned + size_t(kWarpsPerBlock) * kORing + size_t(warp) * (kSmemPerWarp - kORing);
    InStream in;
    in.init(wbase, reinterpret_cast<uint64_t*>(wbase + kInStage));
    __syncwarp();
    fence_proxy_async();
    for (;;) {
        uint32_t t = 0;
        if (lane_id() == 0) t = atomicAdd(P.ticket, 1u);
        t = __shfl_sync(kFull, t, 0);
        if (t >= P.n) break;
        const uint32_t idx = P.order ? P.order[t] : t;
        decode_stream<K>(P, idx, in, rb);
    }
    in.drain_inflight();
}

template <int K>
cudaError_t launch(const DecodeParams& p, int sm_count, cudaStream_t st) {
    constexpr int kWarpsPerBlock = BlockShape<K>::kWarps;
    const int threads = kWarpsPerBlock * 32;
    const size_t smem = size_t(kWarpsPerBlock) * kSmemPerWarp + kORing;   // + alignment slack for the rings
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(decode_bytelz_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    int blocks = sm_count * 2;
    const int needed = int((p.n + kWarpsPerBlock - 1) / kWarpsPerBlock);
    if (needed < blocks) blocks = needed > 0 ? needed : 1;
    decode_bytelz_kernel<K><<<blocks, threads, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_decode_bytelz(const DecodeParams& p, int sm_count, cudaStream_t st) {
    switch (p.format) {
        case AURORA_FMT_LZ4:
        case AURORA_FMT_LZ4_LEGACY: return launch<B_LZ4>(p, sm_count, st);
        case AURORA_FMT_LZ4_BLOCK: return launch<B_LZ4_BLOCK>(p, sm_count, st);
        case AURORA_FMT_SNAPPY: return launch<B_SNAPPY>(p, sm_count, st);
        case AURORA_FMT_SNAPPY_BLOCK: return launch<B_SNAPPY_BLOCK>(p, sm_count, st);
        case AURORA_FMT_LZO: return launch<B_LZO>(p, sm_count, st);
        case AURORA_FMT_PRS: return launch<B_PRS>(p, sm_count, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace aurora
