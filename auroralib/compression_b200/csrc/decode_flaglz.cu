// decode_flaglz.cu — batched decoder for the flag-byte LZ77 family: LZ10, LZ11, Yaz0/Yaz1, LZSS (interleaved
// flag/token streams) and MIO0, Yay0 (three separate sub-streams).  One compressed stream per warp.
//
// Reference semantics restated on the device (paths under /root/reference/src):
//   LZ10   AuroraLib.Compression.Nintendo/Nintendo/LZ10.cs:47-57 (header) :82-111 (body)
//   LZ11   .../Nintendo/LZ11.cs:43-53, :83-133
//   Yaz0   .../Nintendo/Yaz0.cs:58-79 (header + endian retry) -> Yay0.cs:110-144 (token core)
//   Yay0   .../Nintendo/Yay0.cs:50-60, :99-144
//   MIO0   .../Nintendo/MIO0.cs:51-61, :105-149
//   LZSS   AuroraLib.Compression/Formats/Common/LZSS.cs:53-69, :91-130
//   window AuroraLib.Compression/IO/LzWindows.cs:72-115 (BackCopy/OffsetCopy), FlagReader.cs:53-65
//
// Design (B200, sm_100a):
//   * persistent grid, one warp per stream, streams handed out by a global ticket (largest first when
//     the host supplies an order array);
//   * the compressed bytes are staged into a per-warp shared-memory ring by 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx), two 1 KiB chunks in flight per sub-stream;
//   * 32 tokens are parsed per warp iteration: flag words give literal/match per lane, token offsets are
//     popcount prefix sums (LZ10/LZSS/MIO0/Yay0) or a short uniform walk over an extended-token ballot
//     mask (Yaz0/LZ11); output positions come from a shuffle prefix scan over the token lengths;
//   * the decoded bytes live in a per-warp 8 KiB shared-memory ring that always holds the last 4 KiB
//     window, so back-references never touch HBM: literals are scattered in one step, matches are
//     warp-cooperative copies out[dst+i] = out[dst-d + (i mod d)] (the BackCopy chunk loop), and the ring
//     is drained to HBM in 512-byte, 16-byte-per-lane vector stores.
#include "common.cuh"
#include "stage.cuh"

namespace aurora {

namespace {

constexpr int kRing = 8192;          // per-warp output ring (bytes)
constexpr int kRingMask = kRing - 1;
constexpr int kWindow = 4096;        // largest back-reference distance of the family
constexpr int kSubMax = 2048;        // output bytes resolved per sub-batch (ring keeps window + sub-batch + drain slack)
constexpr int kFlush = 512;          // bytes per ring->HBM drain step (16 B per lane)

enum Kind { K_LZ10 = 0, K_LZ11 = 1, K_YAZ0 = 2, K_LZSS = 3, K_MIO0 = 4, K_YAY0 = 5 };

template <int K>
struct Traits {
    static constexpr int kStreams = (K == K_MIO0 || K == K_YAY0) ? 3 : 1;
    static constexpr int kMaxTok = (K == K_LZ10 || K == K_MIO0) ? 18 : (K == K_YAZ0 || K == K_YAY0) ? 273 : (K == K_LZSS) ? 258 : 65808;
    static constexpr bool kNeedSub = kMaxTok * 32 > kSubMax;
    static constexpr int kWarps = kStreams == 1 ? 11 : 8;   // x2 blocks per SM
    static constexpr int kSmemPerWarp = kRing + kStreams * kInRing + 64;
};

// out[dst+i] = out[dst-d + (i mod d)], i < len : LzWindows.BackCopy (IO/LzWindows.cs:72-100) on the flat ring.
// All reads are below dst and all writes at or above it, so one pass has no internal hazard.
__device__ __forceinline__ void ring_copy_match(uint8_t* ring, uint32_t dstp, uint32_t d, uint32_t len) {
    const uint32_t lane = lane_id();
    const uint32_t srcp = dstp - d;
    if (d >= len) {
        for (uint32_t i = lane; i < len; i += 32) ring[(dstp + i) & kRingMask] = ring[(srcp + i) & kRingMask];
    } else {
        for (uint32_t i = lane; i < len; i += 32) ring[(dstp + i) & kRingMask] = ring[(srcp + i % d) & kRingMask];
    }
}

struct OutState {
    uint8_t* ring;
    uint8_t* dst;
    uint32_t limit;    // bytes of the destination that may be written
    uint32_t flushed;  // output position drained to HBM so far (multiple of kFlush)
    bool aligned;
    __device__ __forceinline__ void drain(uint32_t produced) {
        const uint32_t lane = lane_id();
        const uint32_t upto = min(produced, limit);
        while (flushed + kFlush <= upto) {
            if (aligned) {
                const uint4 v = *reinterpret_cast<const uint4*>(ring + ((flushed + lane * 16) & kRingMask));
                *reinterpret_cast<uint4*>(dst + flushed + lane * 16) = v;
            } else {
                for (uint32_t i = lane; i < kFlush; i += 32) dst[flushed + i] = ring[(flushed + i) & kRingMask];
            }
            flushed += kFlush;
        }
    }
    __device__ __forceinline__ void finish(uint32_t produced) {
        const uint32_t lane = lane_id();
        const uint32_t upto = min(produced, limit);
        drain(produced);
        for (uint32_t p = flushed + lane; p < upto; p += 32) dst[p] = ring[p & kRingMask];
    }
};

struct BodyResult {
    int status;
    uint32_t written;
    uint32_t consumed;
};

// ---------------------------------------------------------------------------------------------
// token core shared by all six formats.
//   interleaved formats: in[0] is the whole blob, cursor `cur` = blob offset of the next flag byte.
//   split formats: in[0] flags (relative to blob offset 0x10), in[1] codes (relative to comp_off),
//                  in[2] literals + extended lengths (relative to lit_off).
// ---------------------------------------------------------------------------------------------
template <int K>
__device__ BodyResult decode_body(InStream* in, OutState& out, const uint32_t slen, const uint32_t size,
                                  const uint32_t body_off, const uint32_t comp_off, const uint32_t lit_off,
                                  const LzssParams& lz) {
    const uint32_t lane = lane_id();
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint8_t* ring = out.ring;
    uint32_t written = 0;
    uint32_t cur = (K == K_MIO0 || K == K_YAY0) ? 0u : body_off;
    uint32_t ccur = 0, lcur = 0;   // split formats: relative cursors
    uint32_t consumed = (K == K_MIO0 || K == K_YAY0) ? max(comp_off, lit_off) : body_off;
    int status = AURORA_OK;

    while (written < size) {
        // ------------------------------------------------------------------ parse 32 tokens
        bool ism;            // match?
        uint32_t len;        // decoded bytes of this token
        uint32_t dist = 1;   // back-reference distance (LZSS: raw window offset until resolved below)
        uint32_t lit = 0;    // literal byte
        uint32_t tok_end = 0;// interleaved: blob offset just past the token
        bool bad;            // the token (or its flag byte) lies beyond the end of the input
        bool ext_used = false;   // Yay0: the token consumed an extended-length byte
        uint32_t next_cur;

        if constexpr (K == K_LZ10 || K == K_LZSS) {
            in[0].ensure(cur);
            // flag-byte chain: a group is 1 + 8 + (#matches) bytes.  m = match bits in wire bit order.
            const uint32_t p0 = cur;
            const uint32_t f0 = in[0].at(p0), m0 = (K == K_LZ10) ? f0 : (~f0 & 0xFFu);
            const uint32_t p1 = p0 + 9 + __popc(m0);
            const uint32_t f1 = in[0].at(p1), m1 = (K == K_LZ10) ? f1 : (~f1 & 0xFFu);
            const uint32_t p2 = p1 + 9 + __popc(m1);
            const uint32_t f2 = in[0].at(p2), m2 = (K == K_LZ10) ? f2 : (~f2 & 0xFFu);
            const uint32_t p3 = p2 + 9 + __popc(m2);
            const uint32_t f3 = in[0].at(p3), m3 = (K == K_LZ10) ? f3 : (~f3 & 0xFFu);
            next_cur = p3 + 9 + __popc(m3);
            const uint32_t g = lane >> 3, j = lane & 7;
            const uint32_t pg = g == 0 ? p0 : g == 1 ? p1 : g == 2 ? p2 : p3;
            const uint32_t mg = g == 0 ? m0 : g == 1 ? m1 : g == 2 ? m2 : m3;
            uint32_t before;
            if (K == K_LZ10) {   // MSB first (FlagReader bitOrder Big)
                ism = (mg >> (7 - j)) & 1;
                before = __popc(mg >> (8 - j));
            } else {             // LSB first
                ism = (mg >> j) & 1;
                before = __popc(mg & ((1u << j) - 1u));
            }
            const uint32_t tokoff = pg + 1 + j + before;
            const uint32_t b1 = in[0].at(tokoff), b2 = in[0].at(tokoff + 1);
            if (K == K_LZ10) {
                len = ism ? (b1 >> 4) + 3 : 1;
                dist = (((b1 & 0xF) << 8) | b2) + 1;
            } else {
                len = ism ? (b2 & ((1u << lz.length_bits) - 1u)) + uint32_t(lz.min_length) : 1;
                dist = ((b2 >> lz.length_bits) << 8) | b1;
            }
            lit = b1;
            tok_end = tokoff + (ism ? 2 : 1);
            bad = pg >= slen || tok_end > slen;
        } else if constexpr (K == K_YAZ0 || K == K_LZ11) {
            in[0].ensure(cur);
            uint32_t pg = cur, myoff = 0, mysz = 1, mypg = 0;
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const uint32_t f = in[0].at(pg);
                const uint32_t hi = in[0].at(pg + 1 + lane) >> 4;
                const uint32_t e0 = __ballot_sync(kFull, hi == 0);
                const uint32_t e1 = (K == K_LZ11) ? __ballot_sync(kFull, hi == 1) : 0u;
                uint32_t off = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const uint32_t bit = (f >> (7 - i)) & 1;
                    const bool is_lit = (K == K_YAZ0) ? bit != 0 : bit == 0;
                    const uint32_t sz = is_lit ? 1u : 2u + ((e0 >> off) & 1u) + ((K == K_LZ11) ? 2u * ((e1 >> off) & 1u) : 0u);
                    if (lane == uint32_t(g * 8 + i)) {
                        myoff = pg + 1 + off;
                        mysz = sz;
                        mypg = pg;
                    }
                    off += sz;
                }
                pg += 1 + off;
            }
            next_cur = pg;
            ism = mysz >= 2;
            const uint32_t b1 = in[0].at(myoff), b2 = in[0].at(myoff + 1), b3 = in[0].at(myoff + 2);
            lit = b1;
            if (K == K_YAZ0) {
                dist = (((b1 & 0xF) << 8) | b2) + 1;
                // the extended length byte is read with Stream.ReadByte(): -1 at EOF -> 0x11 (Yay0.cs:131)
                const bool have_ext = myoff + 2 < slen;
                len = !ism ? 1 : (mysz == 3 ? (have_ext ? b3 + 0x12 : 0x11) : (b1 >> 4) + 2);
                tok_end = myoff + (ism ? 2 : 1);
                bad = mypg >= slen || tok_end > slen;
                if (mysz == 3 && have_ext) tok_end = myoff + 3;
            } else {
                const uint32_t b4 = in[0].at(myoff + 3);
                if (mysz == 3) {
                    dist = (((b2 & 0xF) << 8) | b3) + 1;
                    len = (((b1 & 0xF) << 4) | (b2 >> 4)) + 17;
                } else if (mysz == 4) {
                    dist = (((b3 & 0xF) << 8) | b4) + 1;
                    len = (((b1 & 0xF) << 12) | (b2 << 4) | (b3 >> 4)) + 273;
                } else {
                    dist = (((b1 & 0xF) << 8) | b2) + 1;
                    len = ism ? (b1 >> 4) + 1 : 1;
                }
                tok_end = myoff + mysz;
                bad = mypg >= slen || tok_end > slen;
            }
        } else {   // K_MIO0 / K_YAY0
            in[0].ensure(cur, 8);
            in[1].ensure(ccur, 72);
            in[2].ensure(lcur, 72);
            const uint32_t fw = (in[0].at(cur) << 24) | (in[0].at(cur + 1) << 16) | (in[0].at(cur + 2) << 8) | in[0].at(cur + 3);
            next_cur = cur + 4;
            ism = ((fw >> (31 - lane)) & 1) == 0;
            const uint32_t lits_before = lane ? __popc(fw >> (32 - lane)) : 0;
            const uint32_t m_before = lane - lits_before;
            const uint32_t crel = ccur + 2 * m_before;
            const uint32_t b1 = in[1].at(crel), b2 = in[1].at(crel + 1);
            dist = (((b1 & 0xF) << 8) | b2) + 1;
            uint32_t ext_before = 0;
            bool is_ext = false;
            if (K == K_YAY0) {
                is_ext = ism && (b1 >> 4) == 0;
                ext_before = __popc(__ballot_sync(kFull, is_ext) & lt_mask);
            }
            const uint32_t lrel = lcur + lits_before + ext_before;
            const uint32_t lb = in[2].at(lrel);
            lit = lb;
            const uint32_t fabs = 0x10 + cur + (lane >> 3), cabs = comp_off + crel, labs = lit_off + lrel;
            if (K == K_MIO0) {
                len = ism ? (b1 >> 4) + 3 : 1;
            } else {
                ext_used = is_ext && labs < slen;   // ReadByte() == -1 at EOF: nothing consumed, length 0x11
                len = !ism ? 1 : (is_ext ? (ext_used ? lb + 0x12 : 0x11) : (b1 >> 4) + 2);
            }
            bad = fabs >= slen || (ism ? cabs + 2 > slen : labs + 1 > slen);
        }

        // ------------------------------------------------------------------ output positions
        const uint32_t incl = warp_incl_scan(len);
        const uint32_t excl = incl - len;
        const uint32_t remaining = size - written;
        bool active = excl < remaining;
        const uint32_t badmask = __ballot_sync(kFull, active && bad);
        if (badmask) {
            const uint32_t first = __ffs(badmask) - 1;
            active = active && lane < first;
            status = AURORA_END_OF_STREAM;
        }
        const uint32_t amask = __ballot_sync(kFull, active);   // always a prefix of the lanes
        const uint32_t nact = __popc(amask);
        if (nact == 0) break;
        const uint32_t total = __shfl_sync(kFull, incl, nact - 1);
        const uint32_t pos = written + excl;   // output position of this token

        if constexpr (K == K_LZSS) {
            // LZSS.cs:122-126 + LzWindows.OffsetCopy (:108-115): absolute ring offset -> distance
            const uint32_t ring_len = 1u << lz.windows_bits;
            const uint32_t offset = (uint32_t(lz.max_distance) + dist - uint32_t(lz.windows_start)) & uint32_t(lz.max_distance - 1);
            const uint32_t rp = pos & (ring_len - 1);
            uint32_t d = rp >= offset ? rp - offset : rp - offset + ring_len;
            if (d == 0) d = ring_len;   // BackCopy(0, n) re-reads the slot it writes: the data one window back
            dist = d;
        }

        // source.Position after the last executed token
        if constexpr (K == K_MIO0 || K == K_YAY0) {
            ccur += 2 * __popc(__ballot_sync(kFull, active && ism));
            lcur += __popc(__ballot_sync(kFull, active && !ism)) + __popc(__ballot_sync(kFull, active && ext_used));
            consumed = max(comp_off + ccur, lit_off + lcur);
        } else {
            consumed = __shfl_sync(kFull, tok_end, nact - 1);
        }
        cur = next_cur;

        // ------------------------------------------------------------------ copy phase (sub-batches of <= kSubMax bytes)
        uint32_t a = 0;
        while (a < nact) {
            uint32_t b = nact;
            if constexpr (Traits<K>::kNeedSub) {
                const uint32_t base_excl = __shfl_sync(kFull, excl, a);
                const uint32_t inm = __ballot_sync(kFull, active && lane >= a && incl - base_excl <= uint32_t(kSubMax));
                b = a + __popc(inm);
                if (b == a) {
                    // a single token longer than kSubMax (LZ11): periodic copy in segments
                    const uint32_t dstp = __shfl_sync(kFull, pos, a);
                    const uint32_t l = __shfl_sync(kFull, len, a);
                    const uint32_t d = __shfl_sync(kFull, dist, a);
                    for (uint32_t s = 0; s < l; s += kSubMax) {
                        const uint32_t seg = min(uint32_t(kSubMax), l - s);
                        ring_copy_match(ring, dstp + s, d, seg);
                        __syncwarp();
                        out.drain(dstp + s + seg);
                    }
                    a++;
                    continue;
                }
            }
            const bool mine = active && lane >= a && lane < b;
            if (mine && !ism) ring[pos & kRingMask] = uint8_t(lit);
            __syncwarp();
            uint32_t mm = __ballot_sync(kFull, mine && ism);
            while (mm) {
                const int k = __ffs(mm) - 1;
                mm &= mm - 1;
                const uint32_t dstp = __shfl_sync(kFull, pos, k);
                const uint32_t ld = __shfl_sync(kFull, len | (dist << 17), k);
                ring_copy_match(ring, dstp, ld >> 17, ld & 0x1FFFFu);
                __syncwarp();
            }
            out.drain(written + __shfl_sync(kFull, incl, b - 1));
            a = b;
        }
        written += total;
        if (status != AURORA_OK) break;
    }
    out.finish(written);
    if (status == AURORA_OK) {
        if (K == K_LZSS ? written != size : written > size) status = AURORA_SIZE_MISMATCH;
    }
    return BodyResult{status, written, consumed};
}

// pre-history of the window: zeros (LzWindows.cs:53 rents an uncleared array) or LZSS initialFill
__device__ __forceinline__ void ring_prefill(uint8_t* ring, uint32_t fill) {
    const uint32_t w = fill * 0x01010101u;
    const uint4 v = make_uint4(w, w, w, w);
    uint4* p = reinterpret_cast<uint4*>(ring + (kRing - kWindow));
    for (int i = lane_id(); i < kWindow / 16; i += 32) p[i] = v;
    __syncwarp();
}

template <int K>
__device__ void decode_stream(const DecodeParams& P, uint32_t idx, InStream* in, uint8_t* ring) {
    const uint32_t lane = lane_id();
    const uint8_t* src = P.src_base + P.src_off[idx];
    const uint64_t slen64 = P.src_len[idx];
    const uint32_t slen = slen64 > 0xFFFFFFF0ull ? 0xFFFFFFF0u : uint32_t(slen64);
    uint8_t* dst = P.dst_base + P.dst_off[idx];
    const uint64_t cap = P.dst_cap[idx];

    // ---- header (<= 16 bytes), read straight from global memory
    const uint32_t hb = (lane < 16 && lane < slen) ? src[lane] : 0u;
    auto H = [&](int j) { return __shfl_sync(kFull, hb, j); };
    auto be32 = [&](int j) { return (H(j) << 24) | (H(j + 1) << 16) | (H(j + 2) << 8) | H(j + 3); };
    auto le32 = [&](int j) { return H(j) | (H(j + 1) << 8) | (H(j + 2) << 16) | (H(j + 3) << 24); };

    int status = AURORA_OK;
    uint32_t size = 0, body_off = 0, comp_off = 0, lit_off = 0, consumed = 0;
    bool yaz_retry = false;

    if (K == K_LZ10 || K == K_LZ11) {
        const uint32_t id = (K == K_LZ10) ? 0x10 : 0x11;
        if (slen < 1) { status = AURORA_END_OF_STREAM; consumed = slen; }
        else if (H(0) != id) { status = AURORA_INVALID_IDENTIFIER; consumed = 1; }
        else if (slen < 4) { status = AURORA_END_OF_STREAM; consumed = slen; }
        else {
            size = H(1) | (H(2) << 8) | (H(3) << 16);
            body_off = 4;
            if (size == 0) {
                if (slen < 8) { status = AURORA_END_OF_STREAM; consumed = slen; }
                else { size = le32(4); body_off = 8; }
            }
        }
    } else if (K == K_YAZ0 || K == K_LZSS || K == K_MIO0 || K == K_YAY0) {
        uint32_t magic;
        if (K == K_YAZ0) magic = P.format == AURORA_FMT_YAZ1 ? 0x59617A31u : 0x59617A30u;   // "Yaz1" / "Yaz0"
        else if (K == K_LZSS) magic = 0x4C5A5353u;                                           // "LZSS"
        else if (K == K_MIO0) magic = 0x4D494F30u;                                           // "MIO0"
        else magic = 0x59617930u;                                                            // "Yay0"
        if (slen < 4) { status = AURORA_END_OF_STREAM; consumed = slen; }
        else if (be32(0) != magic) { status = AURORA_INVALID_IDENTIFIER; consumed = 4; }
        else if (slen < 16) { status = AURORA_END_OF_STREAM; consumed = slen; }
        else {
            body_off = 16;
            if (K == K_LZSS) {
                size = be32(4);
            } else if (K == K_YAZ0) {
                const bool big = P.byte_order != AURORA_ENDIAN_LITTLE;   // Yaz0.cs:30 default Big
                size = big ? be32(4) : le32(4);
                yaz_retry = true;
            } else {
                bool big;
                if (P.byte_order == AURORA_ENDIAN_LITTLE) big = false;
                else if (P.byte_order == AURORA_ENDIAN_BIG) big = true;
                else {
                    // DetectByteOrder<uint>(3) stand-in: first plausible order, Big preferred
                    const uint32_t cb = be32(8), lb = be32(12), cl = le32(8), ll = le32(12);
                    const bool pb = cb >= 0x10 && cb <= lb && lb <= slen;
                    const bool pl = cl >= 0x10 && cl <= ll && ll <= slen;
                    big = pb || !pl;
                }
                size = big ? be32(4) : le32(4);
                comp_off = big ? be32(8) : le32(8);
                lit_off = big ? be32(12) : le32(12);
                if (K == K_YAY0 && P.byte_order == AURORA_ENDIAN_DEFAULT) (void)0;
                if (comp_off < 0x10 || comp_off > slen || lit_off < 0x10 || lit_off > slen) {
                    status = AURORA_INVALID_DATA;   // Span.Slice -> ArgumentOutOfRangeException
                    consumed = slen;
                }
            }
        }
    }

    uint32_t written = 0;
    if (status == AURORA_OK) {
        for (int attempt = 0; attempt < 2; attempt++) {
            if (uint64_t(size) > cap) {   // destination.SetLength on a non-expandable stream
                status = AURORA_DST_TOO_SMALL;
                written = 0;
                consumed = (K == K_MIO0 || K == K_YAY0) ? slen : body_off;
            } else {
                ring_prefill(ring, K == K_LZSS ? uint32_t(P.lzss.initial_fill) & 0xFFu : 0u);
                OutState out;
                out.ring = ring;
                out.dst = dst;
                out.limit = uint32_t(min(uint64_t(0xFFFFFFFFu), cap));
                out.flushed = 0;
                out.aligned = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
                if constexpr (K == K_MIO0 || K == K_YAY0) {
                    in[0].begin(P.src_base, P.src_limit, src + 0x10);
                    in[1].begin(P.src_base, P.src_limit, src + comp_off);
                    in[2].begin(P.src_base, P.src_limit, src + lit_off);
                } else {
                    in[0].begin(P.src_base, P.src_limit, src);
                }
                const BodyResult r = decode_body<K>(in, out, slen, size, body_off, comp_off, lit_off, P.lzss);
                status = r.status;
                written = r.written;
                consumed = r.consumed;
                if (status == AURORA_END_OF_STREAM) consumed = slen;
                if ((K == K_MIO0 || K == K_YAY0) && status != AURORA_OK) consumed = slen;
            }
            // Yaz0.cs:67-78: on any exception retry once with the byte-swapped size
            if (!(yaz_retry && attempt == 0 && status != AURORA_OK)) break;
            size = bswap32(size);
            status = AURORA_OK;
        }
    }
    if (lane == 0) {
        P.out_len[idx] = written;
        P.consumed[idx] = consumed;
        P.status[idx] = status;
    }
}

template <int K>
__global__ void __launch_bounds__(Traits<K>::kWarps * 32, 2) decode_flaglz_kernel(const DecodeParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5;
    uint8_t* wbase = smem + size_t(warp) * Traits<K>::kSmemPerWarp;
    uint8_t* ring = wbase;
    uint64_t* bars = reinterpret_cast<uint64_t*>(wbase + kRing + Traits<K>::kStreams * kInRing);
    InStream in[Traits<K>::kStreams];
#pragma unroll
    for (int s = 0; s < Traits<K>::kStreams; s++) in[s].init(wbase + kRing + s * kInRing, bars + 2 * s);
    __syncwarp();
    fence_proxy_async();

    for (;;) {
        uint32_t t = 0;
        if (lane_id() == 0) t = atomicAdd(P.ticket, 1u);
        t = __shfl_sync(kFull, t, 0);
        if (t >= P.n) break;
        const uint32_t idx = P.order ? P.order[t] : t;
        decode_stream<K>(P, idx, in, ring);
    }
#pragma unroll
    for (int s = 0; s < Traits<K>::kStreams; s++) in[s].drain_inflight();
}

template <int K>
cudaError_t launch(const DecodeParams& p, int sm_count, cudaStream_t st) {
    const int threads = Traits<K>::kWarps * 32;
    const size_t smem = size_t(Traits<K>::kWarps) * Traits<K>::kSmemPerWarp;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(decode_flaglz_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    const uint32_t warps_needed = p.n;
    int blocks = sm_count * 2;
    const int needed_blocks = int((warps_needed + Traits<K>::kWarps - 1) / Traits<K>::kWarps);
    if (needed_blocks < blocks) blocks = needed_blocks > 0 ? needed_blocks : 1;
    decode_flaglz_kernel<K><<<blocks, threads, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_decode_flaglz(const DecodeParams& p, int sm_count, cudaStream_t st) {
    switch (p.format) {
        case AURORA_FMT_LZ10: return launch<K_LZ10>(p, sm_count, st);
        case AURORA_FMT_LZ11: return launch<K_LZ11>(p, sm_count, st);
        case AURORA_FMT_YAZ0:
        case AURORA_FMT_YAZ1: return launch<K_YAZ0>(p, sm_count, st);
        case AURORA_FMT_LZSS: return launch<K_LZSS>(p, sm_count, st);
        case AURORA_FMT_MIO0: return launch<K_MIO0>(p, sm_count, st);
        case AURORA_FMT_YAY0: return launch<K_YAY0>(p, sm_count, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace aurora
