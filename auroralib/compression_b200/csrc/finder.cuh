// finder.cuh — exact device replay of LzChainMatchFinder (MatchFinder/LzChainMatchFinder.cs) and of FlagWriter
// (IO/FlagWriter.cs), shared by the encoder kernels.  One warp per stream; lane 0 owns the tables, the 32 lanes
// share the common-prefix comparison of each chain candidate.
#pragma once
#include "common.cuh"

namespace aurora {

struct Finder {
    int* head;
    int* chain;
    int* mint;
    int hash_bits, hash_mask, chain_mask, max_chain, lazy, min_len, max_len, min_dist, max_dist;
    uint32_t min_mask;
    bool no_self_overlap, has_min;
    int position;
};

__device__ __forceinline__ uint32_t load_u32le(const uint8_t* p) {
    return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24);
}

// LzChainMatchFinder.cs:288-299
__device__ __forceinline__ void compute_hash(const Finder& f, const uint8_t* d, int& h4, int& hm) {
    const uint32_t prim = 2654435761u;
    uint32_t v = load_u32le(d);
    uint32_t mn = v & f.min_mask;
    v *= prim;
    mn *= prim;
    h4 = int(v >> (32 - f.hash_bits)) & f.hash_mask;
    hm = int((mn >> 16) & 0xFFFF);
}

// :130-140 (lane 0 owns the tables; the values it reads back are broadcast by the callers)
__device__ __forceinline__ void finder_insert(Finder& f, int pos, int h4, int hm) {
    if (lane_id() == 0) {
        if (f.chain_mask != 0) f.chain[pos & f.chain_mask] = f.head[h4];
        f.head[h4] = pos;
        if (f.has_min) f.mint[hm] = pos;
    }
    __syncwarp();
}

// :338-357 — common prefix of data[a..] and data[b..], capped at max; 32 bytes per step
__device__ __forceinline__ int match_length(const uint8_t* data, int a, int b, int max) {
    const int lane = lane_id();
    for (int base = 0; base < max; base += 32) {
        const int i = base + lane;
        const bool eq = i < max && data[a + i] == data[b + i];
        const uint32_t ne = __ballot_sync(kFull, !eq);
        if (ne) return base + __ffs(ne) - 1;
    }
    return max;
}

// :214-282
static __device__ __forceinline__ void match_search(Finder& f, const uint8_t* data, int data_len, int pos, int& best_dist, int& best_len) {
    int h4, hm;
    compute_hash(f, data + pos, h4, hm);
    int cur = __shfl_sync(kFull, lane_id() == 0 ? f.head[h4] : 0, 0);
    const int best_possible = min(data_len - pos, f.max_len);
    best_dist = best_len = 0;
    int best_score = -1;
    int attempts = f.max_chain;
    while (cur != -1 && attempts-- > 0) {
        const int distance = pos - cur;
        if (distance > f.max_dist) break;
        int next = -1;
        if (f.chain_mask != 0) next = __shfl_sync(kFull, lane_id() == 0 ? f.chain[cur & f.chain_mask] : 0, 0);
        if (distance < f.min_dist) {
            cur = next;
            continue;
        }
        // CompatibilityMode cuts a match to its distance (:304): comparing further cannot change the result
        const int len = match_length(data, pos, cur, f.no_self_overlap ? min(best_possible, distance) : best_possible);
        const int score = len - f.min_len;
        if (score > best_score) {
            best_score = score;
            best_len = len;
            best_dist = distance;
            if (best_len == best_possible) break;
        }
        cur = next;
    }
    if (best_len == 0 && f.has_min) {
        cur = __shfl_sync(kFull, lane_id() == 0 ? f.mint[hm] : 0, 0);
        if (cur != -1) {
            int distance = pos - cur;
            if (distance < f.min_dist) distance = f.min_dist;
            // (a raised distance that points in front of the buffer is no match: the reference compares through an unsafe
            //  pointer against whatever memory precedes the source there — undefined; DESIGN.md section 2, deviation 6)
            if (distance <= f.max_dist && pos - distance >= 0) {
                best_len = match_length(data, pos, pos - distance, f.no_self_overlap ? min(best_possible, distance) : best_possible);
                best_dist = distance;
            }
        }
    }
    finder_insert(f, pos, h4, hm);
}

struct Match {
    int offset, distance, length;
};

// :157-212
static __device__ __forceinline__ Match find_next_best_match(Finder& f, const uint8_t* data, int length) {
    const int limit = length - 4;
    while (f.position <= limit) {
        int best_dist, best_len;
        match_search(f, data, length, f.position, best_dist, best_len);
        if (best_len < f.min_len) {
            f.position++;
            continue;
        }
        int skip = 0;
        if (best_len <= f.lazy && f.position + 1 <= limit) {
            const int next_pos = f.position + 1;
            int nd, nl;
            match_search(f, data, length, next_pos, nd, nl);
            if (nl > best_len) {
                best_len = nl;
                best_dist = nd;
                f.position = next_pos;
            } else {
                skip++;
            }
        }
        const Match m{f.position, best_dist, best_len};
        const int end = f.position + best_len;
        f.position++;
        f.position += skip;
        while (f.position < end && f.position <= limit) {
            int h4, hm;
            compute_hash(f, data + f.position, h4, hm);
            finder_insert(f, f.position, h4, hm);
            f.position++;
        }
        return m;
    }
    f.position = length;
    return Match{length, 0, 0};
}

// FlagWriter (IO/FlagWriter.cs) over a bounded output: the flag byte of a group is reserved when the group's
// first byte or bit arrives and patched when the 8th bit (or Dispose) comes — the same byte layout as the
// reference's "flag word, then the buffered token bytes".
struct Writer {
    uint8_t* out;
    uint64_t cap;
    uint64_t pos;
    int64_t flag_pos;
    uint32_t flag_val, bits;
    uint32_t flag_bytes = 1;   // FlagWriter flag size: 1, or 4 (big-endian word: LZHudson)
    bool msb_first, overflow;
    bool negate = false;       // the flag byte is written as (byte)-flag (LZ40.cs:137)
    __device__ __forceinline__ void put(uint64_t at, uint32_t b) {
        if (at < cap) {
            if (lane_id() == 0) out[at] = uint8_t(b);
        } else {
            overflow = true;
        }
    }
    __device__ __forceinline__ void raw_byte(uint32_t b) { put(pos++, b); }
    __device__ __forceinline__ void group() {
        if (flag_pos < 0) {
            flag_pos = int64_t(pos);
            pos += flag_bytes;
            flag_val = 0;
            bits = 0;
        }
    }
    __device__ __forceinline__ void put_flag() {
        if (flag_bytes == 1) {
            put(uint64_t(flag_pos), negate ? (0u - flag_val) & 0xFFu : flag_val);
        } else {
            for (uint32_t i = 0; i < flag_bytes; i++) put(uint64_t(flag_pos) + i, (flag_val >> (8 * (flag_bytes - 1 - i))) & 0xFF);
        }
        flag_pos = -1;
    }
    __device__ __forceinline__ void byte(uint32_t b) {
        group();
        put(pos++, b);
    }
    __device__ __forceinline__ void bit(bool v) {
        group();
        if (v) flag_val |= msb_first ? ((0x80u << (8 * (flag_bytes - 1))) >> bits) : (1u << bits);
        if (++bits == 8 * flag_bytes) put_flag();
    }
    __device__ __forceinline__ void dispose() {
        if (flag_pos >= 0) put_flag();
    }
};


// Reset (:121-128)
__device__ __forceinline__ void finder_reset(Finder& f) {
    const int lane = lane_id();
    const int4 m1 = make_int4(-1, -1, -1, -1);
    int4* h = reinterpret_cast<int4*>(f.head);
    for (int i = lane; i < (f.hash_mask + 1) / 4; i += 32) h[i] = m1;
    if (f.max_chain != 1) {
        int4* c = reinterpret_cast<int4*>(f.chain);
        for (int i = lane; i < (f.chain_mask + 1) / 4; i += 32) c[i] = m1;
    }
    if (f.has_min) {
        int4* t = reinterpret_cast<int4*>(f.mint);
        for (int i = lane; i < 65536 / 4; i += 32) t[i] = m1;
    }
    f.position = 0;
    __syncwarp();
    __threadfence_block();
}

// LzChainMatchFinder.cs:42-106 with the derived parameters of :108-119 (resolved on the host)
__device__ __forceinline__ void finder_setup(Finder& f, const EncodeParams& P, uint8_t* scratch) {
    f.hash_bits = P.hash_bits;
    f.hash_mask = (1 << P.hash_bits) - 1;
    f.max_chain = P.max_chain;
    f.chain_mask = P.max_chain == 1 ? 0 : (1 << P.chain_bits) - 1;
    f.lazy = P.lazy_threshold;
    f.min_len = P.min_length;
    f.max_len = P.max_length;
    f.min_dist = P.min_distance;
    f.max_dist = P.max_distance;
    f.no_self_overlap = P.no_self_overlap != 0;
    f.has_min = P.use_min_table != 0 && P.min_length < 4;
    f.min_mask = f.has_min ? 0xFFFFFFFFu >> ((4 - P.min_length) * 8) : 0u;
    f.head = reinterpret_cast<int*>(scratch);
    f.chain = f.head + (size_t(1) << P.hash_bits);
    f.mint = f.chain + (size_t(1) << P.chain_bits);
    f.position = 0;
}

}  // namespace aurora
