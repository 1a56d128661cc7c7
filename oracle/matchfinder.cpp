// matchfinder.cpp — CPU ORACLE (test infrastructure, NOT product code).
// Restatement of MatchFinder/LzChainMatchFinder.cs (single-LzProperties use only; the multi-property
// scoring path :301-321 is used by FastLZ/aPLib/RefPack/ALLZ, none of which is on the hot path).
#include <cstdlib>
#include <cmath>

#include "oracle_core.hpp"

namespace ora {

// LzChainMatchFinder.cs:111-119
static int GetMaxChain(int q) {
    if (q < 6) return q + 1;
    if (q >= 11) return 1 << (q - 5);
    int baseVal = 1 << (q >> 1);
    return baseVal | (baseVal >> (q & 1));
}

static int isqrt2q(int q) {   // (int)Math.Sqrt(2 * Quality), exact for 0..30
    int r = 0;
    while ((r + 1) * (r + 1) <= 2 * q) r++;
    return r;
}

// LzChainMatchFinder.cs:42-106 (ctor) with the parameter derivation of :108-109
MatchFinder::MatchFinder(const LzProps& p, const Settings& s) {
    const int q = s.Quality;
    int maxChain = GetMaxChain(q);
    int lazyThreshold = 3 + (q / 3);
    int hashBits = 15 + isqrt2q(q);
    int maxChainSizeBits = 17 + isqrt2q(q);
    int maxWindowBits = s.MaxWindowBits;
    bool useMinTable = q >= 10;
    // (investigation knob, tests/test_oracle_golden.py: Benchmarks.md's Q15 columns of the 4 KiB-window formats)
    if (std::getenv("ORACLE_NO_MINTABLE")) useMinTable = false;
    if (const char* e = std::getenv("ORACLE_LAZY")) lazyThreshold = std::atoi(e);
    if (const char* e = std::getenv("ORACLE_MAXCHAIN")) maxChain = std::atoi(e);
    if (const char* e = std::getenv("ORACLE_HASHBITS")) hashBits = std::atoi(e);

    minMatchLength_ = p.MinLength;
    maxMatchLength_ = p.MaxLength;
    minDistance_ = p.MinDistance;
    maxDistance_ = p.MaxDistance;
    int windowsBits = 1;
    if (windowsBits < p.WindowsBits) windowsBits = p.WindowsBits;
    if (maxWindowBits != 0) {
        windowsBits = std::max(windowsBits, maxWindowBits);
        maxDistance_ = std::max(maxDistance_, 1 << maxWindowBits);
    }
    lazyThreshold_ = lazyThreshold;
    noSelfOverlap_ = (s.Strategy & 1) != 0;

    hashBits_ = hashBits;
    hashMask_ = (1 << hashBits) - 1;
    head_.assign(size_t(1) << hashBits, -1);

    maxChain_ = maxChain;
    if (maxChain_ == 1) {
        chain_.assign(1, -1);   // NoChainTable
        chainMask_ = 0;
    } else {
        maxChainSizeBits = std::min(maxChainSizeBits, windowsBits);
        chain_.assign(size_t(1) << maxChainSizeBits, -1);
        chainMask_ = (1 << maxChainSizeBits) - 1;
    }
    if (useMinTable && minMatchLength_ < 4) {
        minMask_ = 0xFFFFFFFFu >> ((4 - minMatchLength_) * 8);
        mint_.assign(65536, -1);
        hasMin_ = true;
    }
    Reset();
}

// :121-128
void MatchFinder::Reset() {
    Position = 0;
    std::fill(head_.begin(), head_.end(), -1);
    if (maxChain_ != 1) std::fill(chain_.begin(), chain_.end(), -1);
    if (hasMin_) std::fill(mint_.begin(), mint_.end(), -1);
}

// :130-140
void MatchFinder::Insert(int pos, int h4, int hm) {
    if (chainMask_ != 0) chain_[size_t(pos & chainMask_)] = head_[size_t(h4)];
    head_[size_t(h4)] = pos;
    if (hasMin_) mint_[size_t(hm)] = pos;
}

// :288-299
void MatchFinder::ComputeHash(const uint8_t* d, int& h4, int& hm) const {
    const uint32_t prim = 2654435761u;
    uint32_t v = uint32_t(d[0]) | uint32_t(d[1]) << 8 | uint32_t(d[2]) << 16 | uint32_t(d[3]) << 24;
    uint32_t mn = v & minMask_;
    v *= prim;
    mn *= prim;
    h4 = int(v >> (32 - hashBits_)) & hashMask_;
    hm = int((mn >> 16) & 0xFFFF);
}

// :301-321 (single-property branch)
int MatchFinder::ScoreMatch(int& length, int distance) const {
    if (noSelfOverlap_ && length > distance) length = distance;
    return length - minMatchLength_;
}

// :338-357
int MatchFinder::GetMatchLength(const uint8_t* a, const uint8_t* b, int max) {
    int len = 0;
    while (len + 8 <= max) {
        uint64_t x, y;
        std::memcpy(&x, a + len, 8);
        std::memcpy(&y, b + len, 8);
        uint64_t diff = x ^ y;
        if (diff != 0) return len + __builtin_ctzll(diff) / 8;
        len += 8;
    }
    while (len < max && a[len] == b[len]) len++;
    return len;
}

// :248-282
void MatchFinder::ChainMatches(const uint8_t* data, int bestPossible, int pos, int cur, int attempts, int& bestDistance, int& bestLength) {
    const uint8_t* dataPos = data + pos;
    bestDistance = bestLength = 0;
    int bestScore = -1;
    while (cur != -1 && attempts-- > 0) {
        int distance = pos - cur;
        if (distance > maxDistance_) break;
        if (distance < minDistance_) {
            cur = GetNext(cur);
            continue;
        }
        int len = GetMatchLength(dataPos, data + cur, bestPossible);
        int score = ScoreMatch(len, distance);
        if (score > bestScore) {
            bestScore = score;
            bestLength = len;
            bestDistance = distance;
            if (bestLength == bestPossible) break;
        }
        cur = GetNext(cur);
    }
}

// :214-246
void MatchFinder::MatchSearch(const uint8_t* data, int dataLength, int pos, int attempts, int& bestDistance, int& bestLength) {
    const uint8_t* dataPos = data + pos;
    int h4, hm;
    ComputeHash(dataPos, h4, hm);
    int cur = head_[size_t(h4)];
    int bestPossible = std::min(dataLength - pos, maxMatchLength_);
    ChainMatches(data, bestPossible, pos, cur, attempts, bestDistance, bestLength);
    if (bestLength == 0 && hasMin_) {
        cur = mint_[size_t(hm)];
        if (cur != -1) {
            int distance = pos - cur;
            if (distance < minDistance_) distance = minDistance_;
            // Raising the distance to MinDistance can point in front of the buffer (pos < MinDistance: LZ10 / LZ11 VRAM mode,
            // BLZ): the reference compares through an unsafe pointer there, i.e. against whatever memory precedes the source
            // (undefined, nondeterministic).  Here — and in the GPU finder — such a candidate is no match.
            if (distance <= maxDistance_ && pos - distance >= 0) {
                bestLength = GetMatchLength(dataPos, data + pos - distance, bestPossible);
                (void)ScoreMatch(bestLength, distance);
                bestDistance = distance;
            }
        }
    }
    Insert(pos, h4, hm);
}

// :157-212
LzMatch MatchFinder::FindNextBestMatch(const uint8_t* data, int length) {
    const int limit = length - 4;
    const int maxChain = maxChain_;
    while (Position <= limit) {
        int bestDistance, bestLength;
        MatchSearch(data, length, Position, maxChain, bestDistance, bestLength);
        if (bestLength < minMatchLength_) {
            Position++;
            continue;
        }
        int skip = 0;
        if (bestLength <= lazyThreshold_ && Position + 1 <= limit) {
            int nextPos = Position + 1;
            int nextDistance, nextLength;
            MatchSearch(data, length, nextPos, maxChain, nextDistance, nextLength);
            if (nextLength > bestLength) {
                bestLength = nextLength;
                bestDistance = nextDistance;
                Position = nextPos;
            } else {
                skip++;
            }
        }
        LzMatch match{Position, bestDistance, bestLength};
        int end = Position + bestLength;
        Position++;
        Position += skip;
        while (Position < end && Position <= limit) {
            int h4, hm;
            ComputeHash(data + Position, h4, hm);
            Insert(Position++, h4, hm);
        }
        return match;
    }
    Position = length;
    return LzMatch{length, 0, 0};
}

}  // namespace ora
