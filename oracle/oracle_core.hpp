// oracle_core.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A C++ restatement of the primitives of Venomalia/AuroraLib.Compression that the LZ hot path is
// built on.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may use anything under oracle/.  The product (libaurora_cuda.so) never links or calls it.
//
// Restated from (paths relative to /root/reference/src/AuroraLib.Compression):
//   IO/LzWindows.cs        ring-buffer output window  (BackCopy :72-100, OffsetCopy :108-115,
//                           CopyFrom :124-135, Write :162-185, InternWrite :187-227, WriteByte :232-237)
//   IO/FlagReader.cs       lazy flag-word bit reader   (Readbit :53-65, ReadInt :75-100)
//   IO/FlagWriter.cs       flag-word bit writer        (WriteBit :70-80, Flush :111-127, FlushIfNecessary :132-139)
//   LzProperties.cs        ctor A :46-55, ctor B :57-66
//   CompressionSettings.cs quality/maxWindowBits/strategy :38-50
// plus the AuroraLib.Core 1.7.0 stream helpers whose source is not in the tree (ReadUInt8 throws at
// EOF, BCL ReadByte returns -1, ReadUInt24/32 default little-endian) as used at the call sites.
//
// One documented deviation: the reference rents its ring from ArrayPool and never clears it
// (LzWindows.cs:53), so back-references before the start of the output read nondeterministic bytes.
// The oracle (and the GPU path) define that pre-history as zeros (or LZSS initialFill).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

namespace ora {

enum Status : int {
    OK = 0, END_OF_STREAM = 1, INVALID_IDENTIFIER = 2, SIZE_MISMATCH = 3, DST_TOO_SMALL = 4,
    INVALID_DATA = 5, NOT_SUPPORTED = 6, INVALID_ARGUMENT = 7, CUDA_ERROR = 8
};

enum Format : int {
    FMT_YAZ0 = 1, FMT_YAZ1 = 2, FMT_YAY0 = 3, FMT_MIO0 = 4, FMT_LZ10 = 5, FMT_LZ11 = 6, FMT_LZSS = 7,
    FMT_LZ4 = 8, FMT_LZ4_BLOCK = 9, FMT_LZ4_LEGACY = 10, FMT_LZO = 11, FMT_SNAPPY = 12,
    FMT_SNAPPY_BLOCK = 13, FMT_PRS = 14,
    FMT_GCLZ = 15, FMT_CXLZ = 16, FMT_COMP = 17, FMT_LZ_3DS = 18, FMT_LZ77 = 19, FMT_LEVEL5 = 20, FMT_LZON = 21, FMT_LEVEL5_LZSS = 22,
    FMT_AKLZ = 23, FMT_LZ01 = 24, FMT_FCMP = 25, FMT_IECP = 26, FMT_MDB4 = 27, FMT_LZSEGA = 28, FMT_GCZ = 29, FMT_SDPC = 30,
    FMT_ECD = 31, FMT_LZ00 = 32,
    FMT_LZHUDSON = 33, FMT_LZ40 = 34, FMT_LZ60 = 35, FMT_SMSR00 = 36, FMT_BLZ = 37
};

struct Error {
    int status;
    long long expected = 0, actual = 0;
};
[[noreturn]] inline void fail(int st, long long e = 0, long long a = 0) { throw Error{st, e, a}; }

enum class Endian : int { Little = 0, Big = 1 };

// ---------------------------------------------------------------------------------------------
// Source stream: System.IO.Stream over a read-only span, as the reference's decoders consume it.
// ---------------------------------------------------------------------------------------------
struct Src {
    const uint8_t* p;
    int64_t len;
    int64_t pos = 0;
    Src(const uint8_t* p_, int64_t n) : p(p_), len(n) {}

    int ReadByte() { return pos < len ? p[pos++] : -1; }                 // BCL Stream.ReadByte
    int PeekByte() const { return pos < len ? p[pos] : -1; }
    uint8_t ReadUInt8() {                                                 // AuroraLib.Core: throws at EOF
        if (pos >= len) fail(END_OF_STREAM);
        return p[pos++];
    }
    void need(int64_t n) { if (pos < 0 || pos + n > len) { pos = std::max<int64_t>(pos, len); fail(END_OF_STREAM); } }
    uint16_t ReadUInt16(Endian e) {
        need(2);
        uint16_t v = e == Endian::Big ? uint16_t(p[pos] << 8 | p[pos + 1]) : uint16_t(p[pos] | p[pos + 1] << 8);
        pos += 2;
        return v;
    }
    uint32_t ReadUInt24(Endian e = Endian::Little) {
        need(3);
        uint32_t v = e == Endian::Big ? uint32_t(p[pos] << 16 | p[pos + 1] << 8 | p[pos + 2])
                                      : uint32_t(p[pos] | p[pos + 1] << 8 | p[pos + 2] << 16);
        pos += 3;
        return v;
    }
    uint32_t ReadUInt32(Endian e = Endian::Little) {
        need(4);
        uint32_t v = e == Endian::Big
                         ? (uint32_t(p[pos]) << 24 | uint32_t(p[pos + 1]) << 16 | uint32_t(p[pos + 2]) << 8 | p[pos + 3])
                         : (uint32_t(p[pos]) | uint32_t(p[pos + 1]) << 8 | uint32_t(p[pos + 2]) << 16 | uint32_t(p[pos + 3]) << 24);
        pos += 4;
        return v;
    }
    uint64_t ReadUInt64LE() {
        uint64_t lo = ReadUInt32(), hi = ReadUInt32();
        return lo | hi << 32;
    }
    void ReadExactly(uint8_t* dst, int64_t n) {
        need(n);
        std::memcpy(dst, p + pos, size_t(n));
        pos += n;
    }
    bool Match(const void* magic, int n) {                                // consumes on success AND on failure
        if (pos + n > len) { return false; }
        bool ok = std::memcmp(p + pos, magic, size_t(n)) == 0;
        pos += n;
        return ok;
    }
    void MatchThrow(const void* magic, int n) {
        if (pos + n > len) { pos = len; fail(END_OF_STREAM); }
        if (!Match(magic, n)) fail(INVALID_IDENTIFIER);
    }
};

// ---------------------------------------------------------------------------------------------
// Destination stream with a fixed capacity (a non-expandable MemoryStream / the batch dst slot).
// Bytes past the capacity are counted but dropped; `pos` is what destination.Position would be
// on an expandable stream.
// ---------------------------------------------------------------------------------------------
struct Sink {
    uint8_t* p;
    int64_t cap;
    int64_t pos = 0;
    int64_t length = 0;
    bool size_only = false;   // size-scan pre-pass: count, never store, never refuse a length
    Sink(uint8_t* p_, int64_t c) : p(p_), cap(c) {}
    void SetLength(int64_t n) {
        if (n > cap && !size_only) fail(DST_TOO_SMALL);
        length = n;
    }
    void Write(const uint8_t* s, int64_t n) {
        if (n <= 0) return;
        if (pos < cap) std::memcpy(p + pos, s, size_t(std::min(n, cap - pos)));
        pos += n;
        length = std::max(length, pos);
    }
    void WriteByte(uint8_t b) { Write(&b, 1); }
    bool overflowed() const { return length > cap; }
};

// Growable output for the encoders (MemoryPoolStream).
struct OutBuf {
    std::vector<uint8_t> v;
    void WriteByte(uint8_t b) { v.push_back(b); }
    void Write(const uint8_t* s, size_t n) { v.insert(v.end(), s, s + n); }
    void WriteU16(uint16_t x, Endian e) {
        if (e == Endian::Big) { WriteByte(uint8_t(x >> 8)); WriteByte(uint8_t(x)); }
        else { WriteByte(uint8_t(x)); WriteByte(uint8_t(x >> 8)); }
    }
    void WriteU24(uint32_t x, Endian e) {
        if (e == Endian::Big) { WriteByte(uint8_t(x >> 16)); WriteByte(uint8_t(x >> 8)); WriteByte(uint8_t(x)); }
        else { WriteByte(uint8_t(x)); WriteByte(uint8_t(x >> 8)); WriteByte(uint8_t(x >> 16)); }
    }
    void WriteU32(uint32_t x, Endian e = Endian::Little) {
        if (e == Endian::Big) { WriteU16(uint16_t(x >> 16), e); WriteU16(uint16_t(x), e); }
        else { WriteU16(uint16_t(x), e); WriteU16(uint16_t(x >> 16), e); }
    }
    size_t size() const { return v.size(); }
    void PatchU32(size_t at, uint32_t x, Endian e) {
        for (int i = 0; i < 4; i++) v[at + i] = uint8_t(e == Endian::Big ? x >> (24 - 8 * i) : x >> (8 * i));
    }
};

// ---------------------------------------------------------------------------------------------
// LzProperties (LzProperties.cs)
// ---------------------------------------------------------------------------------------------
struct LzProps {
    int WindowsBits = 0, LengthBits = 0, MinLength = 0, MaxLength = 0, MaxDistance = 0, MinDistance = 1, WindowsStart = 0;
    static int ceil_log2(long long x) {   // (byte)Math.Ceiling(Math.Log(x, 2)) for the integers used here
        int b = 0;
        while ((1LL << b) < x) b++;
        return b;
    }
    // ctor A (LzProperties.cs:46-55)
    static LzProps Window(int windowsSize, int maxLength, int minLength = 3, int windowsStart = 0, int minDistance = 1) {
        LzProps p;
        p.WindowsBits = ceil_log2(windowsSize);
        p.LengthBits = ceil_log2((long long)maxLength - minLength) & 0xFF;
        p.MinLength = minLength;
        p.MaxDistance = windowsSize;
        p.MaxLength = maxLength;
        p.WindowsStart = windowsStart;
        p.MinDistance = minDistance;
        return p;
    }
    // ctor B (LzProperties.cs:57-66)
    static LzProps Bits(int distanceBits, int lengthBits, int threshold = 2) {
        LzProps p;
        p.WindowsBits = distanceBits;
        p.LengthBits = lengthBits;
        p.MinLength = threshold + 1;
        p.MaxDistance = 1 << distanceBits;
        p.MaxLength = (1 << lengthBits) + threshold;
        p.WindowsStart = p.MaxDistance - (1 << lengthBits) - threshold;
        p.MinDistance = 1;
        return p;
    }
    int GetWindowsFlag() const { return MaxDistance - 1; }
    int GetLengthBitsFlag() const { return (1 << LengthBits) - 1; }
};

struct Settings {
    int Quality = 8;          // default(CompressionSettings) maps to Balanced (CompressionSettings.cs:18-19)
    int MaxWindowBits = 0;
    int Strategy = 0;         // bit0 = CompatibilityMode
};

// ---------------------------------------------------------------------------------------------
// LzWindows (IO/LzWindows.cs): power-of-two ring in front of the destination stream.
// ---------------------------------------------------------------------------------------------
class LzWindows {
    std::vector<uint8_t> buf_;
    int len_;
    int pos_ = 0;
    Sink* dst_;
    bool disposed_ = false;

  public:
    LzWindows(Sink* dst, int windowsBits, uint8_t fill = 0) : buf_(size_t(1) << windowsBits, fill), len_(1 << windowsBits), dst_(dst) {}
    ~LzWindows() { Dispose(); }
    int Position() const { return pos_; }
    int Length() const { return len_; }

    // LzWindows.cs:72-100
    void BackCopy(int distance, int length) {
        const int bufferLength = len_, mask = len_ - 1;
        while (length > 0) {
            int chunk = length;
            int srcPos = (pos_ - distance) & mask;
            if (distance < length && distance != 0) chunk = distance;
            if (srcPos + chunk > bufferLength) chunk = bufferLength - srcPos;
            InternWrite(buf_.data() + srcPos, chunk);
            length -= chunk;
        }
    }
    // LzWindows.cs:108-115
    void OffsetCopy(int offset, int length) {
        int distance = pos_ >= offset ? pos_ - offset : pos_ - offset + len_;
        BackCopy(distance, length);
    }
    // LzWindows.cs:124-135
    void CopyFrom(Src& source, int length) {
        while (length != 0) {
            int l = std::min(length, len_ - pos_);
            source.ReadExactly(buf_.data() + pos_, l);
            pos_ = (pos_ + l) & (len_ - 1);
            length -= l;
            if (pos_ == 0) FlushToDestination(len_);
        }
    }
    // LzWindows.cs:162-185
    void Write(const uint8_t* s, int n) {
        if (n < len_) { InternWrite(s, n); return; }
        int offset = 0;
        while (n - offset >= len_) { InternWrite(s + offset, len_); offset += len_; }
        if (offset < n) InternWrite(s + offset, n - offset);
    }
    // LzWindows.cs:232-237
    void WriteByte(uint8_t v) {
        buf_[size_t(pos_)] = v;
        pos_ = (pos_ + 1) & (len_ - 1);
        if (pos_ == 0) FlushToDestination(len_);
    }
    // LzWindows.cs:269-278
    void Dispose() {
        if (disposed_) return;
        disposed_ = true;
        if (pos_ != 0) FlushToDestination(pos_);
    }

  private:
    void FlushToDestination(int n) { if (dst_) dst_->Write(buf_.data(), n); }
    // LzWindows.cs:187-227.  `src` may alias the ring itself (BackCopy); memmove keeps the
    // non-overlapping-chunk semantics of Unsafe.CopyBlockUnaligned for the chunks BackCopy forms.
    void InternWrite(const uint8_t* src, int len) {
        if (len == 0) return;
        uint8_t* dst = buf_.data();
        if (len_ > pos_ + len) {
            std::memmove(dst + pos_, src, size_t(len));
            pos_ += len;
        } else {
            int left = len_ - pos_;
            int remaining = len - left;
            if (left > 0) std::memmove(dst + pos_, src, size_t(left));
            dst_->Write(dst, len_);
            if (remaining != 0) std::memmove(dst, src + len - remaining, size_t(remaining));
            pos_ = remaining;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// FlagReader (IO/FlagReader.cs)
// ---------------------------------------------------------------------------------------------
class FlagReader {
    Src* base_;
    int flagSize_;
    bool bitOrderBe_;
    Endian byteOrder_;
    int current_ = 0;

  public:
    int BitsLeft = 0;
    FlagReader(Src* source, Endian bitOrder, int flagSizeBytes = 1, Endian byteOrder = Endian::Little)
        : base_(source), flagSize_(flagSizeBytes * 8), bitOrderBe_(bitOrder == Endian::Big), byteOrder_(byteOrder) {}
    bool Readbit() {
        if (BitsLeft == 0) {
            switch (flagSize_) {
                case 8: current_ = (int8_t)base_->ReadUInt8(); break;
                case 16: current_ = base_->ReadUInt16(byteOrder_); break;
                case 24: current_ = int(base_->ReadUInt24(byteOrder_)); break;
                default: current_ = int(base_->ReadUInt32(byteOrder_)); break;
            }
            BitsLeft = flagSize_;
        }
        int shift = bitOrderBe_ ? BitsLeft - 1 : flagSize_ - BitsLeft;
        BitsLeft--;
        return (current_ & (1 << shift)) != 0;
    }
    int ReadInt(int bits, bool reverseOrder = false) {
        int value = 0;
        if (!reverseOrder) {
            for (int i = 0; i < bits; i++) if (Readbit()) value |= 1 << i;
        } else {
            for (int i = 0; i < bits; i++) { value <<= 1; if (Readbit()) value |= 1; }
        }
        return value;
    }
};

// ---------------------------------------------------------------------------------------------
// FlagWriter (IO/FlagWriter.cs)
// ---------------------------------------------------------------------------------------------
class FlagWriter {
    OutBuf* base_;
    int flagSize_;
    bool bitOrderBe_;
    Endian byteOrder_;
    int current_ = 0;

  public:
    int BitsLeft;
    OutBuf Buffer;
    bool negate8 = false;   // WriteFlagDelegate = i => destination.WriteByte((byte)-i)  (LZ40.cs:137)
    FlagWriter(OutBuf* destination, Endian bitOrder, int flagSizeBytes = 1, Endian byteOrder = Endian::Little)
        : base_(destination), flagSize_(8 * flagSizeBytes), bitOrderBe_(bitOrder == Endian::Big), byteOrder_(byteOrder), BitsLeft(8 * flagSizeBytes) {}
    void WriteBit(bool bit) {
        if (bit) {
            int shift = bitOrderBe_ ? BitsLeft - 1 : flagSize_ - BitsLeft;
            current_ |= (1 << shift);
        }
        BitsLeft--;
        if (BitsLeft == 0) Flush();
    }
    void WriteInt(int value, int bits, bool reverseOrder = false) {
        if (!reverseOrder) for (int i = 0; i < bits; i++) WriteBit(((value >> i) & 1) == 1);
        else for (int i = bits - 1; i >= 0; i--) WriteBit(((value >> i) & 1) == 1);
    }
    void Flush() {
        if (BitsLeft != flagSize_) {
            switch (flagSize_) {
                case 8: base_->WriteByte(negate8 ? uint8_t(-current_) : uint8_t(current_)); break;
                case 16: base_->WriteU16(uint16_t(current_), byteOrder_); break;
                case 24: base_->WriteU24(uint32_t(current_), byteOrder_); break;
                default: base_->WriteU32(uint32_t(current_), byteOrder_); break;
            }
            BitsLeft = flagSize_;
            current_ = 0;
        }
        if (Buffer.size() != 0) {
            // Buffer may alias *base_ only in Yaz0 (flag.Buffer is all three sub-streams); there the
            // buffer is a distinct object from the destination, so a plain append is right.
            base_->Write(Buffer.v.data(), Buffer.v.size());
            Buffer.v.clear();
        }
    }
    void FlushIfNecessary() {
        if (BitsLeft == flagSize_ && Buffer.size() != 0) {
            base_->Write(Buffer.v.data(), Buffer.v.size());
            Buffer.v.clear();
        }
    }
    void Dispose() { Flush(); }
};

// ---------------------------------------------------------------------------------------------
// LzChainMatchFinder (MatchFinder/LzChainMatchFinder.cs) — see matchfinder.cpp
// ---------------------------------------------------------------------------------------------
struct LzMatch {
    int Offset, Distance, Length;
};

class MatchFinder {
    int minMatchLength_, maxMatchLength_, minDistance_, maxDistance_;
    int chainMask_, hashBits_, hashMask_, maxChain_;
    uint32_t minMask_ = 0;
    int lazyThreshold_;
    bool noSelfOverlap_;
    std::vector<int> head_, chain_, mint_;
    bool hasMin_ = false;
    int Position = 0;

  public:
    MatchFinder(const LzProps& p, const Settings& s);
    void Reset();
    LzMatch FindNextBestMatch(const uint8_t* data, int length);

  private:
    void Insert(int pos, int h4, int hm);
    void ComputeHash(const uint8_t* d, int& h4, int& hm) const;
    void MatchSearch(const uint8_t* data, int dataLength, int pos, int attempts, int& bestDistance, int& bestLength);
    void ChainMatches(const uint8_t* data, int bestPossible, int pos, int cur, int attempts, int& bestDistance, int& bestLength);
    int GetNext(int pos) const { return chain_[size_t(pos & chainMask_)]; }
    int ScoreMatch(int& length, int distance) const;
    static int GetMatchLength(const uint8_t* a, const uint8_t* b, int max);
};

// ---------------------------------------------------------------------------------------------
// hashes used by the containers / KATs (public algorithms restated from their specifications)
// ---------------------------------------------------------------------------------------------
uint32_t xxh32(const uint8_t* p, size_t n, uint32_t seed = 0);
uint64_t xxh64(const uint8_t* p, size_t n, uint64_t seed = 0);
uint32_t crc32c(const uint8_t* p, size_t n);

// ---------------------------------------------------------------------------------------------
// codec options mirrored from include/aurora_cuda.h (kept layout-compatible; see capi.cpp)
// ---------------------------------------------------------------------------------------------
struct CodecOpts {
    Endian byteOrder = Endian::Big;
    bool byteOrderDefault = true;
    Settings settings;
    int vramMode = -1;
    LzProps lzss = LzProps::Bits(12, 4, 2);
    int lzssInitialFill = 0;
    uint32_t lz4BlockSize = 0x400000;
    bool lz4Verify = false;
    uint32_t yaz0Alignment = 0;
    int lz77Type = 0x10, lz77ChunkSize = 0x1000, level5Type = 1;   // LZ77.Type / LZ77.ChunkSize / Level5.Type
    uint32_t lz00Key = 0;       // LZ00.Compress(source, destination, key, settings)
    int ecdPlainSize = 4;       // ECD.PlainSize
};

struct DecodeResult {
    int status = OK;
    int64_t out_len = 0;
    int64_t consumed = 0;
};

// per-format entry points (formats_*.cpp).  Decoders throw ora::Error; the dispatcher catches.
bool is_wrapper_format(int fmt);
void wrapper_decode(int fmt, Src& s, Sink& d, const CodecOpts& o);
void wrapper_encode(int fmt, const uint8_t* src, int n, OutBuf& out, const CodecOpts& o, int lz77_type, int chunk_size, int level5_type);
uint32_t wrapper_decoded_size(int fmt, Src& s);
void lz10_decode(Src& s, Sink& d);
void lz11_decode(Src& s, Sink& d);
void yaz0_decode(Src& s, Sink& d, const CodecOpts& o, const char* magic);
void yay0_decode(Src& s, Sink& d, const CodecOpts& o);
void lz40_decode(Src& s, Sink& d, uint8_t id);
void lz40_encode(const uint8_t* src, int n, OutBuf& out, const CodecOpts& o, uint8_t id);
void blz_decode(Src& s, Sink& d);
void blz_encode(const uint8_t* src, int n, OutBuf& out, const CodecOpts& o);
void smsr00_decode(Src& s, Sink& d);
void smsr00_encode(const uint8_t* src, int n, OutBuf& out, const CodecOpts& o);
void lzhudson_decode(Src& s, Sink& d);
void lzhudson_encode(const uint8_t* src, int n, OutBuf& out, const CodecOpts& o);
void mio0_decode(Src& s, Sink& d, const CodecOpts& o);
void lzss_decode(Src& s, Sink& d, const CodecOpts& o);
void lz4_decode(Src& s, Sink& d, const CodecOpts& o);
void lz4_block_decode(Src& s, Sink& d);
void lzo_decode(Src& s, Sink& d);
void snappy_decode(Src& s, Sink& d);
void snappy_block_decode(Src& s, Sink& d);
void prs_decode(Src& s, Sink& d);

void lz10_encode(const uint8_t* src, int n, OutBuf& out, const CodecOpts& o);
void lz11_encode(const uint8_t* src, int n, OutBuf& out, const CodecOpts& o);
void yaz0_encode(const uint8_t* src, int n, OutBuf& out, const CodecOpts& o, const char* magic);
void yay0_encode(const uint8_t* src, int n, OutBuf& out, const CodecOpts& o);
void mio0_encode(const uint8_t* src, int n, OutBuf& out, const CodecOpts& o);
void lzss_encode(const uint8_t* src, int n, OutBuf& out, const CodecOpts& o);
void lz4_encode(const uint8_t* src, int n, OutBuf& out, const CodecOpts& o, bool legacy);
void lz4_block_encode(const uint8_t* src, int n, OutBuf& out, const CodecOpts& o);
void lzo_encode(const uint8_t* src, int n, OutBuf& out, const CodecOpts& o);
void snappy_encode(const uint8_t* src, int n, OutBuf& out, const CodecOpts& o);
void snappy_block_encode(const uint8_t* src, int n, OutBuf& out, const CodecOpts& o);
void prs_encode(const uint8_t* src, int n, OutBuf& out, const CodecOpts& o);

// IProvidesDecompressedSize / IsMatch
uint32_t decoded_size(int fmt, Src& s, const CodecOpts& o);   // throws NOT_SUPPORTED where absent
bool is_match(int fmt, Src& s, const CodecOpts& o);

DecodeResult decode_one(int fmt, const CodecOpts& o, const uint8_t* src, int64_t n, uint8_t* dst, int64_t cap, bool size_only = false);
int encode_one(int fmt, const CodecOpts& o, const uint8_t* src, int64_t n, std::vector<uint8_t>& out);

}  // namespace ora
