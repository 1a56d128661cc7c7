// aurora_codecs.hpp — header-only C++ host mirror of the reference's operator interface over the C ABI.
//
// The reference is compiled (managed C#) code and its toolchain is absent from the build image, so this is the
// compiled-language host side of the drop-in boundary: the same class names, method names, argument meaning and error
// behaviour as /root/reference/src/AuroraLib.Compression (ICompressionAlgorithm = ICompressionDecoder
// [Interfaces/ICompressionDecoder.cs:24] + ICompressionEncoder [Interfaces/ICompressionEncoder.cs:19],
// IProvidesDecompressedSize [:20], IEndianDependentFormat [:13]), with std::istream / std::ostream standing in for
// System.IO.Stream.  Every call is a 1-element batch through libaurora_cuda.so; there is no CPU implementation here.
#pragma once
#include <algorithm>
#include <cstdint>
#include <istream>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "aurora_cuda.h"

namespace aurora {

// ---- exception taxonomy (Exceptions/DecompressedSizeException.cs:8-25) ----
struct EndOfStreamException : std::runtime_error { EndOfStreamException() : std::runtime_error("EndOfStreamException") {} };
struct InvalidIdentifierException : std::runtime_error { InvalidIdentifierException() : std::runtime_error("InvalidIdentifierException") {} };
struct DecompressedSizeException : std::runtime_error {
    long long expected, actual;
    DecompressedSizeException(long long e, long long a)
        : std::runtime_error("Expected " + std::to_string(e) + " bytes, but write " + std::to_string(a) + "bytes."), expected(e), actual(a) {}
};
struct InvalidDataException : std::runtime_error { InvalidDataException() : std::runtime_error("InvalidDataException") {} };
struct NotSupportedException : std::runtime_error { explicit NotSupportedException(const char* m = "NotSupportedException") : std::runtime_error(m) {} };
struct ArgumentException : std::invalid_argument { ArgumentException() : std::invalid_argument("ArgumentException") {} };
struct CudaException : std::runtime_error { explicit CudaException(const std::string& m) : std::runtime_error(m) {} };

enum class Endian { Little = AURORA_ENDIAN_LITTLE, Big = AURORA_ENDIAN_BIG };

// CompressionSettings.cs:38-50 (default(CompressionSettings) is quality 8)
struct CompressionSettings {
    int Quality = 8, MaxWindowBits = 0, Strategy = 0;
    CompressionSettings() = default;
    CompressionSettings(int quality, int maxWindowBits = 0, int strategy = 0) : Quality(quality), MaxWindowBits(maxWindowBits), Strategy(strategy) {
        if (quality < 0 || quality > 15 || (maxWindowBits != 0 && (maxWindowBits < 7 || maxWindowBits > 28))) throw ArgumentException();
    }
    static CompressionSettings Fastest() { return {0}; }
    static CompressionSettings Fast() { return {4}; }
    static CompressionSettings Balanced() { return {8}; }
    static CompressionSettings High() { return {12}; }
    static CompressionSettings Maximum() { return {15}; }
};

// process-wide context (all visible B200s); throws when there is none: no CPU fallback
inline aurora_ctx* Context() {
    struct Holder {
        aurora_ctx* ctx;
        Holder() : ctx(aurora_init(0)) {}
        ~Holder() { if (ctx) aurora_shutdown(ctx); }
    };
    static Holder h;
    if (!h.ctx) throw CudaException(std::string("aurora_init failed: ") + aurora_last_error_string(nullptr) + " (no CPU fallback)");
    return h.ctx;
}

inline void ThrowFor(int status, long long expected = 0, long long actual = 0) {
    switch (status) {
        case AURORA_OK: return;
        case AURORA_END_OF_STREAM: throw EndOfStreamException();
        case AURORA_INVALID_IDENTIFIER: throw InvalidIdentifierException();
        case AURORA_SIZE_MISMATCH: throw DecompressedSizeException(expected, actual);
        case AURORA_DST_TOO_SMALL: throw NotSupportedException("destination stream is not expandable");
        case AURORA_INVALID_DATA: throw InvalidDataException();
        case AURORA_NOT_SUPPORTED: throw NotSupportedException();
        case AURORA_INVALID_ARGUMENT: throw ArgumentException();
        default: throw CudaException(aurora_last_error_string(Context()));
    }
}

class CompressionAlgorithm {
  public:
    virtual ~CompressionAlgorithm() = default;
    virtual int Format() const = 0;
    virtual const char* Name() const = 0;

    // IFormatInfoProvider.IsMatch(Stream): a peek, the stream position is restored
    bool IsMatch(std::istream& stream) const {
        const std::vector<uint8_t> data = Remaining(stream, true);
        aurora_codec_opts o = Options(nullptr);
        const uint64_t off = 0, len = data.size();
        uint8_t m = 0;
        const uint8_t dummy = 0;
        Check(aurora_is_match_batch(Context(), Format(), &o, 1, data.empty() ? &dummy : data.data(), &off, &len, &m));
        return m != 0;
    }

    // ICompressionDecoder.Decompress(Stream source, Stream destination): source consumed from its position and left
    // just past the compressed bytes, decoded bytes appended to destination
    void Decompress(std::istream& source, std::ostream& destination) const {
        const std::streampos start = source.tellg();
        const std::vector<uint8_t> data = Remaining(source, false);
        aurora_codec_opts o = Options(nullptr);
        const uint8_t dummy = 0;
        const uint8_t* src = data.empty() ? &dummy : data.data();
        const uint64_t off = 0, len = data.size();
        uint64_t size = 0, out_len = 0, consumed = 0, doff = 0;
        int32_t st = 0;
        Check(aurora_decoded_size_batch(Context(), Format(), &o, 1, src, &off, &len, 1, &size, &st));
        uint64_t cap = st == AURORA_OK ? Capacity(size, len) : 0;
        std::vector<uint8_t> out(std::max<uint64_t>(cap, 1));
        Check(aurora_decode_batch(Context(), Format(), &o, 1, src, &off, &len, out.data(), &doff, &cap, &out_len, &consumed, &st));
        source.clear();
        source.seekg(start + std::streamoff(consumed));
        destination.write(reinterpret_cast<const char*>(out.data()), std::streamsize(std::min(out_len, cap)));
        ThrowFor(st, (long long)size, (long long)out_len);
    }

    // ICompressionEncoder.Compress(ReadOnlySpan<byte> source, Stream destination, CompressionSettings settings = default)
    void Compress(const uint8_t* source, size_t length, std::ostream& destination, const CompressionSettings& settings = {}) const {
        aurora_codec_opts o = Options(&settings);
        const uint8_t dummy = 0;
        const uint64_t off = 0, len = length, doff = 0;
        uint64_t cap = aurora_encode_bound(EncodeFormat(), len), out_len = 0;
        std::vector<uint8_t> out(cap);
        int32_t st = 0;
        Check(aurora_encode_batch(Context(), EncodeFormat(), &o, 1, length ? source : &dummy, &off, &len, out.data(), &doff, &cap, &out_len, &st));
        ThrowFor(st);
        destination.write(reinterpret_cast<const char*>(out.data()), std::streamsize(out_len));
    }

  protected:
    virtual int EncodeFormat() const { return Format(); }
    virtual uint64_t Capacity(uint64_t size, uint64_t /*srcLen*/) const { return size; }
    virtual void Fill(aurora_codec_opts& /*o*/, bool /*encoding*/) const {}

    aurora_codec_opts Options(const CompressionSettings* s) const {
        aurora_codec_opts o;
        aurora_codec_opts_init(&o);
        if (s) {
            o.quality = s->Quality;
            o.max_window_bits = s->MaxWindowBits;
            o.strategy = s->Strategy;
        }
        Fill(o, s != nullptr);
        return o;
    }
    static void Check(int rc) {
        if (rc == AURORA_OK) return;
        if (rc == AURORA_NOT_SUPPORTED) throw NotSupportedException(aurora_last_error_string(Context()));
        if (rc == AURORA_INVALID_ARGUMENT) throw ArgumentException();
        throw CudaException(aurora_last_error_string(Context()));
    }
    static std::vector<uint8_t> Remaining(std::istream& s, bool restore) {
        const std::streampos pos = s.tellg();
        std::vector<uint8_t> v((std::istreambuf_iterator<char>(s)), std::istreambuf_iterator<char>());
        s.clear();
        if (restore) s.seekg(pos);
        return v;
    }
};

// IProvidesDecompressedSize.GetDecompressedSize(Stream): a peek
class SizedAlgorithm : public CompressionAlgorithm {
  public:
    uint32_t GetDecompressedSize(std::istream& source) const {
        std::vector<uint8_t> data = Remaining(source, true);
        data.resize(std::min<size_t>(data.size(), PeekBytes()));
        aurora_codec_opts o = Options(nullptr);
        const uint8_t dummy = 0;
        const uint64_t off = 0, len = data.size();
        uint64_t size = 0;
        int32_t st = 0;
        Check(aurora_decoded_size_batch(Context(), Format(), &o, 1, data.empty() ? &dummy : data.data(), &off, &len, 0, &size, &st));
        ThrowFor(st);
        return uint32_t(size);
    }

  protected:
    virtual size_t PeekBytes() const { return 16; }   // header bytes GetDecompressedSize looks at
};

#define AURORA_FORMAT(cls, fmt, name)             \
    int Format() const override { return fmt; }   \
    const char* Name() const override { return name; }

// Nintendo/Yaz0.cs
class Yaz0 : public SizedAlgorithm {
  public:
    AURORA_FORMAT(Yaz0, AURORA_FMT_YAZ0, "Nintendo Yaz0")
    Endian FormatByteOrder = Endian::Big;
    uint32_t MemoryAlignment = 0;

  protected:
    void Fill(aurora_codec_opts& o, bool) const override {
        o.byte_order = int(FormatByteOrder);
        o.yaz0_alignment = MemoryAlignment;
    }
    uint64_t Capacity(uint64_t size, uint64_t srcLen) const override {   // room for the swapped-size retry (Yaz0.cs:67-78)
        const uint64_t swapped = __builtin_bswap32(uint32_t(size));
        const uint64_t limit = 64 * std::max<uint64_t>(srcLen, 1) + 4096;
        uint64_t cap = size <= limit ? size : 0;
        if (swapped <= limit) cap = std::max(cap, swapped);
        return cap ? cap : size;
    }
};
class Yaz1 : public Yaz0 {
  public:
    AURORA_FORMAT(Yaz1, AURORA_FMT_YAZ1, "Nintendo Yaz1")
};
// Nintendo/Yay0.cs, MIO0.cs: decode detects the order, FormatByteOrder drives Compress
class Yay0 : public SizedAlgorithm {
  public:
    AURORA_FORMAT(Yay0, AURORA_FMT_YAY0, "Nintendo Yay0")
    Endian FormatByteOrder = Endian::Big;

  protected:
    void Fill(aurora_codec_opts& o, bool encoding) const override { o.byte_order = encoding ? int(FormatByteOrder) : AURORA_ENDIAN_DEFAULT; }
};
class MIO0 : public Yay0 {
  public:
    AURORA_FORMAT(MIO0, AURORA_FMT_MIO0, "Nintendo MIO0")
};
// Nintendo/LZ10.cs (GbaVramCompatibilityMode defaults to true, :33), LZ11.cs (false, :29)
class LZ10 : public SizedAlgorithm {
  public:
    AURORA_FORMAT(LZ10, AURORA_FMT_LZ10, "Nintendo LZ10")
    bool GbaVramCompatibilityMode = true;

  protected:
    void Fill(aurora_codec_opts& o, bool) const override { o.vram_mode = GbaVramCompatibilityMode ? 1 : 0; }
};
class LZ11 : public SizedAlgorithm {
  public:
    AURORA_FORMAT(LZ11, AURORA_FMT_LZ11, "Nintendo LZ11")
    bool GbaVramCompatibilityMode = false;

  protected:
    void Fill(aurora_codec_opts& o, bool) const override { o.vram_mode = GbaVramCompatibilityMode ? 1 : 0; }
};
// Formats/Common/LZSS.cs: LZSS(LzProperties), default ((byte)12, 4, 2)
class LZSS : public SizedAlgorithm {
  public:
    AURORA_FORMAT(LZSS, AURORA_FMT_LZSS, "Lempel-Ziv-Storer-Szymanski")
    aurora_lz_props LZ;
    LZSS() { aurora_lz_props_bits(&LZ, 12, 4, 2); }
    explicit LZSS(const aurora_lz_props& lz) : LZ(lz) {}

  protected:
    void Fill(aurora_codec_opts& o, bool) const override { o.lzss = LZ; }
};
// Formats/Common/LZ4.cs, LZ4Legacy.cs, LZO.cs, Snappy.cs; Sega/PRS.cs
class LZ4 : public CompressionAlgorithm {
  public:
    AURORA_FORMAT(LZ4, AURORA_FMT_LZ4, "LZ4 Frame Compression")
};
class LZ4Legacy : public CompressionAlgorithm {
  public:
    AURORA_FORMAT(LZ4Legacy, AURORA_FMT_LZ4_LEGACY, "LZ4 Legacy Compression")
};
class LZO : public CompressionAlgorithm {
  public:
    AURORA_FORMAT(LZO, AURORA_FMT_LZO, "Lempel-Ziv-Oberhumer")
};
class Snappy : public CompressionAlgorithm {
  public:
    AURORA_FORMAT(Snappy, AURORA_FMT_SNAPPY, "Snappy Frame")
};
class PRS : public CompressionAlgorithm {
  public:
    AURORA_FORMAT(PRS, AURORA_FMT_PRS, "SEGA PRS")
    Endian FormatByteOrder = Endian::Big;

  protected:
    void Fill(aurora_codec_opts& o, bool) const override { o.byte_order = int(FormatByteOrder); }
};
// Wrapper formats of AuroraLib.Compression.Nintendo: a header around one of the cores above (aurora_cuda.h, formats 15-22)
class GCLZ : public SizedAlgorithm {   // Nintendo/GCLZ.cs
  public:
    AURORA_FORMAT(GCLZ, AURORA_FMT_GCLZ, "GCLZ")
};
class CXLZ : public SizedAlgorithm {   // Sega/CXLZ.cs
  public:
    AURORA_FORMAT(CXLZ, AURORA_FMT_CXLZ, "CXLZ")
};
class COMP : public SizedAlgorithm {   // Sega/COMP.cs
  public:
    AURORA_FORMAT(COMP, AURORA_FMT_COMP, "COMP")
};
class LZ_3DS : public SizedAlgorithm {   // Nintendo/3DS-LZ.cs
  public:
    AURORA_FORMAT(LZ_3DS, AURORA_FMT_LZ_3DS, "3DS-LZ")
};
class LZ77 : public SizedAlgorithm {   // Nintendo/LZ77.cs: Type (:30) and ChunkSize (:35)
  public:
    AURORA_FORMAT(LZ77, AURORA_FMT_LZ77, "Nintendo LZ77")
    enum CompressionType : uint32_t { LZ10_ = 0x10, LZ11_ = 0x11, ChunkLZ10 = 0xF7 };
    uint32_t Type = LZ10_;
    uint32_t ChunkSize = 0x1000;

  protected:
    void Fill(aurora_codec_opts& o, bool) const override {
        o.lz77_type = Type;
        o.lz77_chunk_size = ChunkSize;
    }
};
class Level5 : public SizedAlgorithm {   // Level5/Level5.cs: OnlySave / LZ10 (the Huffman, RLE and zlib types are not LZ codecs)
  public:
    AURORA_FORMAT(Level5, AURORA_FMT_LEVEL5, "Level5 compression")
    enum CompressionType : uint32_t { OnlySave = 0, LZ10_ = 1 };
    uint32_t Type = LZ10_;

  protected:
    void Fill(aurora_codec_opts& o, bool) const override { o.level5_type = Type; }
};
class LZOn : public SizedAlgorithm {   // Nintendo/LZOn.cs
  public:
    AURORA_FORMAT(LZOn, AURORA_FMT_LZON, "LZOn")
};
class Level5LZSS : public SizedAlgorithm {   // Level5/Level5LZSS.cs
  public:
    AURORA_FORMAT(Level5LZSS, AURORA_FMT_LEVEL5_LZSS, "Level5 lzss")
};
// The LZSS-property family: a fixed header + LZSS.DecompressHeaderless (aurora_cuda.h, formats 23-29)
class AKLZ : public SizedAlgorithm {   // Sega/AKLZ.cs
  public:
    AURORA_FORMAT(AKLZ, AURORA_FMT_AKLZ, "AKLZ")
};
class LZ01 : public SizedAlgorithm {   // Sega/LZ01.cs
  public:
    AURORA_FORMAT(LZ01, AURORA_FMT_LZ01, "LZ01")
};
class FCMP : public SizedAlgorithm {   // Extended/Marvelous/FCMP.cs
  public:
    AURORA_FORMAT(FCMP, AURORA_FMT_FCMP, "FCMP")
};
class IECP : public SizedAlgorithm {   // Extended/Marvelous/IECP.cs
  public:
    AURORA_FORMAT(IECP, AURORA_FMT_IECP, "IECP")
};
class MDB4 : public SizedAlgorithm {   // Extended/Specialized/MDB4.cs
  public:
    AURORA_FORMAT(MDB4, AURORA_FMT_MDB4, "MDB4")
};
class LZSega : public SizedAlgorithm {   // Sega/LZSega.cs (no identifier: IsMatch is not provided)
  public:
    AURORA_FORMAT(LZSega, AURORA_FMT_LZSEGA, "LZSega")
};
class GCZ : public SizedAlgorithm {   // Extended/Konami/GCZ.cs (no identifier: IsMatch is not provided)
  public:
    AURORA_FORMAT(GCZ, AURORA_FMT_GCZ, "Konami GCZ")
};
class SDPC : public SizedAlgorithm {   // Extended/Specialized/SDPC.cs: "SDPC" + size + LZO
  public:
    AURORA_FORMAT(SDPC, AURORA_FMT_SDPC, "SDPC")
};
class LZHudson : public SizedAlgorithm {   // HudsonSoft/LZHudson.cs: u32 BE size + Yay0 tokens under 4-byte big-endian flag words
  public:
    AURORA_FORMAT(LZHudson, AURORA_FMT_LZHUDSON, "LZHudson")
};
class BLZ : public SizedAlgorithm {   // Nintendo/BLZ.cs: backwards from the footer at the end of the stream
  public:
    AURORA_FORMAT(BLZ, AURORA_FMT_BLZ, "Nintendo BLZ")

  protected:
    size_t PeekBytes() const override { return SIZE_MAX; }   // the footer sits at Length - 8
};
class SMSR00 : public SizedAlgorithm {   // Nintendo/SMSR00.cs: MIO0 tokens, 16-bit masks interleaved with the codes
  public:
    AURORA_FORMAT(SMSR00, AURORA_FMT_SMSR00, "Nintendo SMSR00")
};
class LZ40 : public SizedAlgorithm {   // Nintendo/LZ40.cs: negated flag bytes, little-endian 2 / 3 / 4-byte match tokens
  public:
    AURORA_FORMAT(LZ40, AURORA_FMT_LZ40, "Nintendo LZ40")
    bool GbaVramCompatibilityMode = false;

  protected:
    void Fill(aurora_codec_opts& o, bool) const override { o.vram_mode = GbaVramCompatibilityMode ? 1 : 0; }
};
class LZ60 : public SizedAlgorithm {   // Nintendo/LZ60.cs: the LZ40 codec under identifier 0x60
  public:
    AURORA_FORMAT(LZ60, AURORA_FMT_LZ60, "Nintendo LZ60")
    bool GbaVramCompatibilityMode = false;

  protected:
    void Fill(aurora_codec_opts& o, bool) const override { o.vram_mode = GbaVramCompatibilityMode ? 1 : 0; }
};
class ECD : public SizedAlgorithm {   // Extended/Specialized/ECD.cs: plain bytes + LZSS(0x400, 0x42, 3, 0x3BE), or stored
  public:
    AURORA_FORMAT(ECD, AURORA_FMT_ECD, "ECD lzss")
    uint8_t PlainSize = 4;   // ECD.cs:33

  protected:
    void Fill(aurora_codec_opts& o, bool) const override { o.ecd_plain_size = PlainSize; }
    size_t PeekBytes() const override { return SIZE_MAX; }   // GetDecompressedSize compares the compressed size with the stream length
};
class LZ00 : public SizedAlgorithm {   // Sega/LZ00.cs: 64-byte header + LZSS (Lzss0) under a per-byte LCG keystream
  public:
    AURORA_FORMAT(LZ00, AURORA_FMT_LZ00, "LZ00")
    uint32_t Key = 0;   // Compress(source, destination, key, settings); the reference's keyless overload takes the Unix time

  protected:
    void Fill(aurora_codec_opts& o, bool) const override { o.lz00_key = Key; }
    size_t PeekBytes() const override { return 64; }
};
#undef AURORA_FORMAT

// The new batch entry point: many independent blobs at once, sharded over all GPUs of the box.
struct BatchResult {
    std::vector<std::vector<uint8_t>> outputs;
    std::vector<uint64_t> out_len, consumed;
    std::vector<int32_t> status;
};
inline BatchResult DecompressBatch(int format, const std::vector<std::vector<uint8_t>>& sources, const std::vector<uint64_t>& capacities,
                                   const aurora_codec_opts* opts = nullptr) {
    const size_t n = sources.size();
    std::vector<uint64_t> soff(n), slen(n), doff(n), dcap(capacities);
    uint64_t st = 0, dt = 0;
    for (size_t i = 0; i < n; i++) {
        soff[i] = st;
        slen[i] = sources[i].size();
        st += (slen[i] + 15) & ~15ull;
        doff[i] = dt;
        dt += (dcap[i] + 15) & ~15ull;
    }
    uint8_t* ps = static_cast<uint8_t*>(aurora_pinned_alloc(st + 16));
    uint8_t* pd = static_cast<uint8_t*>(aurora_pinned_alloc(dt + 16));
    if (!ps || !pd) throw CudaException("aurora_pinned_alloc failed");
    for (size_t i = 0; i < n; i++) std::copy(sources[i].begin(), sources[i].end(), ps + soff[i]);
    BatchResult r;
    r.out_len.resize(n);
    r.consumed.resize(n);
    r.status.resize(n);
    const int rc = aurora_decode_batch(Context(), format, opts, n, ps, soff.data(), slen.data(), pd, doff.data(), dcap.data(), r.out_len.data(),
                                       r.consumed.data(), r.status.data());
    if (rc == AURORA_OK) {
        r.outputs.resize(n);
        for (size_t i = 0; i < n; i++) r.outputs[i].assign(pd + doff[i], pd + doff[i] + std::min(r.out_len[i], dcap[i]));
    }
    aurora_pinned_free(ps);
    aurora_pinned_free(pd);
    if (rc != AURORA_OK) throw CudaException(aurora_last_error_string(Context()));
    return r;
}

}  // namespace aurora
