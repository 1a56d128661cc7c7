// C++ host mirror exercise: the reference's EncodingAndDecodingMatchTest / DataRecognitionTest shapes
// (CompressionTest/CompressionAlgorithmTest.cs:60-130) through include/aurora_codecs.hpp.
// usage: mirror_test <Test.bmp> [expect-no-gpu]
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "aurora_codecs.hpp"

using namespace aurora;

template <typename T>
static int round_trip(T&& algo, const std::string& raw, const CompressionSettings& s, bool sized, bool check_match = true) {
    std::stringstream comp, out;
    algo.Compress(reinterpret_cast<const uint8_t*>(raw.data()), raw.size(), comp, s);
    comp.seekg(0);
    if (check_match && !algo.IsMatch(comp)) { std::printf("%s: IsMatch false\n", algo.Name()); return 1; }
    if (comp.tellg() != std::streampos(0)) { std::printf("%s: IsMatch moved the stream\n", algo.Name()); return 1; }
    algo.Decompress(comp, out);
    if (out.str() != raw) { std::printf("%s: round trip differs\n", algo.Name()); return 1; }
    size_t end = comp.str().size();
    if (std::string(algo.Name()) == "Nintendo BLZ") end -= uint8_t(comp.str()[end - 5]);   // BLZ leaves the source in front of its padding + footer
    if (size_t(comp.tellg()) != end) { std::printf("%s: source not left at the end of the compressed bytes\n", algo.Name()); return 1; }
    (void)sized;
    std::printf("%s: %zu -> %zu bytes ok\n", algo.Name(), raw.size(), comp.str().size());
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::ifstream f(argv[1], std::ios::binary);
    std::string bmp((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const bool expect_no_gpu = argc > 2 && std::strcmp(argv[2], "expect-no-gpu") == 0;
    try {
        int bad = 0;
        const std::string raw = bmp.substr(0, 10240);
        bad += round_trip(Yaz0(), raw, CompressionSettings::Balanced(), true);
        bad += round_trip(Yaz1(), raw, CompressionSettings::Maximum(), true);
        bad += round_trip(Yay0(), raw, CompressionSettings::Balanced(), true);
        bad += round_trip(MIO0(), raw, CompressionSettings::Fastest(), true);
        bad += round_trip(LZ10(), bmp.substr(0, 1 << 20), CompressionSettings::Fastest(), true);
        bad += round_trip(LZ11(), raw, CompressionSettings::Balanced(), true);
        bad += round_trip(LZSS(), raw, CompressionSettings::Balanced(), true);
        bad += round_trip(LZ4(), raw, CompressionSettings::Balanced(), false);
        bad += round_trip(LZ4Legacy(), raw, CompressionSettings::Fast(), false);
        bad += round_trip(LZO(), raw, CompressionSettings::Balanced(), false);
        bad += round_trip(Snappy(), raw, CompressionSettings::Balanced(), false);
        bad += round_trip(PRS(), raw, CompressionSettings::Balanced(), false);
        // wrapper formats (headers resolved on the host, cores on the device)
        bad += round_trip(GCLZ(), raw, CompressionSettings::Balanced(), true);
        bad += round_trip(CXLZ(), raw, CompressionSettings::Balanced(), true);
        bad += round_trip(COMP(), raw, CompressionSettings::Balanced(), true);
        bad += round_trip(LZ_3DS(), raw, CompressionSettings::Balanced(), true);
        bad += round_trip(LZ77(), raw, CompressionSettings::Balanced(), true);
        {
            LZ77 chunked;
            chunked.Type = LZ77::ChunkLZ10;
            chunked.ChunkSize = 0x800;
            bad += round_trip(chunked, raw, CompressionSettings::Balanced(), true);
        }
        bad += round_trip(Level5(), raw, CompressionSettings::Balanced(), true, false);   // Level5.IsMatch is not provided (zlib, file name)
        bad += round_trip(LZOn(), raw, CompressionSettings::Balanced(), true);
        bad += round_trip(Level5LZSS(), raw, CompressionSettings::Balanced(), true);
        bad += round_trip(AKLZ(), raw, CompressionSettings::Balanced(), true);
        bad += round_trip(LZ01(), raw, CompressionSettings::Balanced(), true);
        bad += round_trip(FCMP(), raw, CompressionSettings::Balanced(), true);
        bad += round_trip(IECP(), raw, CompressionSettings::Balanced(), true);
        bad += round_trip(MDB4(), raw, CompressionSettings::Balanced(), true);
        bad += round_trip(LZSega(), raw, CompressionSettings::Balanced(), true, false);
        bad += round_trip(GCZ(), raw, CompressionSettings::Balanced(), true, false);
        bad += round_trip(SDPC(), raw, CompressionSettings::Balanced(), true);
        bad += round_trip(LZHudson(), raw, CompressionSettings::Balanced(), true);
        bad += round_trip(SMSR00(), raw, CompressionSettings::Balanced(), true);
        bad += round_trip(BLZ(), raw, CompressionSettings::Balanced(), true);
        bad += round_trip(LZ40(), raw, CompressionSettings::Maximum(), true);
        bad += round_trip(LZ60(), raw, CompressionSettings::Balanced(), true);
        bad += round_trip(ECD(), raw, CompressionSettings::Balanced(), true);
        bad += round_trip(ECD(), raw, CompressionSettings::Fastest(), true);   // quality 0: stored
        {
            LZ00 keyed;
            keyed.Key = 0x5EEDC0DEu;
            bad += round_trip(keyed, raw, CompressionSettings::Balanced(), true);
        }
        // GetDecompressedSize (DataRecognitionTest: 256 zero bytes)
        {
            std::string zeros(0x100, '\0');
            std::stringstream comp;
            LZ10().Compress(reinterpret_cast<const uint8_t*>(zeros.data()), zeros.size(), comp, CompressionSettings::Fastest());
            comp.seekg(0);
            if (LZ10().GetDecompressedSize(comp) != 0x100) { std::printf("GetDecompressedSize wrong\n"); bad++; }
        }
        // exceptions
        {
            std::stringstream comp, out;
            LZ10().Compress(reinterpret_cast<const uint8_t*>(raw.data()), raw.size(), comp);
            std::string c = comp.str();
            std::stringstream trunc(c.substr(0, 100)), wrong("\x11" + c.substr(1));
            try { LZ10().Decompress(trunc, out); bad++; std::printf("no EndOfStreamException\n"); } catch (const EndOfStreamException&) {}
            try { LZ10().Decompress(wrong, out); bad++; std::printf("no InvalidIdentifierException\n"); } catch (const InvalidIdentifierException&) {}
        }
        // batch entry point
        {
            std::vector<std::vector<uint8_t>> src;
            std::vector<uint64_t> caps;
            for (int i = 0; i < 50; i++) {
                std::stringstream comp;
                const std::string r = bmp.substr(size_t(i) * 1000, 3000 + 100 * i);
                Yaz0().Compress(reinterpret_cast<const uint8_t*>(r.data()), r.size(), comp);
                const std::string c = comp.str();
                src.emplace_back(c.begin(), c.end());
                caps.push_back(r.size());
            }
            BatchResult r = DecompressBatch(AURORA_FMT_YAZ0, src, caps);
            for (int i = 0; i < 50; i++)
                if (r.status[i] != AURORA_OK || std::string(r.outputs[i].begin(), r.outputs[i].end()) != bmp.substr(size_t(i) * 1000, 3000 + 100 * i)) bad++;
            std::printf("batch of 50 Yaz0 streams: %s\n", bad ? "BAD" : "ok");
        }
        if (expect_no_gpu) { std::printf("expected a failure without a GPU\n"); return 1; }
        std::printf("mirror_test: %d failures\n", bad);
        return bad ? 1 : 0;
    } catch (const CudaException& e) {
        std::printf("CudaException: %s\n", e.what());
        return expect_no_gpu ? 0 : 3;
    }
}
