// P/Invoke binding of libaurora_cuda.so (include/aurora_cuda.h).  UNTESTED: no .NET toolchain exists in the build
// image; this file is the reference-side stub a maintainer adds (see INTEGRATION.md).
using System;
using System.Runtime.InteropServices;

namespace AuroraLib.Compression.Cuda
{
    public enum AuroraStatus : int
    {
        Ok = 0, EndOfStream = 1, InvalidIdentifier = 2, SizeMismatch = 3, DstTooSmall = 4,
        InvalidData = 5, NotSupported = 6, InvalidArgument = 7, CudaError = 8
    }

    public enum AuroraFormat : int
    {
        Yaz0 = 1, Yaz1 = 2, Yay0 = 3, MIO0 = 4, LZ10 = 5, LZ11 = 6, LZSS = 7, LZ4 = 8, LZ4Block = 9,
        LZ4Legacy = 10, LZO = 11, Snappy = 12, SnappyBlock = 13, PRS = 14,
        // wrapper formats (AuroraLib.Compression.Nintendo): a header around one of the cores above
        GCLZ = 15, CXLZ = 16, COMP = 17, LZ_3DS = 18, LZ77 = 19, Level5 = 20, LZOn = 21, Level5LZSS = 22,
        // the LZSS-property family (AuroraLib.Compression.Sega, AuroraLib.Compression-Extended)
        AKLZ = 23, LZ01 = 24, FCMP = 25, IECP = 26, MDB4 = 27, LZSega = 28, GCZ = 29, SDPC = 30,
        // LZSS wrappers with extra work around the core: stored prefix / stored fallback (ECD), LCG keystream on the device (LZ00)
        ECD = 31, LZ00 = 32,
        // a core format: Yay0 tokens under 32-bit big-endian flag words (HudsonSoft/LZHudson.cs)
        LZHudson = 33,
        // core formats: LZ11-like tokens, little-endian with the length in the low nibble, negated flag bytes
        LZ40 = 34, LZ60 = 35,
        // core format: MIO0 tokens, 16-bit big-endian mask words interleaved with the codes, literals in their own section
        SMSR00 = 36,
        // core format with its own kernel: parsed and written backwards from the footer at the end of the stream
        BLZ = 37
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct AuroraLzProps
    {
        public int WindowsBits, LengthBits, MinLength, MaxLength, MaxDistance, MinDistance, WindowsStart, Reserved;
    }

    [StructLayout(LayoutKind.Sequential)]
    public unsafe struct AuroraCodecOpts
    {
        public uint StructSize;
        public int ByteOrder;          // 0 little, 1 big, 2 class default
        public int Quality;            // -1 = default(CompressionSettings)
        public int MaxWindowBits;
        public int Strategy;
        public int VramMode;           // -1 class default
        public AuroraLzProps Lzss;
        public int LzssInitialFill;
        public uint Lz4BlockSize;
        public int Lz4Verify;
        public uint Yaz0Alignment;
        public uint Balance;           // 0 auto, 1 largest-first on, 2 off
        public uint Lz77Type;          // LZ77.Type: 0 -> 0x10; 0x11; 0xF7 (ChunkLZ10)
        public uint Lz77ChunkSize;     // LZ77.ChunkSize: 0 -> 0x1000
        public uint Level5Type;        // Level5.Type: 0 -> 1 (LZ10)
        public uint Lz00Key;           // LZ00.Compress(source, destination, key, settings)
        public uint EcdPlainSize;      // ECD.PlainSize: 0 -> 4
    }

    internal static unsafe class Native
    {
        private const string Lib = "aurora_cuda";   // libaurora_cuda.so

        [DllImport(Lib)] public static extern IntPtr aurora_init(uint deviceMask);
        [DllImport(Lib)] public static extern void aurora_shutdown(IntPtr ctx);
        [DllImport(Lib)] public static extern int aurora_device_count();
        [DllImport(Lib)] public static extern int aurora_ctx_device_count(IntPtr ctx);
        [DllImport(Lib)] public static extern int aurora_abi_version();
        [DllImport(Lib)] public static extern IntPtr aurora_last_error_string(IntPtr ctx);
        [DllImport(Lib)] public static extern IntPtr aurora_pinned_alloc(UIntPtr bytes);
        [DllImport(Lib)] public static extern void aurora_pinned_free(IntPtr p);
        [DllImport(Lib)] public static extern void aurora_codec_opts_init(AuroraCodecOpts* opts);
        [DllImport(Lib)] public static extern void aurora_lz_props_window(AuroraLzProps* o, int windowsSize, int maxLength, int minLength, int windowsStart, int minDistance);
        [DllImport(Lib)] public static extern void aurora_lz_props_bits(AuroraLzProps* o, int distanceBits, int lengthBits, int threshold);
        [DllImport(Lib)] public static extern ulong aurora_encode_bound(int format, ulong rawLen);

        [DllImport(Lib)]
        public static extern int aurora_decoded_size_batch(IntPtr ctx, int format, AuroraCodecOpts* opts, UIntPtr n, byte* srcBase,
            ulong* srcOff, ulong* srcLen, int sizeScan, ulong* outSize, int* status);

        [DllImport(Lib)]
        public static extern int aurora_is_match_batch(IntPtr ctx, int format, AuroraCodecOpts* opts, UIntPtr n, byte* srcBase,
            ulong* srcOff, ulong* srcLen, byte* match);

        [DllImport(Lib)]
        public static extern int aurora_decode_batch(IntPtr ctx, int format, AuroraCodecOpts* opts, UIntPtr n, byte* srcBase,
            ulong* srcOff, ulong* srcLen, byte* dstBase, ulong* dstOff, ulong* dstCap, ulong* outLen, ulong* consumed, int* status);

        [DllImport(Lib)]
        public static extern int aurora_encode_batch(IntPtr ctx, int format, AuroraCodecOpts* opts, UIntPtr n, byte* srcBase,
            ulong* srcOff, ulong* srcLen, byte* dstBase, ulong* dstOff, ulong* dstCap, ulong* outLen, int* status);
    }
}
