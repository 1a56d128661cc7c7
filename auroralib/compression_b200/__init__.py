"""auroralib.compression_b200 — B200-native batched LZ-family codec engine behind AuroraLib.Compression's
ICompressionAlgorithm surface.  Hot path: hand-written sm_100a kernels in csrc/ behind the C ABI of
include/aurora_cuda.h (libaurora_cuda.so).  No CPU fallback."""
from . import _abi
from ._abi import (FMT_BLZ, FMT_LZ4, FMT_LZ4_BLOCK, FMT_LZ4_LEGACY, FMT_LZ10, FMT_LZ11, FMT_LZ40, FMT_LZ60, FMT_LZHUDSON, FMT_LZO, FMT_LZSS, FMT_MIO0, FMT_PRS,
                   FMT_SMSR00, FMT_SNAPPY, FMT_SNAPPY_BLOCK, FMT_YAY0, FMT_YAZ0, FMT_YAZ1, make_opts)
from .batch import AuroraError, BatchCodec, default_codec, layout, pack
from .codecs import (AKLZ, COMP, CXLZ, ECD, FCMP, GCLZ, GCZ, IECP, LZ00, LZ01, LZ77, LZ_3DS, MDB4, SDPC, WRAPPERS, Level5, Level5LZSS, LZOn,
                     LZSega)
from .codecs import (ALGORITHMS, BLZ, LZ4, LZ10, LZ11, LZO, LZSS, MIO0, PRS, CompressionSettings, DecompressedSizeException,
                     Endian, EndOfStreamException, InvalidDataException, InvalidIdentifierException, LZ4Legacy, LZHudson, LZ40, LZ60, SMSR00,
                     LzProperties, NotSupportedException, Snappy, Yay0, Yaz0, Yaz1)

__all__ = [n for n in dir() if not n.startswith("_")]
