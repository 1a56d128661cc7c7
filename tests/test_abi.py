"""CPU tests of the boundary: libaurora_cuda.so loads, exports every symbol include/aurora_cuda.h declares, the
blittable structs agree between C and ctypes, and the engine refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from auroralib.compression_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    from auroralib.compression_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    return _lib.load()


def _declared_functions():
    hdr = open(os.path.join(ROOT, "include", "aurora_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(aurora_[a-z0-9_]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported(lib):
    from auroralib.compression_b200 import _lib
    declared = _declared_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"libaurora_cuda.so does not export {name}"
    assert sorted(_lib.EXPORTS) == declared


def test_oracle_exports_the_same_batch_shapes(oracle):
    L = oracle.lib()
    for name in ("ora_decode_batch", "ora_encode_batch", "ora_decoded_size_batch", "ora_is_match_batch"):
        assert hasattr(L, name)


def test_abi_version_and_struct_layout(lib):
    assert lib.aurora_abi_version() == A.ABI_VERSION
    o = A.CodecOpts()
    lib.aurora_codec_opts_init(C.byref(o))
    assert o.struct_size == C.sizeof(A.CodecOpts) and o.byte_order == A.ENDIAN_DEFAULT and o.quality == -1 and o.vram_mode == -1
    p = A.LzProps()
    lib.aurora_lz_props_bits(C.byref(p), 10, 6, 2)
    q = A.lz_props_bits(10, 6, 2)
    assert [getattr(p, f[0]) for f in A.LzProps._fields_] == [getattr(q, f[0]) for f in A.LzProps._fields_]
    assert (p.min_length, p.max_length, p.max_distance, p.windows_start) == (3, 66, 1024, 958)
    lib.aurora_lz_props_window(C.byref(p), 0x1000, 0xff + 0x12, 3, 0, 1)
    q = A.lz_props_window(0x1000, 0xff + 0x12, 3, 0, 1)
    assert [getattr(p, f[0]) for f in A.LzProps._fields_] == [getattr(q, f[0]) for f in A.LzProps._fields_]
    assert (p.windows_bits, p.max_length) == (12, 273)


def test_status_strings_and_bounds(lib):
    assert [lib.aurora_status_string(i).decode() for i in range(9)] == A.STATUS_NAMES
    for fmt in range(1, 15):
        for n in (0, 1, 100, 65536, 1 << 20):
            assert lib.aurora_encode_bound(fmt, n) >= n + n // 8 or fmt in (A.FMT_LZ4, A.FMT_LZ4_BLOCK, A.FMT_LZ4_LEGACY, A.FMT_LZO, A.FMT_SNAPPY, A.FMT_SNAPPY_BLOCK)
            assert lib.aurora_encode_bound(fmt, n) >= n


def test_no_cpu_fallback(lib):
    """Without a CUDA device the engine must fail loudly, not decode on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    from auroralib.compression_b200 import AuroraError, BatchCodec
    assert lib.aurora_device_count() == 0
    with pytest.raises(AuroraError):
        BatchCodec()
    from auroralib.compression_b200 import LZ10
    import io
    with pytest.raises(AuroraError):
        LZ10().Decompress(io.BytesIO(b"\x10\x04\x00\x00\x00abcd"), io.BytesIO())


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "auroralib", "compression_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("the oracle", "").replace("oracle encoder", "").replace("oracle decoder", "").replace("CPU oracle", "").replace("oracle's", "") \
                    or f in ("corpus.py",), f"{f} references oracle/"


def test_strategy_bits_match_the_header():
    """The library-specific bits of opts.strategy (which byte-identical match finder encodes) are the same in the C header and
    in the Python mirror, and leave bit 0 (CompresionStrategy.CompatibilityMode, the only bit the oracle reads) alone."""
    import re
    hdr = open(os.path.join(ROOT, "include", "aurora_cuda.h")).read()
    vals = {m.group(1): int(m.group(2), 16) for m in re.finditer(r"#define (AURORA_STRATEGY_\w+)\s+(0x[0-9A-Fa-f]+)", hdr)}
    assert vals == {"AURORA_STRATEGY_PARALLEL_FINDER": A.STRATEGY_PARALLEL_FINDER, "AURORA_STRATEGY_SERIAL_FINDER": A.STRATEGY_SERIAL_FINDER}
    assert A.STRATEGY_COMPATIBILITY == 1 and (A.STRATEGY_PARALLEL_FINDER | A.STRATEGY_SERIAL_FINDER) & 1 == 0
