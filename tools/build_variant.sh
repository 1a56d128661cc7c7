#!/bin/bash
# Developer tool: build a variant of libaurora_cuda.so with extra -D flags for A/B probes on the GPU box.
#   tools/build_variant.sh NAME "-DAURORA_FLAG_WARPS=4 ..."   ->  auroralib/compression_b200/variants/libaurora_cuda_NAME.so
# Load it with AURORA_CUDA_LIB=<path> (auroralib/compression_b200/_lib.py).
set -e
NAME=$1; shift
FLAGS="$*"
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/auroralib/compression_b200/csrc
OUT=$ROOT/auroralib/compression_b200/variants
B=$SRC/build/variant_$NAME
mkdir -p "$OUT" "$B"
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in api wrappers decode_flaglz decode_bytelz decode_blz encode_lz encode_lz_par encode_bytelz ismatch keystream; do
  if [ "$f" = decode_flaglz ] || [ "$f" = decode_bytelz ] || [ "$f" = decode_blz ] || [ "$f" = encode_lz_par ] || [ ! -f "$SRC/build/$f.o" ]; then
    nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -cudart static $FLAGS -c "$SRC/$f.cu" -o "$B/$f.o" &
  else
    cp "$SRC/build/$f.o" "$B/$f.o"
  fi
done
wait
nvcc $ARCH -shared -cudart static -o "$OUT/libaurora_cuda_$NAME.so" "$B"/*.o -lpthread
echo "$OUT/libaurora_cuda_$NAME.so"
