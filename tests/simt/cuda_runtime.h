// tests/simt/cuda_runtime.h — stand-in for the CUDA runtime header when a kernel source file is compiled for the CPU lane
// emulation (tests/simt/simt.hpp).  Only what csrc/common.cuh's host part and csrc/encode_lz_par.cu's device part mention.
#pragma once
#include <stddef.h>
#include <stdint.h>
typedef int cudaError_t;
typedef void* cudaStream_t;
struct uint4 {
    uint32_t x, y, z, w;
};
struct uint2 {
    uint32_t x, y;
};
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
struct int4 {
    int x, y, z, w;
};
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
#define __constant__
static inline void __threadfence_block() {}
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) {   // one emulated block at a time: no concurrency
    const unsigned old = *p;
    *p += v;
    return old;
}
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
