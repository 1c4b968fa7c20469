// Minimal stand-in for <cuda_runtime.h> so that lambdipy_b200/csrc/plan.cu compiles with g++ for the
// warp emulator (tests/emu/plan_emu.cpp).  TEST HARNESS ONLY -- never part of the product build.
#pragma once
#include <cstdint>
#include <cstring>
#include <cstdio>
struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct uint3 { uint32_t x, y, z; };
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }
typedef void *cudaStream_t;
#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)
#define __restrict__
