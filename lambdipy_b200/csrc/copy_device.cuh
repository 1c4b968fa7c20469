// copy_device.cuh -- warp-level byte movers shared by the compaction kernels.
#pragma once
#include "lb2_common.cuh"

namespace lb2 {

__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(uint4 *p, const uint4 &v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// dst 16-aligned vector <- 16 source bytes starting `sh` bytes (1..15) into `lo`, continuing in `hi`
__device__ __forceinline__ uint4 funnel16(const uint4 &lo, const uint4 &hi, uint32_t sh) {
  const uint32_t w = sh >> 2, b = (sh & 3) * 8;
  uint32_t a0, a1, a2, a3, a4;
  switch (w) {
    case 0: a0 = lo.x; a1 = lo.y; a2 = lo.z; a3 = lo.w; a4 = hi.x; break;
    case 1: a0 = lo.y; a1 = lo.z; a2 = lo.w; a3 = hi.x; a4 = hi.y; break;
    case 2: a0 = lo.z; a1 = lo.w; a2 = hi.x; a3 = hi.y; a4 = hi.z; break;
    default: a0 = lo.w; a1 = hi.x; a2 = hi.y; a3 = hi.z; a4 = hi.w; break;
  }
  uint4 r;
  r.x = __funnelshift_r(a0, a1, b);
  r.y = __funnelshift_r(a1, a2, b);
  r.z = __funnelshift_r(a2, a3, b);
  r.w = __funnelshift_r(a3, a4, b);
  return r;
}

constexpr int UNROLL = 8;  // 16-byte loads in flight per lane

__device__ __forceinline__ void warp_copy_tile(const uint8_t *src, uint8_t *dst, uint32_t len, int lane) {
  // head: bytes up to the first 16-aligned destination address
  uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
  if (head > len) head = len;
  if (lane < (int)head) dst[lane] = __ldg(src + lane);
  src += head; dst += head; len -= head;
  const uint32_t nvec = len >> 4;
  uint4 *dv = reinterpret_cast<uint4 *>(dst);
  const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15);
  if (sh == 0) {
    const uint4 *sv = reinterpret_cast<const uint4 *>(src);
    uint32_t i = lane;
    for (; i + 32 * (UNROLL - 1) < nvec; i += 32 * UNROLL) {
      uint4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; u++) v[u] = ld_stream(sv + i + 32 * u);
#pragma unroll
      for (int u = 0; u < UNROLL; u++) st_stream(dv + i + 32 * u, v[u]);
    }
    for (; i < nvec; i += 32) st_stream(dv + i, ld_stream(sv + i));
  } else {
    // every aligned 16-byte word that holds at least one source byte is inside the allocation
    const uint4 *sv = reinterpret_cast<const uint4 *>(src - sh);
    for (uint32_t i = lane; i < nvec; i += 32) {
      uint4 lo = __ldg(sv + i), hi = __ldg(sv + i + 1);
      st_stream(dv + i, funnel16(lo, hi, sh));
    }
  }
  const uint32_t tail = len & 15, done = nvec << 4;
  if (lane < (int)tail) dst[done + lane] = __ldg(src + done + lane);
}

__device__ __forceinline__ void warp_zero_tile(uint8_t *dst, uint32_t len, int lane) {
  uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
  if (head > len) head = len;
  if (lane < (int)head) dst[lane] = 0;
  dst += head; len -= head;
  const uint32_t nvec = len >> 4;
  uint4 *dv = reinterpret_cast<uint4 *>(dst);
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (uint32_t i = lane; i < nvec; i += 32) st_stream(dv + i, z);
  const uint32_t tail = len & 15, done = nvec << 4;
  if (lane < (int)tail) dst[done + lane] = 0;
}


struct TileView {
  const uint8_t *src;  // nullptr = zero fill
  uint8_t *dst;
  uint32_t len;
};
__device__ __forceinline__ TileView load_tile(const CompactArgs &a, unsigned long long t) {
  const uint4 *tp = reinterpret_cast<const uint4 *>(a.tiles + t);
  const uint4 q0 = __ldg(tp), q1 = __ldg(tp + 1);
  TileView v;
  uint64_t src = (uint64_t)q0.x | ((uint64_t)q0.y << 32);
  if (src - a.rebase_lo < a.rebase_len) src += a.rebase_delta;
  v.src = reinterpret_cast<const uint8_t *>(src);
  v.dst = a.out + __ldg(a.out_off + q1.y) + ((uint64_t)q0.z | ((uint64_t)q0.w << 32));
  v.len = q1.x;
  return v;
}

}  // namespace lb2
