// corpus.cu -- payload generator for the synthetic ELF corpus: fills
// regions of an HBM arena with counter-based pseudo-random bytes so a 100 GB corpus never has to
// exist on the host.  byte(o) = byte (o & 7) of splitmix64(seed + (o >> 3)): a pure function of the
// arena offset, so the host-side generator (lambdipy_b200/corpus.py) reproduces any file exactly.
#include "lb2_common.cuh"

namespace lb2 {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256) lb2_fill_kernel(uint8_t *arena, const FillRegion *regions, uint32_t n, uint64_t seed) {
  for (uint32_t r = blockIdx.x; r < n; r += gridDim.x) {
    const uint64_t off = regions[r].offset, len = regions[r].len;
    uint64_t b = off, e = off + len;
    // byte head up to 16-alignment
    uint64_t b16 = (b + 15) & ~15ull;
    if (b16 > e) b16 = e;
    for (uint64_t o = b + threadIdx.x; o < b16; o += blockDim.x) arena[o] = (uint8_t)(splitmix64(seed + (o >> 3)) >> ((o & 7) * 8));
    const uint64_t e16 = b16 + ((e - b16) & ~15ull);
    for (uint64_t o = b16 + (uint64_t)threadIdx.x * 16; o < e16; o += (uint64_t)blockDim.x * 16) {
      const uint64_t lo = splitmix64(seed + (o >> 3)), hi = splitmix64(seed + (o >> 3) + 1);
      *reinterpret_cast<uint4 *>(arena + o) = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
    }
    for (uint64_t o = e16 + threadIdx.x; o < e; o += blockDim.x) arena[o] = (uint8_t)(splitmix64(seed + (o >> 3)) >> ((o & 7) * 8));
  }
}

void launch_fill(uint8_t *arena, const FillRegion *d_regions, uint32_t n, uint64_t seed, int grid, cudaStream_t s) {
  lb2_fill_kernel<<<grid, 256, 0, s>>>(arena, d_regions, n, seed);
}

}  // namespace lb2
