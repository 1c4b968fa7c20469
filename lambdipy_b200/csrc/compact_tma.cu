// compact_tma.cu -- compaction through the bulk-copy (TMA) engine: cp.async.bulk global->shared
// signalled on an mbarrier, then cp.async.bulk shared->global, 16 KB stages in a shared-memory
// ring.  No registers and no LSU instructions touch the payload; one producer lane and one storer
// lane per CTA drive the engine, the remaining warps take what the engine cannot: heads/tails of
// < 16 bytes, tiles whose source and destination are not congruent mod 16, zero fills.
//
// One persistent CTA per SM (grid = SM count).  SASS: UBLKCP (see profiles/).
#include "lb2_common.cuh"
#include "copy_device.cuh"

namespace lb2 {

constexpr int TMA_STAGES = 12;           // 12 x 16 KB = 192 KB of the 227 KB shared memory
constexpr int TMA_THREADS = 256;         // warp 0: producer, warp 1: storer, warps 2..7: helpers
constexpr int TMA_HELPERS = TMA_THREADS / 32 - 2;
constexpr uint32_t TMA_MIN_BODY = 2048;  // smaller aligned bodies go through the LSU path
constexpr int TMA_STORES_IN_FLIGHT = 6;  // bulk stores allowed to be still reading shared memory

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}

// How a tile splits between the engine (16-byte aligned body) and the LSU helpers (head, tail).
struct BulkSplit { uint32_t head, body; };
__device__ __forceinline__ BulkSplit bulk_split(const TileView &v) {
  BulkSplit s{0, 0};
  if (!v.src) return s;
  if (((reinterpret_cast<uintptr_t>(v.src) ^ reinterpret_cast<uintptr_t>(v.dst)) & 15) != 0) return s;
  uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(v.dst) & 15)) & 15);
  if (head >= v.len) return s;
  uint32_t body = (v.len - head) & ~15u;
  if (body < TMA_MIN_BODY) return s;
  s.head = head; s.body = body;
  return s;
}

__global__ void __launch_bounds__(TMA_THREADS, 1) lb2_compact_tma_kernel(CompactArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[TMA_STAGES], empty_bar[TMA_STAGES];
  if (a.ctr->overflow) return;
  const unsigned long long n_tiles = a.ctr->n_tiles;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TMA_STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == 0) {
    // ---- producer: one lane queues bulk loads, running up to TMA_STAGES tiles ahead
    if (lane == 0) {
      uint32_t it = 0;
      for (unsigned long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const TileView v = load_tile(a, t);
        const BulkSplit sp = bulk_split(v);
        if (!sp.body) continue;
        const uint32_t s = it % TMA_STAGES, round = it / TMA_STAGES;
        if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1);
        mbar_expect_tx(&full_bar[s], sp.body);
        bulk_g2s(smem + (size_t)s * TILE_BYTES, v.src + sp.head, sp.body, &full_bar[s]);
        it++;
      }
    }
  } else if (warp == 1) {
    // ---- storer: as each stage lands, queue its bulk store; release stages whose store has
    //      finished reading shared memory
    if (lane == 0) {
      uint32_t it = 0;
      for (unsigned long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const TileView v = load_tile(a, t);
        const BulkSplit sp = bulk_split(v);
        if (!sp.body) continue;
        const uint32_t s = it % TMA_STAGES, round = it / TMA_STAGES;
        mbar_wait(&full_bar[s], round & 1);
        bulk_s2g(v.dst + sp.head, smem + (size_t)s * TILE_BYTES, sp.body);
        bulk_commit();
        if (it >= (uint32_t)TMA_STORES_IN_FLIGHT) {
          bulk_wait_read<TMA_STORES_IN_FLIGHT>();
          mbar_arrive(&empty_bar[(it - TMA_STORES_IN_FLIGHT) % TMA_STAGES]);
        }
        it++;
      }
      bulk_wait_read<0>();
      // (no one waits on the remaining empty barriers)
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores globally performed before exit
    }
  } else {
    // ---- helpers: heads/tails of engine tiles, and every tile the engine does not take
    const int h = warp - 2;
    unsigned long long j = 0;
    for (unsigned long long t = blockIdx.x; t < n_tiles; t += gridDim.x, j++) {
      if ((int)(j % TMA_HELPERS) != h) continue;
      const TileView v = load_tile(a, t);
      const BulkSplit sp = bulk_split(v);
      if (sp.body) {
        if (lane < (int)sp.head) v.dst[lane] = __ldg(v.src + lane);
        const uint32_t done = sp.head + sp.body, tail = v.len - done;
        if (lane < (int)tail) v.dst[done + lane] = __ldg(v.src + done + lane);
      } else if (v.src) {
        warp_copy_tile(v.src, v.dst, v.len, lane);
      } else {
        warp_zero_tile(v.dst, v.len, lane);
      }
    }
  }
}

void launch_compact_tma(const CompactArgs &a, int grid, cudaStream_t s) {
  static bool configured = false;
  const size_t smem = (size_t)TMA_STAGES * TILE_BYTES;
  if (!configured) {
    cudaFuncSetAttribute(lb2_compact_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = true;
  }
  lb2_compact_tma_kernel<<<grid, TMA_THREADS, smem, s>>>(a);
}

}  // namespace lb2
