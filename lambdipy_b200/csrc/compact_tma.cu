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

  // Producer and storer warps read tile descriptors 32 at a time (one per lane, one round trip to L2
  // for the whole batch) and hand them to their issuing lane by shuffle: a per-tile dependent load
  // in front of every 16 KB bulk copy would cap an SM at one tile per L2 latency.
  if (warp == 0 || warp == 1) {
    uint32_t it = 0;
    for (unsigned long long tb = blockIdx.x; tb < n_tiles; tb += (unsigned long long)gridDim.x * 32) {
      const unsigned long long t = tb + (unsigned long long)lane * gridDim.x;
      uint64_t src = 0, dst = 0;
      uint32_t body = 0;
      if (t < n_tiles) {
        const TileView v = load_tile(a, t);
        const BulkSplit sp = bulk_split(v);
        body = sp.body;
        src = reinterpret_cast<uint64_t>(v.src) + sp.head;
        dst = reinterpret_cast<uint64_t>(v.dst) + sp.head;
      }
      unsigned todo = __ballot_sync(0xffffffffu, body != 0);
      while (todo) {
        const int l = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint64_t s_ = __shfl_sync(0xffffffffu, src, l), d_ = __shfl_sync(0xffffffffu, dst, l);
        const uint32_t b_ = __shfl_sync(0xffffffffu, body, l);
        if (lane == 0) {
          const uint32_t st = it % TMA_STAGES, round = it / TMA_STAGES;
          if (warp == 0) {
            // ---- producer: queue the bulk load, running up to TMA_STAGES tiles ahead of the stores
            if (round > 0) mbar_wait(&empty_bar[st], (round - 1) & 1);
            mbar_expect_tx(&full_bar[st], b_);
            bulk_g2s(smem + (size_t)st * TILE_BYTES, reinterpret_cast<const void *>(s_), b_, &full_bar[st]);
          } else {
            // ---- storer: as the stage lands queue its bulk store; release the stage whose store
            //      has finished reading shared memory
            mbar_wait(&full_bar[st], round & 1);
            bulk_s2g(reinterpret_cast<void *>(d_), smem + (size_t)st * TILE_BYTES, b_);
            bulk_commit();
            if (it >= (uint32_t)TMA_STORES_IN_FLIGHT) {
              bulk_wait_read<TMA_STORES_IN_FLIGHT>();
              mbar_arrive(&empty_bar[(it - TMA_STORES_IN_FLIGHT) % TMA_STAGES]);
            }
          }
        }
        it++;
      }
    }
    if (warp == 1 && lane == 0) {
      bulk_wait_read<0>();
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores performed before the CTA exits
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
