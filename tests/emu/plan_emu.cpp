// plan_emu.cpp -- TEST HARNESS: runs the PRODUCT's device planner source (lambdipy_b200/csrc/plan.cu) on the
// CPU by emulating one CTA (4 warps) with 128 host threads (barrier-based __syncthreads / __syncwarp / __ballot_sync / __shfl*), then
// executes the emitted tile list with memcpy to materialise the stripped file.  It lets the CPU test-suite
// (-m "not gpu") check the planner logic that ships in the CUDA library against the oracle and GNU strip
// without a GPU.  It is NOT a CPU fallback: it lives under tests/, is built only by the tests, is never
// imported by lambdipy_b200/, and emulates a single file per call at kHz speeds.
#include <atomic>
#include <barrier>
#include <cstdlib>
#include <thread>
#include <vector>

#include "cuda_runtime.h"  // the shim

// ---- CTA emulation state: 4 warps of 32 host threads; warp collectives exchange through a per-warp slot
//      array and barrier, __syncthreads is a barrier over all 128 threads
static constexpr int EMU_THREADS = 128, EMU_WARPS = EMU_THREADS / 32;
static thread_local uint3 threadIdx, blockIdx;
static std::barrier<> *g_wbar[EMU_WARPS];
static std::barrier<> *g_bbar;
static uint64_t g_xchg[EMU_WARPS][32];
#define EMU_W (threadIdx.x >> 5)
#define EMU_L (threadIdx.x & 31)
static inline void __syncwarp() { g_wbar[EMU_W]->arrive_and_wait(); }
static inline void __syncthreads() { g_bbar->arrive_and_wait(); }
static inline unsigned __ballot_sync(unsigned, int pred) {
  g_xchg[EMU_W][EMU_L] = pred ? 1 : 0;
  g_wbar[EMU_W]->arrive_and_wait();
  unsigned m = 0;
  for (int i = 0; i < 32; i++) m |= (unsigned)g_xchg[EMU_W][i] << i;
  g_wbar[EMU_W]->arrive_and_wait();
  return m;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) {
  uint64_t raw = 0; memcpy(&raw, &v, sizeof v);
  g_xchg[EMU_W][EMU_L] = raw;
  g_wbar[EMU_W]->arrive_and_wait();
  T r; memcpy(&r, &g_xchg[EMU_W][src & 31], sizeof r);
  g_wbar[EMU_W]->arrive_and_wait();
  return r;
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) {
  uint64_t raw = 0; memcpy(&raw, &v, sizeof v);
  g_xchg[EMU_W][EMU_L] = raw;
  g_wbar[EMU_W]->arrive_and_wait();
  T r = v;
  if ((int)EMU_L >= d) memcpy(&r, &g_xchg[EMU_W][EMU_L - d], sizeof r);
  g_wbar[EMU_W]->arrive_and_wait();
  return r;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) {
  uint64_t raw = 0; memcpy(&raw, &v, sizeof v);
  g_xchg[EMU_W][EMU_L] = raw;
  g_wbar[EMU_W]->arrive_and_wait();
  T r; memcpy(&r, &g_xchg[EMU_W][(EMU_L ^ m) & 31], sizeof r);
  g_wbar[EMU_W]->arrive_and_wait();
  return r;
}
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline long long clock64() { return 0; }

#define LB2_HOST_EMULATION 1
#include "../../lambdipy_b200/csrc/plan.cu"

using namespace lb2;

static void run_block(const PlanArgs &a) {
  std::barrier<> bbar(EMU_THREADS);
  std::barrier<> w0(32), w1(32), w2(32), w3(32);
  g_bbar = &bbar;
  g_wbar[0] = &w0; g_wbar[1] = &w1; g_wbar[2] = &w2; g_wbar[3] = &w3;
  std::vector<std::thread> th;
  for (int l = 0; l < EMU_THREADS; l++)
    th.emplace_back([&, l] {
      threadIdx = uint3{(uint32_t)l, 0, 0};
      blockIdx = uint3{0, 0, 0};
      lb2_plan_kernel(a);
      // threads leave together (every early exit of the kernel is taken by the whole CTA)
      g_wbar[l >> 5]->arrive_and_drop();
      bbar.arrive_and_drop();
    });
  for (auto &t : th) t.join();
}

extern "C" int lb2emu_strip(const uint8_t *in, uint64_t n, uint32_t flags, uint8_t **out_p, uint64_t *out_n) {
  *out_p = nullptr; *out_n = 0;
  // 256-aligned private copy, as the arena contract requires
  const uint64_t cap = ((n + 255) & ~255ull) + 256;
  uint8_t *arena = static_cast<uint8_t *>(aligned_alloc(256, cap));
  memset(arena, 0, cap);
  memcpy(arena, in, n);
  uint8_t *scratch = static_cast<uint8_t *>(aligned_alloc(256, SCR_STRIDE));
  memset(scratch, 0, SCR_STRIDE);
  const uint64_t tile_cap = n / TILE_BYTES + 2 * MAX_EXT + 16 + 4096;
  std::vector<Tile> tiles(tile_cap);
  uint64_t in_off = 0, in_size = n, out_size = 0;
  int32_t status = -99;
  BatchCounters ctr;
  memset(&ctr, 0, sizeof ctr);
  PlanArgs a;
  a.in = arena; a.in_off = &in_off; a.in_size = &in_size; a.n_files = 1; a.flags = flags;
  a.scratch = scratch; a.out_size = &out_size; a.status = &status; a.tiles = tiles.data(); a.tile_cap = tile_cap; a.ctr = &ctr; a.up_ranges = nullptr; a.up_cap = 0;
  // (LB2EMU_BIG_CAP=0 leaves no room in the list: the planning CTA then writes the tiles of huge extents itself)
  const char *cap_env = getenv("LB2EMU_BIG_CAP");
  std::vector<BigExt> big(256);
  const uint32_t big_cap_used = cap_env ? (uint32_t)atoi(cap_env) : (uint32_t)big.size();
  a.big = big.data(); a.big_cap = big_cap_used < big.size() ? big_cap_used : (uint32_t)big.size();
  run_block(a);
  int rc = status;
  if (status == ST_OK && !ctr.overflow) {
    // what the extra CTAs of the scan launch do on the device: the tiles of the extents the planner only recorded
    for (uint32_t e = 0; e < ctr.n_big && e < a.big_cap; e++) {
      const BigExt &r = big[e];
      const uint32_t cnt = (uint32_t)((r.dst + r.len - 1) / TILE_BYTES - r.dst / TILE_BYTES + 1);
      for (uint32_t k = 0; k < cnt; k++) tiles[r.tile_index + k] = extent_tile(r.src, r.dst, r.len, r.file, k);
    }
    uint8_t *out = static_cast<uint8_t *>(malloc(out_size ? out_size : 1));
    memset(out, 0xA5, out_size);  // poison: every byte must be produced by exactly one tile
    std::vector<uint8_t> hit(out_size, 0);
    for (unsigned long long t = 0; t < ctr.n_tiles; t++) {
      const Tile &tl = tiles[t];
      if (tl.dst_rel + tl.len > out_size) { rc = -1000; break; }
      if (tl.src) memcpy(out + tl.dst_rel, reinterpret_cast<const void *>(tl.src), tl.len);
      else memset(out + tl.dst_rel, 0, tl.len);
      for (uint32_t q = 0; q < tl.len; q++) hit[tl.dst_rel + q]++;
    }
    for (uint64_t q = 0; q < out_size && rc == 0; q++) if (hit[q] != 1) rc = -1001;  // gap or double write
    if (rc == 0) { *out_p = out; *out_n = out_size; } else free(out);
  }
  free(arena); free(scratch);
  return rc;
}
extern "C" void lb2emu_free(uint8_t *p) { free(p); }
extern "C" void lb2emu_path_counts(int *out) { for (int k = 0; k < 4; k++) out[k] = lb2::lb2_path_counts[k]; }
