// api.cu -- C ABI (include/lambdipy_b200.h) over the plan / scan / compaction kernels, the
// chunked host pipeline (pinned H2D -> kernels -> D2H on rotating streams) and the in-place tree
// walker that stands where the reference runs `find ... -name "*.so" | xargs strip`
// (/root/reference/lambdipy/project_build.py:260).
#include "lb2_common.cuh"
#include "../../include/lambdipy_b200.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <climits>
#include <dirent.h>
#include <fcntl.h>
#include <spawn.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

extern char **environ;

using namespace lb2;

static thread_local std::string g_create_error;

// Device workspace of one batch in flight.
struct Workspace {
  uint64_t *d_in_off = nullptr, *d_in_size = nullptr, *d_out_size = nullptr, *d_out_off = nullptr;
  int32_t *d_status = nullptr;
  uint8_t *d_scratch = nullptr;
  Tile *d_tiles = nullptr;
  BatchCounters *d_ctr = nullptr;
  uint64_t *h_stage = nullptr;  // pinned: 2*(n+1) offsets/sizes up, counters + total down
  BatchCounters *h_ctr = nullptr;
  uint32_t cap_files = 0;
  uint64_t cap_tiles = 0;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  // last batch
  uint32_t n_files = 0;
  cudaStream_t stream = nullptr;
  bool in_flight = false;
};

struct lb2_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  Workspace ws;               // lb2_strip_device_async / lb2_plan_device
  std::string err;
  int compact_ctas_per_sm = 4;
  int use_tma = 1;             // bulk-copy engine kernel (0.97 of copy peak) ; LB2_COMPACT_TMA=0 selects the LSU kernel (0.90)
  // host pipeline slots
  struct Slot {
    Workspace ws;
    cudaStream_t stream = nullptr;
    uint8_t *d_in = nullptr, *d_out = nullptr;
    uint64_t cap_in = 0, cap_out = 0;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr}, ev_planned = nullptr;
  } slot[3];
  // tree: reusable pinned arenas
  uint8_t *h_tree_in = nullptr, *h_tree_out = nullptr;
  uint64_t cap_tree_in = 0, cap_tree_out = 0;
};

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                               \
      return LB2_E_CUDA;                                                                           \
    }                                                                                              \
  } while (0)

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void ws_free(Workspace &w) {
  cudaFree(w.d_in_off); cudaFree(w.d_in_size); cudaFree(w.d_out_size); cudaFree(w.d_out_off); cudaFree(w.d_status);
  cudaFree(w.d_scratch); cudaFree(w.d_tiles); cudaFree(w.d_ctr);
  if (w.h_stage) cudaFreeHost(w.h_stage);
  if (w.h_ctr) cudaFreeHost(w.h_ctr);
  for (auto &e : w.ev) if (e) cudaEventDestroy(e);
  w = Workspace();
}

static int ws_reserve(lb2_ctx *ctx, Workspace &w, uint32_t n_files, uint64_t n_tiles) {
  if (!w.d_ctr) {
    CK(cudaMalloc(&w.d_ctr, sizeof(BatchCounters)));
    CK(cudaHostAlloc(&w.h_ctr, sizeof(BatchCounters) + 64, cudaHostAllocDefault));
    for (auto &e : w.ev) CK(cudaEventCreate(&e));
  }
  if (n_files > w.cap_files) {
    uint32_t cap = std::max<uint32_t>(n_files, 256u);
    cap = std::max<uint32_t>(cap, w.cap_files + w.cap_files / 2);
    cudaFree(w.d_in_off); cudaFree(w.d_in_size); cudaFree(w.d_out_size); cudaFree(w.d_out_off); cudaFree(w.d_status); cudaFree(w.d_scratch);
    if (w.h_stage) cudaFreeHost(w.h_stage);
    w.cap_files = 0;
    CK(cudaMalloc(&w.d_in_off, (cap + 1) * sizeof(uint64_t)));
    CK(cudaMalloc(&w.d_in_size, (cap + 1) * sizeof(uint64_t)));
    CK(cudaMalloc(&w.d_out_size, (cap + 1) * sizeof(uint64_t)));
    CK(cudaMalloc(&w.d_out_off, (cap + 1) * sizeof(uint64_t)));
    CK(cudaMalloc(&w.d_status, (cap + 1) * sizeof(int32_t)));
    CK(cudaMalloc(&w.d_scratch, (uint64_t)cap * SCR_STRIDE));
    CK(cudaHostAlloc(&w.h_stage, (2ull * cap + 2) * sizeof(uint64_t), cudaHostAllocDefault));
    w.cap_files = cap;
  }
  if (n_tiles > w.cap_tiles) {
    uint64_t cap = std::max<uint64_t>(n_tiles, w.cap_tiles + w.cap_tiles / 2);
    cudaFree(w.d_tiles);
    w.cap_tiles = 0;
    CK(cudaMalloc(&w.d_tiles, cap * sizeof(Tile)));
    w.cap_tiles = cap;
  }
  return LB2_OK;
}

// Upper bound on the tiles a batch can emit: an extent of l bytes yields at most l / TILE + 2 tiles, a
// file has at most MAX_EXT extents, and re-laid-out files may grow (LOAD alignment padding) -- 1 GB of
// growth per batch is allowed for before the plan kernel reports overflow.
static uint64_t tile_bound(const uint64_t *sizes, uint32_t n) {
  uint64_t t = 65536;
  for (uint32_t i = 0; i < n; i++) t += sizes[i] / TILE_BYTES + 2 * MAX_EXT + 16;
  return t;
}

// Enqueue plan -> scan -> (compact) for one batch on `s`.  h_off/h_sizes are host arrays.
static int enqueue_batch(lb2_ctx *ctx, Workspace &w, const uint8_t *d_in, const uint64_t *h_off, const uint64_t *h_sizes,
                         uint32_t n, uint8_t *d_out, uint64_t out_cap, uint32_t flags, cudaStream_t s, bool compact) {
  for (uint32_t i = 0; i < n; i++)
    if (h_off[i] & 15) { ctx->err = "input offsets must be multiples of 16"; return LB2_E_ARG; }
  // sizes -> staging (pinned), upload
  std::vector<uint64_t> tmp;
  if (w.in_flight) { ctx->err = "previous batch on this workspace not collected"; return LB2_E_STATE; }
  uint64_t *st_off = nullptr, *st_size = nullptr;
  {
    // need sizes before reserve to bound tiles
    tmp.resize(n);
    for (uint32_t i = 0; i < n; i++) tmp[i] = h_sizes ? h_sizes[i] : (h_off[i + 1] - h_off[i]);
  }
  int rc = ws_reserve(ctx, w, n, tile_bound(tmp.data(), n));
  if (rc) return rc;
  st_off = w.h_stage;
  st_size = w.h_stage + (w.cap_files + 1);
  memcpy(st_off, h_off, (size_t)n * sizeof(uint64_t));
  memcpy(st_size, tmp.data(), (size_t)n * sizeof(uint64_t));
  CK(cudaMemcpyAsync(w.d_in_off, st_off, (size_t)n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(w.d_in_size, st_size, (size_t)n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(w.d_ctr, 0, sizeof(BatchCounters), s));
  CK(cudaEventRecord(w.ev[0], s));
  PlanArgs pa;
  pa.in = d_in; pa.in_off = w.d_in_off; pa.in_size = w.d_in_size; pa.n_files = n; pa.flags = flags;
  pa.scratch = w.d_scratch; pa.out_size = w.d_out_size; pa.status = w.d_status;
  pa.tiles = w.d_tiles; pa.tile_cap = w.cap_tiles; pa.ctr = w.d_ctr;
  launch_plan(pa, s);
  launch_scan(w.d_out_size, w.d_out_off, n, compact ? out_cap : ~0ull, w.d_ctr, s);
  CK(cudaEventRecord(w.ev[1], s));
  if (compact) {
    CompactArgs ca;
    ca.tiles = w.d_tiles; ca.ctr = w.d_ctr; ca.out_off = w.d_out_off; ca.out = d_out;
    if (ctx->use_tma) launch_compact_tma(ca, ctx->sm_count, s);
    else launch_compact(ca, ctx->sm_count * ctx->compact_ctas_per_sm, s);
  }
  CK(cudaEventRecord(w.ev[2], s));
  CK(cudaMemcpyAsync(w.h_ctr, w.d_ctr, sizeof(BatchCounters), cudaMemcpyDeviceToHost, s));
  CK(cudaGetLastError());
  w.n_files = n;
  w.stream = s;
  w.in_flight = true;
  return LB2_OK;
}

static int collect_batch(lb2_ctx *ctx, Workspace &w, uint64_t *h_out_off, uint64_t *h_out_sizes, int32_t *h_status,
                         lb2_stats *stats) {
  if (!w.in_flight) { ctx->err = "no batch in flight"; return LB2_E_STATE; }
  const uint32_t n = w.n_files;
  cudaStream_t s = w.stream;
  if (h_out_off) CK(cudaMemcpyAsync(h_out_off, w.d_out_off, (size_t)(n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
  if (h_out_sizes && n) CK(cudaMemcpyAsync(h_out_sizes, w.d_out_size, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
  if (h_status && n) CK(cudaMemcpyAsync(h_status, w.d_status, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  uint64_t *h_total = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(w.h_ctr) + sizeof(BatchCounters));
  CK(cudaMemcpyAsync(h_total, w.d_out_off + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  w.in_flight = false;
  if (stats) {
    const BatchCounters &c = *w.h_ctr;
    memset(stats, 0, sizeof(*stats));
    stats->n_files = n; stats->n_ok = c.n_ok; stats->n_unsupported = c.n_unsupported; stats->overflow = c.overflow;
    stats->in_bytes = c.in_bytes; stats->out_bytes = c.out_bytes; stats->copy_bytes = c.copy_bytes;
    stats->header_bytes = c.header_bytes; stats->n_tiles = c.n_tiles; stats->out_bytes_needed = *h_total;
    cudaEventElapsedTime(&stats->plan_ms, w.ev[0], w.ev[1]);
    cudaEventElapsedTime(&stats->compact_ms, w.ev[1], w.ev[2]);
  }
  if (w.h_ctr->overflow) { ctx->err = "output arena (or tile buffer) too small for this batch"; return LB2_E_CAPACITY; }
  return LB2_OK;
}

// ============================================================================ C ABI
extern "C" {

const char *lb2_version(void) { return "lambdipy_b200 0.1 (sm_100a; GNU strip 2.42 semantics)"; }

int lb2_ctx_create(int device, lb2_ctx **out) {
  if (!out) return LB2_E_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return LB2_E_NODEVICE;
  }
  if (device < 0 || device >= count) { g_create_error = "device index out of range"; return LB2_E_ARG; }
  lb2_ctx *ctx = new lb2_ctx();
  ctx->device = device;
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    g_create_error = std::string("cuda init: ") + cudaGetErrorString(e);
    delete ctx;
    return LB2_E_CUDA;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  ctx->sm_count = prop.multiProcessorCount;
  if (prop.major < 10) {
    g_create_error = "this library is built for sm_100a (B200) only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return LB2_E_NODEVICE;
  }
  if (const char *v = getenv("LB2_COMPACT_CTAS_PER_SM")) ctx->compact_ctas_per_sm = std::max(1, atoi(v));
  if (const char *v = getenv("LB2_COMPACT_TMA")) ctx->use_tma = atoi(v);
  *out = ctx;
  return LB2_OK;
}

void lb2_ctx_destroy(lb2_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  ws_free(ctx->ws);
  for (auto &sl : ctx->slot) {
    ws_free(sl.ws);
    cudaFree(sl.d_in); cudaFree(sl.d_out);
    if (sl.stream) cudaStreamDestroy(sl.stream);
    for (auto &e : sl.ev_h2d) if (e) cudaEventDestroy(e);
    for (auto &e : sl.ev_d2h) if (e) cudaEventDestroy(e);
    if (sl.ev_planned) cudaEventDestroy(sl.ev_planned);
  }
  if (ctx->h_tree_in) cudaFreeHost(ctx->h_tree_in);
  if (ctx->h_tree_out) cudaFreeHost(ctx->h_tree_out);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char *lb2_last_error(const lb2_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
int lb2_sm_count(const lb2_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

void *lb2_dev_alloc(lb2_ctx *ctx, uint64_t bytes) {
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes ? bytes : 256);
  if (e != cudaSuccess) { if (ctx) ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return nullptr; }
  return p;
}
void lb2_dev_free(lb2_ctx *, void *p) { if (p) cudaFree(p); }
void *lb2_pinned_alloc(lb2_ctx *ctx, uint64_t bytes) {
  void *p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 256, cudaHostAllocMapped | cudaHostAllocPortable);
  if (e != cudaSuccess) { if (ctx) ctx->err = std::string("cudaHostAlloc: ") + cudaGetErrorString(e); return nullptr; }
  return p;
}
void lb2_pinned_free(lb2_ctx *, void *p) { if (p) cudaFreeHost(p); }
int lb2_memcpy_h2d(lb2_ctx *ctx, void *d, const void *h, uint64_t n) { CK(cudaMemcpy(d, h, n, cudaMemcpyHostToDevice)); return LB2_OK; }
int lb2_memcpy_d2h(lb2_ctx *ctx, void *h, const void *d, uint64_t n) { CK(cudaMemcpy(h, d, n, cudaMemcpyDeviceToHost)); return LB2_OK; }
int lb2_memset_d(lb2_ctx *ctx, void *d, int v, uint64_t n) { CK(cudaMemset(d, v, n)); return LB2_OK; }

int lb2_strip_device_async(lb2_ctx *ctx, const void *d_in, const uint64_t *h_in_off, const uint64_t *h_in_sizes,
                           uint32_t n_files, void *d_out, uint64_t out_capacity, uint32_t flags, void *stream) {
  if (!ctx || !d_in || !h_in_off || !d_out) { if (ctx) ctx->err = "NULL argument"; return LB2_E_ARG; }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
  return enqueue_batch(ctx, ctx->ws, static_cast<const uint8_t *>(d_in), h_in_off, h_in_sizes, n_files,
                       static_cast<uint8_t *>(d_out), out_capacity, flags, s, true);
}

int lb2_batch_results(lb2_ctx *ctx, uint64_t *h_out_off, uint64_t *h_out_sizes, int32_t *h_status, lb2_stats *stats) {
  if (!ctx) return LB2_E_ARG;
  return collect_batch(ctx, ctx->ws, h_out_off, h_out_sizes, h_status, stats);
}

int lb2_plan_device(lb2_ctx *ctx, const void *d_in, const uint64_t *h_in_off, const uint64_t *h_in_sizes, uint32_t n_files,
                    uint32_t flags, uint64_t *h_out_sizes, int32_t *h_status, lb2_stats *stats) {
  if (!ctx || !d_in || !h_in_off) { if (ctx) ctx->err = "NULL argument"; return LB2_E_ARG; }
  CK(cudaSetDevice(ctx->device));
  int rc = enqueue_batch(ctx, ctx->ws, static_cast<const uint8_t *>(d_in), h_in_off, h_in_sizes, n_files, nullptr, 0, flags,
                         ctx->stream, false);
  if (rc) return rc;
  return collect_batch(ctx, ctx->ws, nullptr, h_out_sizes, h_status, stats);
}

// ---------------------------------------------------------------------------- host pipeline
static uint64_t env_u64(const char *name, uint64_t dflt) {
  const char *v = getenv(name);
  return v ? strtoull(v, nullptr, 10) : dflt;
}

int lb2_strip_host(lb2_ctx *ctx, const void *h_in_v, const uint64_t *h_in_off, const uint64_t *h_in_sizes, uint32_t n_files,
                   void *h_out_v, uint64_t out_capacity, uint64_t *h_out_off, uint64_t *h_out_sizes, int32_t *h_status,
                   uint32_t flags, lb2_stats *stats) {
  if (!ctx || !h_in_v || !h_in_off || !h_out_v || !h_out_off || !h_out_sizes || !h_status) { if (ctx) ctx->err = "NULL argument"; return LB2_E_ARG; }
  CK(cudaSetDevice(ctx->device));
  const uint8_t *h_in = static_cast<const uint8_t *>(h_in_v);
  uint8_t *h_out = static_cast<uint8_t *>(h_out_v);
  // Zero-copy path: when both arenas are pinned and mapped into the device address space the
  // kernels work on them directly over PCIe -- the plan kernel pulls only headers, the compaction
  // kernel pulls only the KEPT extents and pushes the stripped files straight into host memory.
  // Dropped sections (.symtab/.strtab/.debug_*) never cross the bus, uploads and downloads run
  // concurrently in both directions, and nothing has to fit in HBM.
  if (env_u64("LB2_HOST_ZEROCOPY", 1) && n_files) {
    void *d_in_alias = nullptr, *d_out_alias = nullptr;
    cudaPointerAttributes ai, ao;
    const bool ok = cudaPointerGetAttributes(&ai, h_in) == cudaSuccess && ai.type == cudaMemoryTypeHost &&
                    cudaPointerGetAttributes(&ao, h_out) == cudaSuccess && ao.type == cudaMemoryTypeHost &&
                    cudaHostGetDevicePointer(&d_in_alias, const_cast<uint8_t *>(h_in), 0) == cudaSuccess &&
                    cudaHostGetDevicePointer(&d_out_alias, h_out, 0) == cudaSuccess;
    cudaGetLastError();  // clear "invalid value" from probing pageable memory
    if (ok) {
      int rc = enqueue_batch(ctx, ctx->ws, static_cast<const uint8_t *>(d_in_alias), h_in_off, h_in_sizes, n_files,
                             static_cast<uint8_t *>(d_out_alias), out_capacity, flags, ctx->stream, true);
      if (rc) return rc;
      lb2_stats st;
      rc = collect_batch(ctx, ctx->ws, h_out_off, h_out_sizes, h_status, &st);
      if (stats) *stats = st;
      return rc;
    }
  }
  const uint64_t chunk_bytes = env_u64("LB2_CHUNK_MB", 256) << 20;
  lb2_stats total;
  memset(&total, 0, sizeof total);
  total.n_files = n_files;

  // chunk boundaries: consecutive files, <= chunk_bytes of arena span each (a bigger file stands alone)
  struct Chunk { uint32_t f0, f1; uint64_t in_base, in_span; };
  std::vector<Chunk> chunks;
  for (uint32_t f = 0; f < n_files;) {
    uint32_t g = f + 1;
    while (g < n_files && h_in_off[g + 1] - h_in_off[f] <= chunk_bytes) g++;
    chunks.push_back({f, g, h_in_off[f], ((h_in_off[g] - h_in_off[f]) + 255) & ~255ull});
    f = g;
  }
  const int NS = 3;
  for (int k = 0; k < NS; k++) {
    auto &sl = ctx->slot[k];
    if (!sl.stream) {
      CK(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
      for (auto &e : sl.ev_h2d) CK(cudaEventCreate(&e));
      for (auto &e : sl.ev_d2h) CK(cudaEventCreate(&e));
      CK(cudaEventCreateWithFlags(&sl.ev_planned, cudaEventDisableTiming));
    }
  }
  std::vector<uint64_t> rel_off;
  uint64_t out_base = 0;  // running 256-aligned position in h_out
  int rc = LB2_OK;

  auto finish = [&](size_t ci) -> int {
    auto &sl = ctx->slot[ci % NS];
    const Chunk &c = chunks[ci];
    const uint32_t n = c.f1 - c.f0;
    // wait for plan+scan (+compact, same stream) and learn the chunk's output size
    lb2_stats st;
    std::vector<uint64_t> coff(n + 1);
    int r = collect_batch(ctx, sl.ws, coff.data(), h_out_sizes + c.f0, h_status + c.f0, &st);
    if (r == LB2_E_CAPACITY) {
      // device-side output slot too small (re-laid-out files can grow): enlarge and redo this chunk
      cudaFree(sl.d_out);
      sl.cap_out = 0;
      uint64_t need = st.out_bytes_needed + (1u << 20);
      CK(cudaMalloc(&sl.d_out, need));
      sl.cap_out = need;
      std::vector<uint64_t> ro(n + 1);
      for (uint32_t i = 0; i <= n; i++) ro[i] = h_in_off[c.f0 + i] - c.in_base;
      r = enqueue_batch(ctx, sl.ws, sl.d_in, ro.data(), h_in_sizes ? h_in_sizes + c.f0 : nullptr, n, sl.d_out, sl.cap_out, flags, sl.stream, true);
      if (r) return r;
      r = collect_batch(ctx, sl.ws, coff.data(), h_out_sizes + c.f0, h_status + c.f0, &st);
    }
    if (r) return r;
    const uint64_t bytes = coff[n];
    if (out_base + bytes > out_capacity) { ctx->err = "host output arena too small"; total.out_bytes_needed = out_base + bytes; return LB2_E_CAPACITY; }
    CK(cudaEventRecord(sl.ev_d2h[0], sl.stream));
    if (bytes) CK(cudaMemcpyAsync(h_out + out_base, sl.d_out, bytes, cudaMemcpyDeviceToHost, sl.stream));
    CK(cudaEventRecord(sl.ev_d2h[1], sl.stream));
    for (uint32_t i = 0; i < n; i++) h_out_off[c.f0 + i] = out_base + coff[i];
    out_base += bytes;
    total.n_ok += st.n_ok; total.n_unsupported += st.n_unsupported; total.in_bytes += st.in_bytes; total.out_bytes += st.out_bytes;
    total.copy_bytes += st.copy_bytes; total.header_bytes += st.header_bytes; total.n_tiles += st.n_tiles;
    total.plan_ms += st.plan_ms; total.compact_ms += st.compact_ms;
    return LB2_OK;
  };
  auto reap_copy_times = [&](size_t ci) {
    auto &sl = ctx->slot[ci % NS];
    cudaEventSynchronize(sl.ev_d2h[1]);
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, sl.ev_h2d[0], sl.ev_h2d[1]);
    cudaEventElapsedTime(&b, sl.ev_d2h[0], sl.ev_d2h[1]);
    total.h2d_ms += a; total.d2h_ms += b;
  };

  for (size_t ci = 0; ci < chunks.size() && rc == LB2_OK; ci++) {
    auto &sl = ctx->slot[ci % NS];
    const Chunk &c = chunks[ci];
    const uint32_t n = c.f1 - c.f0;
    if (ci >= (size_t)NS) reap_copy_times(ci - NS);  // slot is free once its D2H finished
    if (sl.cap_in < c.in_span + 256) {
      cudaFree(sl.d_in); sl.cap_in = 0;
      uint64_t need = std::max<uint64_t>(c.in_span + 256, std::min<uint64_t>(chunk_bytes, 64ull << 20));
      CK(cudaMalloc(&sl.d_in, need));
      sl.cap_in = need;
    }
    const uint64_t want_out = c.in_span + (uint64_t)n * 4096 + (8u << 20);
    if (sl.cap_out < want_out) {
      cudaFree(sl.d_out); sl.cap_out = 0;
      CK(cudaMalloc(&sl.d_out, want_out));
      sl.cap_out = want_out;
    }
    CK(cudaEventRecord(sl.ev_h2d[0], sl.stream));
    CK(cudaMemcpyAsync(sl.d_in, h_in + c.in_base, h_in_off[c.f1] - c.in_base, cudaMemcpyHostToDevice, sl.stream));
    CK(cudaEventRecord(sl.ev_h2d[1], sl.stream));
    rel_off.resize(n + 1);
    for (uint32_t i = 0; i <= n; i++) rel_off[i] = h_in_off[c.f0 + i] - c.in_base;
    rc = enqueue_batch(ctx, sl.ws, sl.d_in, rel_off.data(), h_in_sizes ? h_in_sizes + c.f0 : nullptr, n, sl.d_out, sl.cap_out, flags, sl.stream, true);
    if (rc) break;
    // with the next chunk's upload and kernels queued, turn to the previous chunk's download
    if (ci >= 1) rc = finish(ci - 1);
  }
  if (rc == LB2_OK && !chunks.empty()) rc = finish(chunks.size() - 1);
  for (size_t ci = chunks.size() >= (size_t)NS ? chunks.size() - NS : 0; ci < chunks.size(); ci++) reap_copy_times(ci);
  for (int k = 0; k < NS; k++) { cudaStreamSynchronize(ctx->slot[k].stream); ctx->slot[k].ws.in_flight = false; }
  h_out_off[n_files] = out_base;
  if (total.out_bytes_needed == 0) total.out_bytes_needed = out_base;
  if (stats) *stats = total;
  return rc;
}

}  // extern "C"

// ---------------------------------------------------------------------------- tree walker
static bool ends_with(const char *s, const char *suf) {
  size_t a = strlen(s), b = strlen(suf);
  return a >= b && memcmp(s + a - b, suf, b) == 0;
}

struct TreeFile { std::string path; uint64_t size; mode_t mode; uint32_t times; };

// `find ROOT/ -name "*SUFFIX"`: every directory entry whose basename matches, of any type; find does
// not descend into symlinked directories.  What `strip` then does with each path decides the rest:
//   regular file           -> stripped in place
//   symlink to a file      -> the TARGET is rewritten, the link stays (so libfoo.so -> libfoo.so.1
//                             strips libfoo.so.1 even though that name does not match)
//   directory / dangling   -> strip fails -> xargs exits 123 -> the reference's script aborts
// A file reached through k matching paths is stripped k times by the reference; `times` keeps k.
static void walk(const std::string &dir, const char *suffix, std::vector<TreeFile> &files, lb2_tree_stats *st) {
  DIR *d = opendir(dir.c_str());
  if (!d) return;
  while (dirent *e = readdir(d)) {
    if (!strcmp(e->d_name, ".") || !strcmp(e->d_name, "..")) continue;
    std::string p = dir + "/" + e->d_name;
    struct stat lsb, sb;
    if (lstat(p.c_str(), &lsb) != 0) continue;
    const bool match = ends_with(e->d_name, suffix);
    if (match) {
      st->n_selected++;
      if (stat(p.c_str(), &sb) != 0 || !S_ISREG(sb.st_mode)) {
        st->n_failed++;  // directory, dangling link, device ...: strip errors out
      } else {
        char real[PATH_MAX];
        if (!realpath(p.c_str(), real)) st->n_failed++;
        else {
          if (S_ISLNK(lsb.st_mode)) st->n_skipped++;  // the link itself is left alone
          files.push_back({real, (uint64_t)sb.st_size, sb.st_mode, 1});
        }
      }
    }
    if (S_ISDIR(lsb.st_mode)) walk(p, suffix, files, st);
  }
  closedir(d);
}

static void dedupe(std::vector<TreeFile> &files) {
  std::sort(files.begin(), files.end(), [](const TreeFile &a, const TreeFile &b) { return a.path < b.path; });
  size_t w = 0;
  for (size_t i = 0; i < files.size(); i++) {
    if (w && files[w - 1].path == files[i].path) files[w - 1].times++;
    else files[w++] = files[i];
  }
  files.resize(w);
}

// Files are read and written in <= 8 MB pieces so that one 900 MB library is handled by many I/O
// threads instead of one (the tree of BASELINE config 3 is dominated by two such files).
static const uint64_t IO_PIECE = 8ull << 20;

static bool read_piece(const std::string &p, uint8_t *dst, uint64_t off, uint64_t n) {
  int fd = open(p.c_str(), O_RDONLY | O_CLOEXEC);
  if (fd < 0) return false;
  uint64_t got = 0;
  while (got < n) {
    ssize_t r = pread(fd, dst + got, n - got, (off_t)(off + got));
    if (r <= 0) { if (r < 0 && errno == EINTR) continue; break; }
    got += (uint64_t)r;
  }
  close(fd);
  return got == n;
}

static bool write_piece(const char *tmp, const uint8_t *src, uint64_t off, uint64_t n) {
  int fd = open(tmp, O_WRONLY | O_CLOEXEC);
  if (fd < 0) return false;
  uint64_t put = 0;
  bool ok = true;
  while (put < n) {
    ssize_t r = pwrite(fd, src + put, n - put, (off_t)(off + put));
    if (r <= 0) { if (r < 0 && errno == EINTR) continue; ok = false; break; }
    put += (uint64_t)r;
  }
  close(fd);
  return ok;
}

static int host_strip(const std::string &p) {
  const char *argv[] = {"strip", p.c_str(), nullptr};
  pid_t pid;
  if (posix_spawnp(&pid, "strip", nullptr, nullptr, const_cast<char *const *>(argv), environ) != 0) return 127;
  int status = 0;
  while (waitpid(pid, &status, 0) < 0 && errno == EINTR) {}
  return WIFEXITED(status) ? WEXITSTATUS(status) : 128;
}

template <class F> static void parallel_for(size_t n, int threads, F f) {
  std::atomic<size_t> next{0};
  std::vector<std::thread> pool;
  threads = (int)std::min<size_t>((size_t)threads, n);
  for (int t = 0; t < threads; t++)
    pool.emplace_back([&] { for (size_t i; (i = next.fetch_add(1)) < n;) f(i); });
  for (auto &th : pool) th.join();
}

extern "C" {

int lb2_strip_tree(lb2_ctx *ctx, const char *root, const char *suffix, uint32_t flags, lb2_tree_stats *st_out) {
  if (!ctx || !root || !suffix) { if (ctx) ctx->err = "NULL argument"; return LB2_E_ARG; }
  CK(cudaSetDevice(ctx->device));
  lb2_tree_stats st;
  memset(&st, 0, sizeof st);
  double t0 = now_s();
  std::vector<TreeFile> files;
  struct stat rsb;
  if (stat(root, &rsb) != 0 || !S_ISDIR(rsb.st_mode)) { ctx->err = std::string("not a directory: ") + root; return LB2_E_IO; }
  std::string r = root;
  while (r.size() > 1 && r.back() == '/') r.pop_back();
  walk(r, suffix, files, &st);
  dedupe(files);
  const uint32_t n = (uint32_t)files.size();
  std::vector<uint64_t> off(n + 1), sizes(n), out_off(n + 1), out_sizes(n);
  std::vector<int32_t> status(n);
  uint64_t pos = 0;
  for (uint32_t i = 0; i < n; i++) { off[i] = pos; sizes[i] = files[i].size; pos += (files[i].size + 255) & ~255ull; }
  off[n] = pos;
  const uint64_t in_cap = pos + 256, out_cap = pos + (uint64_t)n * 4096 + (16u << 20);
  if (ctx->cap_tree_in < in_cap) {
    if (ctx->h_tree_in) cudaFreeHost(ctx->h_tree_in);
    ctx->cap_tree_in = 0;
    CK(cudaHostAlloc(&ctx->h_tree_in, in_cap, cudaHostAllocDefault));
    ctx->cap_tree_in = in_cap;
  }
  if (ctx->cap_tree_out < out_cap) {
    if (ctx->h_tree_out) cudaFreeHost(ctx->h_tree_out);
    ctx->cap_tree_out = 0;
    CK(cudaHostAlloc(&ctx->h_tree_out, out_cap, cudaHostAllocDefault));
    ctx->cap_tree_out = out_cap;
  }
  const int io_threads = (int)env_u64("LB2_IO_THREADS", std::max(4u, std::min(32u, std::thread::hardware_concurrency())));
  std::atomic<int> read_fail{0};
  struct Piece { uint32_t file; uint64_t off, len; };
  std::vector<Piece> rpieces;
  for (uint32_t i = 0; i < n; i++)
    for (uint64_t o = 0; o < sizes[i] || (o == 0 && sizes[i] == 0); o += IO_PIECE) {
      rpieces.push_back({i, o, std::min(IO_PIECE, sizes[i] - o)});
      if (sizes[i] == 0) break;
    }
  parallel_for(rpieces.size(), io_threads, [&](size_t k) {
    const Piece &pc = rpieces[k];
    if (pc.len && !read_piece(files[pc.file].path, ctx->h_tree_in + off[pc.file] + pc.off, pc.off, pc.len)) read_fail++;
  });
  if (read_fail) { ctx->err = "could not read some selected files"; return LB2_E_IO; }
  st.walk_read_s = now_s() - t0;

  t0 = now_s();
  int rc = LB2_OK;
  if (n) rc = lb2_strip_host(ctx, ctx->h_tree_in, off.data(), sizes.data(), n, ctx->h_tree_out, ctx->cap_tree_out, out_off.data(),
                             out_sizes.data(), status.data(), flags & 0xffu, &st.batch);
  // files the reference would strip more than once (reached through several matching names):
  // run the extra passes on the previous pass's output (strip is not always idempotent)
  for (uint32_t pass = 1; rc == LB2_OK; pass++) {
    std::vector<uint32_t> again;
    for (uint32_t i = 0; i < n; i++) if (files[i].times > pass && status[i] == LB2_ST_OK) again.push_back(i);
    if (again.empty()) break;
    const uint32_t m = (uint32_t)again.size();
    std::vector<uint64_t> off2(m + 1), sz2(m), ooff2(m + 1), osz2(m);
    std::vector<int32_t> st2(m);
    uint64_t p2 = 0;
    for (uint32_t k = 0; k < m; k++) { off2[k] = p2; sz2[k] = out_sizes[again[k]]; p2 += (sz2[k] + 255) & ~255ull; }
    off2[m] = p2;
    // previous outputs become inputs (the input arena is free to reuse: it is at least as large)
    for (uint32_t k = 0; k < m; k++) memcpy(ctx->h_tree_in + off2[k], ctx->h_tree_out + out_off[again[k]], sz2[k]);
    std::vector<uint8_t> keep_out;  // outputs of files not in this pass stay where they are; new ones go to a scratch arena
    uint8_t *h_tmp = nullptr;
    const uint64_t cap2 = p2 + (uint64_t)m * 4096 + (16u << 20);
    CK(cudaHostAlloc(&h_tmp, cap2, cudaHostAllocDefault));
    lb2_stats b2;
    rc = lb2_strip_host(ctx, ctx->h_tree_in, off2.data(), sz2.data(), m, h_tmp, cap2, ooff2.data(), osz2.data(), st2.data(), flags & 0xffu, &b2);
    if (rc == LB2_OK) {
      for (uint32_t k = 0; k < m; k++) {
        const uint32_t i = again[k];
        if (st2[k] != LB2_ST_OK) { status[i] = st2[k]; continue; }
        // a re-stripped file never grows beyond its 256-byte-rounded slot by more than the slack between files
        if (osz2[k] <= ((out_sizes[i] + 255) & ~255ull)) { memcpy(ctx->h_tree_out + out_off[i], h_tmp + ooff2[k], osz2[k]); out_sizes[i] = osz2[k]; }
        else status[i] = LB2_ST_UNSUPPORTED_LAYOUT;  // hand to the host strip
      }
    }
    cudaFreeHost(h_tmp);
  }
  st.gpu_s = now_s() - t0;
  if (rc) { if (st_out) *st_out = st; return rc; }

  t0 = now_s();
  std::atomic<uint32_t> n_gpu{0}, n_failed{0}, n_skipped{0};
  std::atomic<uint64_t> in_b{0}, out_b{0};
  std::vector<uint32_t> fallback;
  for (uint32_t i = 0; i < n; i++) if (status[i] != LB2_ST_OK) fallback.push_back(i);
  if (!(flags & LB2_TREE_DRY_RUN)) {
    // temp file next to the target (pre-sized), pieces written in parallel, then fchmod + rename by
    // whichever thread finishes the file's last piece -- what strip does, minus the single thread
    std::vector<std::string> tmp_path(n);

    std::vector<std::atomic<int>> remaining(n);
    std::vector<std::atomic<int>> piece_fail(n);
    std::vector<Piece> wpieces;
    for (uint32_t i = 0; i < n; i++) {
      remaining[i] = 0; piece_fail[i] = 0;
      if (status[i] != LB2_ST_OK) continue;
      std::string t = files[i].path + ".lb2XXXXXX";
      std::vector<char> tb(t.begin(), t.end());
      tb.push_back(0);
      int fd = mkstemp(tb.data());
      if (fd < 0) { n_failed++; continue; }
      if (ftruncate(fd, (off_t)out_sizes[i]) != 0) { close(fd); unlink(tb.data()); n_failed++; continue; }
      close(fd);
      tmp_path[i] = tb.data();
      int cnt = 0;
      for (uint64_t o = 0; o < out_sizes[i]; o += IO_PIECE) { wpieces.push_back({i, o, std::min(IO_PIECE, out_sizes[i] - o)}); cnt++; }
      if (cnt == 0) { wpieces.push_back({i, 0, 0}); cnt = 1; }
      remaining[i] = cnt;
    }
    parallel_for(wpieces.size(), io_threads, [&](size_t k) {
      const Piece &pc = wpieces[k];
      const uint32_t i = pc.file;
      // (pieces of one file still serialise on the inode lock inside the kernel -- measured: a shared
      //  mmap is no faster on tmpfs -- but pieces of different files, and all reads, run in parallel)
      if (pc.len && !write_piece(tmp_path[i].c_str(), ctx->h_tree_out + out_off[i] + pc.off, pc.off, pc.len)) piece_fail[i]++;
      if (--remaining[i] == 0) {
        bool ok = piece_fail[i] == 0 && chmod(tmp_path[i].c_str(), files[i].mode & 07777) == 0 &&
                  rename(tmp_path[i].c_str(), files[i].path.c_str()) == 0;
        if (ok) { n_gpu++; in_b += sizes[i]; out_b += out_sizes[i]; }
        else { unlink(tmp_path[i].c_str()); n_failed++; }
      }
    });
  } else {
    for (uint32_t i = 0; i < n; i++) if (status[i] == LB2_ST_OK) { n_gpu++; in_b += sizes[i]; out_b += out_sizes[i]; }
  }
  st.write_s = now_s() - t0;

  t0 = now_s();
  std::atomic<uint32_t> n_fb{0};
  parallel_for(fallback.size(), io_threads, [&](size_t k) {
    const uint32_t i = fallback[k];
    const bool non_elf = status[i] == LB2_ST_NOT_ELF;
    if (non_elf && (flags & LB2_TREE_TOLERATE_NON_ELF)) { n_skipped++; return; }
    if ((flags & LB2_TREE_FALLBACK_HOST_STRIP) && !(flags & LB2_TREE_DRY_RUN)) {
      // the reference's own tool decides (and fails the build exactly when the reference would)
      bool ok = true;
      for (uint32_t k = 0; k < files[i].times && ok; k++) ok = host_strip(files[i].path) == 0;
      if (ok) n_fb++; else n_failed++;
    } else {
      n_failed++;
    }
  });
  st.fallback_s = now_s() - t0;
  st.n_gpu = n_gpu; st.n_fallback = n_fb; st.n_failed += n_failed; st.n_skipped += n_skipped;
  st.in_bytes = in_b; st.out_bytes = out_b;
  if (st_out) *st_out = st;
  return LB2_OK;
}

// ---------------------------------------------------------------------------- corpus fill
int lb2_corpus_fill(lb2_ctx *ctx, void *d_arena, const lb2_fill_region *h_regions, uint32_t n_regions, uint64_t seed, void *stream) {
  if (!ctx || !d_arena || (!h_regions && n_regions)) { if (ctx) ctx->err = "NULL argument"; return LB2_E_ARG; }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
  if (!n_regions) return LB2_OK;
  // split regions into <= 4 MB pieces so the grid balances
  std::vector<lb2_fill_region> pieces;
  const uint64_t P = 4ull << 20;
  for (uint32_t i = 0; i < n_regions; i++)
    for (uint64_t o = 0; o < h_regions[i].len; o += P) pieces.push_back({h_regions[i].offset + o, std::min(P, h_regions[i].len - o)});
  lb2_fill_region *d_regions = nullptr;
  CK(cudaMalloc(&d_regions, pieces.size() * sizeof(lb2_fill_region)));
  CK(cudaMemcpyAsync(d_regions, pieces.data(), pieces.size() * sizeof(lb2_fill_region), cudaMemcpyHostToDevice, s));
  launch_fill(static_cast<uint8_t *>(d_arena), reinterpret_cast<const FillRegion *>(d_regions), (uint32_t)pieces.size(), seed, ctx->sm_count * 8, s);
  CK(cudaStreamSynchronize(s));
  CK(cudaFree(d_regions));
  CK(cudaGetLastError());
  return LB2_OK;
}

int lb2_corpus_scatter(lb2_ctx *ctx, void *d_arena, const void *h_data, uint64_t data_bytes, const uint64_t *h_dst,
                       const uint64_t *h_src, const uint64_t *h_len, uint32_t n) {
  if (!ctx || !d_arena || (n && (!h_data || !h_dst || !h_src || !h_len))) { if (ctx) ctx->err = "NULL argument"; return LB2_E_ARG; }
  CK(cudaSetDevice(ctx->device));
  if (!n) return LB2_OK;
  cudaStream_t s = ctx->stream;
  uint8_t *d_stage = nullptr;
  CK(cudaMalloc(&d_stage, data_bytes + 256));
  CK(cudaMemcpyAsync(d_stage, h_data, data_bytes, cudaMemcpyHostToDevice, s));
  std::vector<Tile> tiles;
  for (uint32_t i = 0; i < n; i++)
    for (uint64_t o = 0; o < h_len[i]; o += TILE_BYTES) {
      Tile t;
      t.src = reinterpret_cast<uint64_t>(d_stage) + h_src[i] + o;
      t.dst_rel = h_dst[i] + o;
      t.len = (uint32_t)std::min<uint64_t>(TILE_BYTES, h_len[i] - o);
      t.file = 0;
      tiles.push_back(t);
    }
  Tile *d_tiles = nullptr;
  BatchCounters *d_ctr = nullptr;
  uint64_t *d_off = nullptr;
  CK(cudaMalloc(&d_tiles, tiles.size() * sizeof(Tile)));
  CK(cudaMalloc(&d_ctr, sizeof(BatchCounters)));
  CK(cudaMalloc(&d_off, sizeof(uint64_t)));
  BatchCounters c;
  memset(&c, 0, sizeof c);
  c.n_tiles = tiles.size();
  uint64_t zero = 0;
  CK(cudaMemcpyAsync(d_tiles, tiles.data(), tiles.size() * sizeof(Tile), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_ctr, &c, sizeof c, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_off, &zero, sizeof zero, cudaMemcpyHostToDevice, s));
  CompactArgs ca;
  ca.tiles = d_tiles; ca.ctr = d_ctr; ca.out_off = d_off; ca.out = static_cast<uint8_t *>(d_arena);
  launch_compact(ca, ctx->sm_count * 4, s);
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  cudaFree(d_stage); cudaFree(d_tiles); cudaFree(d_ctr); cudaFree(d_off);
  return LB2_OK;
}

}  // extern "C"
