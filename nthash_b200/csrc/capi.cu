// capi.cu — the extern "C" boundary declared in include/nthash_b200.h.
// Argument validation mirrors the reference constructors (src/kmer.cpp:212-225,
// src/seed.cpp:85-104, :467-469) but reports through return codes instead of exit(1).
#include "../../include/nthash_b200.h"
#include "engine.hpp"
#include "seed_plan.hpp"

#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace nthb {

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define NTH_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t e__ = (expr);                                                                       \
    if (e__ != cudaSuccess) return fail(NTHASH_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

uint32_t span_bound(uint32_t seg, uint32_t segs, uint32_t k)
{
  // KMER_NT consecutive items advance by seg bytes each, plus a (k-1)-base tail per read end crossed
  const uint64_t b = (uint64_t)KMER_NT * seg + ((uint64_t)KMER_NT / segs + 2) * (k - 1) + 64;
  return b > 0xffffffffull ? 0xffffffffu : (uint32_t)b;
}

static const uint32_t SMEM_MAX = 227 * 1024;

bool plan_uniform(uint64_t n_reads, uint32_t read_len, uint32_t k, KmerGeom& g, uint32_t& tile_cap)
{
  g = KmerGeom();
  g.read_len = read_len;
  g.nk = read_len >= k ? read_len - k + 1 : 0;
  if (g.nk == 0 || n_reads == 0) {
    g.n_items = 0;
    tile_cap = 0;
    return true;
  }
  if ((uint64_t)KMER_NT * read_len + 64 <= TILE_BUDGET) { // one item per read
    g.seg = g.nk;
    g.segs = 1;
    tile_cap = KMER_NT * read_len + 64;
  } else {
    // Cut reads into items.  Prefer a segment length that divides nk (all items full => the TMA tile
    // store applies) and whose byte stride spreads lanes over shared-memory banks.
    uint32_t best = 0, best_score = 0;
    if (const char* e = getenv("NTHASH_B200_SEG")) { // experiments: force the segment length
      const uint32_t v = (uint32_t)atoi(e);
      if (v >= 16 && v % 2 == 0 && g.nk % v == 0) best = v, best_score = 0xffffffffu;
    }
    for (uint32_t seg = 160; seg <= 384; seg += 2) {
      if (g.nk % seg) continue;
      const uint32_t score = (seg % 8 == 4 ? 4 : seg % 4 == 2 ? 3 : seg % 16 == 8 ? 2 : 1) * 1000 - (seg > SEG_LONG ? seg - SEG_LONG : SEG_LONG - seg);
      if (score > best_score) {
        best_score = score;
        best = seg;
      }
    }
    g.seg = best ? best : SEG_LONG;
    g.segs = (g.nk + g.seg - 1) / g.seg;
    tile_cap = span_bound(g.seg, g.segs, k);
  }
  g.n_items = n_reads * g.segs;
  return kmer_smem_bytes(tile_cap) <= SMEM_MAX;
}

static int check_kh(uint32_t k, uint32_t h)
{
  if (k < 3 || k > 65535) return fail(NTHASH_ERR_INVALID_ARG, "k=%u outside the supported range [3, 65535]", k);
  if (h < 1 || h > 255) return fail(NTHASH_ERR_INVALID_ARG, "num_hashes=%u outside [1, 255]", h);
  return NTHASH_OK;
}

static int check_outputs(const void* out, const void* fwd, const void* rev)
{
  if (!out) return fail(NTHASH_ERR_INVALID_ARG, "out must not be NULL");
  if ((fwd == nullptr) != (rev == nullptr))
    return fail(NTHASH_ERR_INVALID_ARG, "out_fwd and out_rev must be given together");
  return NTHASH_OK;
}

static int check_device_ready()
{
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(NTHASH_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10)
    return fail(NTHASH_ERR_NO_DEVICE, "device %d is sm_%d?: the engine is built for sm_100a only", dev, major);
  // Scratch (item tables, staging buffers) comes from the device's stream-ordered pool; tell it once to keep freed
  // memory for the next call instead of returning it to the driver at every synchronisation (a 200 MB item table
  // for ragged long reads cost ~20 ms per call that way).
  static std::atomic<uint64_t> pool_tuned{ 0 };
  if (dev < 64 && !(pool_tuned.load() >> dev & 1)) {
    cudaMemPool_t pool = nullptr;
    uint64_t keep = ~0ull;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    pool_tuned.fetch_or(1ull << dev);
  }
  return NTHASH_OK;
}

// Pre-sets the validity bitmap: `rows` known on the host, or (ragged *_dev entries, which must not read anything back)
// taken from device memory at *d_rows with `rows_bound` only sizing the grid.
static cudaError_t preset_valid(uint32_t* d_valid, uint64_t rows, const uint64_t* d_rows, uint64_t rows_bound, cudaStream_t st)
{
  if (!d_valid) return cudaSuccess;
  if (d_rows) return launch_fill_valid(d_valid, d_rows, rows_bound, st);
  return rows ? cudaMemsetAsync(d_valid, 0xFF, ((rows + 31) / 32) * 4, st) : cudaSuccess;
}

static int run_kmer(KmerParams& P, uint64_t rows, const uint64_t* d_rows, uint64_t rows_bound, cudaStream_t st)
{
  if (getenv("NTHASH_B200_DISABLE_TMA_STORE")) P.use_tma = false; // A/B switch for tests and profiling
  NTH_CUDA(preset_valid(P.valid_bits, rows, d_rows, rows_bound, st));
  if (P.g.n_items == 0) return NTHASH_OK;
  const cudaError_t e = launch_kmer(P, st);
  if (e == cudaErrorInvalidConfiguration || e == cudaErrorNotSupported) {
    cudaGetLastError();
    return fail(NTHASH_ERR_UNSUPPORTED, "k=%u, num_hashes=%u on reads of up to %u bases: no kernel fits this request in shared memory%s", P.k,
                P.h, P.g.read_len, P.bloom_mode ? " (the Bloom consumer has no general-kernel form)" : "");
  }
  NTH_CUDA(e);
  return NTHASH_OK;
}


// Ragged batches: either every read is one item (read_off/koff are the item arrays) or reads are cut
// into SEG_LONG-window items listed in a scratch table (freed by the caller with cudaFreeAsync).
struct RaggedItems
{
  KmerGeom g;
  uint32_t tile_cap = 0;
  uint32_t tile_cap_256 = 0;           // KmerParams::tile_cap_256 (plan handles only: it takes a read-back)
  uint64_t* d_items = nullptr;         // [item_byte | item_out | item_read]
  const uint64_t* item_read = nullptr; // NULL when items are reads
  uint8_t* d_perm = nullptr;           // KmerGeom::item_perm (the items of every block of 256 in length-class order)
  void release_async(cudaStream_t st)
  {
    if (d_items) cudaFreeAsync(d_items, st);
    if (d_perm) cudaFreeAsync(d_perm, st);
    d_items = nullptr;
    d_perm = nullptr;
  }
};

// the block-wise length-class order the fast kernel deals its items out by (without it every CTA sorts its own)
static int plan_item_perm(RaggedItems& R, cudaStream_t st)
{
  if (R.g.n_items == 0 || getenv("NTHASH_B200_NO_ITEM_PERM")) return NTHASH_OK;
  NTH_CUDA(cudaMallocAsync(&R.d_perm, (R.g.n_items + 255) / 256 * 256, st));
  const cudaError_t e = launch_item_perm(R.g.item_out, R.g.n_items, R.d_perm, st);
  if (e != cudaSuccess) {
    cudaFreeAsync(R.d_perm, st);
    R.d_perm = nullptr;
    NTH_CUDA(e);
  }
  R.g.item_perm = R.d_perm;
  return NTHASH_OK;
}

// Nothing is read back: the item tables are sized by a host-side bound of the item count (every read contributes at
// most ceil(windows / SEG_LONG) <= bases / SEG_LONG + 1 items) and the surplus is padded with empty items.
static int plan_ragged(const uint64_t* d_read_off, const uint64_t* d_koff, uint64_t n_reads, uint64_t max_read_len,
                       uint64_t n_bases, uint32_t k, uint32_t tile_budget, cudaStream_t st, RaggedItems& R, bool want_perm = true)
{
  if ((uint64_t)KMER_NT * max_read_len + 64 <= tile_budget) { // every read is one item
    R.g.item_byte = d_read_off;
    R.g.item_out = d_koff;
    R.g.n_items = n_reads;
    R.tile_cap = (uint32_t)(KMER_NT * max_read_len + 64);
    return want_perm ? plan_item_perm(R, st) : NTHASH_OK;
  }
  R.tile_cap = span_bound(SEG_LONG, 1, k);
  const uint64_t cap = n_reads + n_bases / SEG_LONG + 1;
  R.g.n_items = cap;
  NTH_CUDA(cudaMallocAsync(&R.d_items, 3 * (cap + 1) * sizeof(uint64_t), st));
  const cudaError_t e = launch_item_fill(d_read_off, d_koff, n_reads, k, SEG_LONG, R.d_items, R.d_items + cap + 1,
                                         R.d_items + 2 * (cap + 1), cap, st);
  if (e != cudaSuccess) {
    cudaFreeAsync(R.d_items, st);
    R.d_items = nullptr;
    NTH_CUDA(e);
  }
  R.g.item_byte = R.d_items;
  R.g.item_out = R.d_items + cap + 1;
  R.item_read = R.d_items + 2 * (cap + 1);
  if (int rc = want_perm ? plan_item_perm(R, st) : NTHASH_OK) {
    cudaFreeAsync(R.d_items, st);
    R.d_items = nullptr;
    return rc;
  }
  return NTHASH_OK;
}

} // namespace nthb

// Opaque handle of include/nthash_b200.h: a compiled seed set resident on one device.
struct nthash_seed_plan
{
  nthb::SeedPlanHost host;
  uint8_t* d_blob = nullptr;
  int device = 0;
  nthb::SeedJit* jit = nullptr; // specialised kernel, or NULL (then jit_note says why)
  std::string jit_note;
};

// Opaque handle of include/nthash_b200.h: the layout of one ragged batch (dense row offsets, totals, item tables),
// computed once so that steady-state calls neither re-plan nor synchronise.
struct nthash_ragged_plan
{
  int device = 0;
  uint32_t k = 0;
  uint64_t n_reads = 0, rows = 0, max_len = 0, n_bases = 0;
  const uint64_t* d_read_off = nullptr; // borrowed: the caller keeps it alive and unchanged
  uint64_t* d_koff = nullptr;           // owned, n_reads + 1
  nthb::RaggedItems items;              // for the k-mer kernels' tile budget (item table owned when reads are cut up)
};

namespace nthb {

static void fill_seed_params(const nthash_seed_plan* plan, SeedParams& P)
{
  const SeedPlanHost& h = plan->host;
  P.k = h.k;
  P.h = h.h;
  P.n_seeds = h.n_seeds;
  P.plan_blob = plan->d_blob;
  P.plan_smem_bytes = h.smem_bytes;
  P.groups_off = h.groups_off;
  P.tables_off = h.tables_off;
  P.care_off = h.care_off;
  P.refblk_off = h.refblk_off;
  P.care_words = h.care_words;
  P.any_ignore = h.any_ignore ? 1u : 0u;
}

static int run_seed(const nthash_seed_plan* plan, SeedParams& P, uint64_t n_reads, uint64_t rows, const uint64_t* d_rows,
                    uint64_t rows_bound, cudaStream_t st)
{
  NTH_CUDA(preset_valid(P.valid_bits, rows, d_rows, rows_bound, st));
  if (P.g.n_items == 0) return NTHASH_OK;
  if (seed_smem_bytes(P.plan_smem_bytes, P.tile_cap) > SMEM_MAX)
    return fail(NTHASH_ERR_UNSUPPORTED, "seed tables (%u B) plus a %u-byte base tile exceed shared memory",
                P.plan_smem_bytes, P.tile_cap);
  uint8_t* d_dirty = nullptr;
  NTH_CUDA(cudaMallocAsync(&d_dirty, n_reads, st));
  cudaError_t e = cudaMemsetAsync(d_dirty, 0, n_reads, st);
  P.read_dirty = d_dirty;
  if (e == cudaSuccess) {
    if (seed_jit_applies(plan->jit, P) && !getenv("NTHASH_B200_DISABLE_SEED_JIT")) {
      e = launch_seed_jit(plan->jit, P, st);
      if (e == cudaSuccess) e = launch_seed_emit(P, n_reads, st);
    } else {
      e = launch_seed(P, n_reads, st);
    }
  }
  cudaFreeAsync(d_dirty, st);
  NTH_CUDA(e);
  return NTHASH_OK;
}

// One device-resident batch, as the internal runners see it.
struct DevBatch
{
  const uint8_t* d_bases = nullptr;
  uint64_t n_bases = 0;
  const uint64_t* d_read_off = nullptr; // ragged (uniform_len == 0)
  const uint64_t* d_koff = nullptr;
  uint64_t n_reads = 0, max_len = 0;
  uint32_t uniform_len = 0; // > 0: reads of this length back to back from d_bases
  uint64_t* d_out = nullptr;
  uint32_t* d_valid = nullptr;
  uint64_t valid_row0 = 0; // row 0 of this batch is bit valid_row0 of d_valid
  uint64_t memset_rows = 0; // > 0: pre-set that many validity bits first
  const uint64_t* d_rows = nullptr; // ragged *_dev entries: the row count lives on the device (koff[n_reads]) ...
  uint64_t rows_bound = 0;          // ... and only this bound of it is known here
  const RaggedItems* items = nullptr; // ragged layout planned earlier (nthash_ragged_plan)
  uint64_t *d_fwd = nullptr, *d_rev = nullptr;
  uint64_t* d_reduce = nullptr; // fused consumer output {windows, sum, xor}; then d_out etc. are NULL
  uint64_t rows = 0;            // dense rows of this batch (set by the host pipeline)
  uint32_t* d_bloom = nullptr;  // Bloom-filter consumer: filter words, size in bits, 1 = insert / 2 = query
  const uint8_t* d_packed = nullptr; // 2-bit packed input hashed directly (uniform batches the nibble-strip kernel takes); then d_bases is unused
  const uint32_t* d_inv = nullptr;
  uint32_t packed_first = 0;
  uint64_t bloom_bits = 0;
  uint32_t bloom_mode = 0;
};

static int kmer_dev_run(const DevBatch& B, uint32_t k, uint32_t h, cudaStream_t st)
{
  KmerParams P;
  P.bases = B.d_bases;
  P.n_bases = B.n_bases;
  P.k = k;
  P.h = h;
  P.out = B.d_out;
  P.valid_bits = B.d_valid;
  P.valid_row0 = B.valid_row0;
  P.out_fwd = B.d_fwd;
  P.out_rev = B.d_rev;
  P.reduce_out = B.d_reduce;
  P.bloom_words = B.d_bloom;
  P.bloom_bits = B.bloom_bits;
  P.bloom_mode = B.bloom_mode;
  P.packed = B.d_packed;
  P.inv_bits = B.d_inv;
  P.packed_first = B.packed_first;
  RaggedItems R;
  if (B.uniform_len) {
    // the fast kernel plans its own (smaller) CTAs, so a tile too large for the general kernel is not fatal yet
    P.general_fits = plan_uniform(B.n_reads, B.uniform_len, k, P.g, P.tile_cap);
  } else {
    if (!B.items)
      if (int rc = plan_ragged(B.d_read_off, B.d_koff, B.n_reads, B.max_len, B.n_bases, k, TILE_BUDGET, st, R)) return rc;
    const RaggedItems& I = B.items ? *B.items : R;
    P.g = I.g;
    P.tile_cap = I.tile_cap;
    P.tile_cap_256 = I.tile_cap_256;
    P.general_fits = kmer_smem_bytes(P.tile_cap) <= SMEM_MAX;
  }
  int rc = run_kmer(P, B.memset_rows, B.d_rows, B.rows_bound, st);
  R.release_async(st);
  return rc;
}

static int seed_dev_run(const nthash_seed_plan* plan, const DevBatch& B, cudaStream_t st)
{
  SeedParams P;
  fill_seed_params(plan, P);
  P.bases = B.d_bases;
  P.n_bases = B.n_bases;
  P.out = B.d_out;
  P.valid_bits = B.d_valid;
  P.valid_row0 = B.valid_row0;
  P.out_fwd = B.d_fwd;
  P.out_rev = B.d_rev;
  RaggedItems R;
  if (B.uniform_len) {
    if (!plan_uniform(B.n_reads, B.uniform_len, P.k, P.g, P.tile_cap))
      return fail(NTHASH_ERR_UNSUPPORTED, "k=%u with read_len=%u needs a %u-byte tile", P.k, B.uniform_len, P.tile_cap);
  } else {
    P.read_off = B.d_read_off;
    P.koff = B.d_koff;
    const uint32_t budget = TILE_BUDGET > P.plan_smem_bytes / 2 ? TILE_BUDGET - P.plan_smem_bytes / 2 : 0;
    // a layout planned for the k-mer kernels' budget can be taken over when it is the one-item-per-read form and that
    // still fits this (smaller) budget, or when it is already the cut-up form
    const bool reuse = B.items && (B.items->d_items || (uint64_t)KMER_NT * B.max_len + 64 <= budget);
    if (!reuse)
      if (int rc = plan_ragged(B.d_read_off, B.d_koff, B.n_reads, B.max_len, B.n_bases, P.k, budget, st, R, false)) return rc; // (the seed kernels deal their items out themselves)
    const RaggedItems& I = reuse ? *B.items : R;
    P.g = I.g;
    P.tile_cap = I.tile_cap;
    P.item_read = I.item_read;
  }
  int rc = run_seed(plan, P, B.n_reads, B.memset_rows, B.d_rows, B.rows_bound, st);
  R.release_async(st);
  return rc;
}

// ---- host-buffer entry points: chunked H2D / kernel / D2H pipeline over a few streams ----------
struct HostBatch
{
  const char* bases;
  const uint64_t* read_off;
  uint64_t n_reads;
  uint32_t k;
  uint64_t H, strand_cols; // u64 per row in out / in out_fwd,out_rev (0: no strands)
  uint64_t* out;
  uint32_t* valid_bits;
  uint64_t *out_fwd, *out_rev;
  uint64_t* reduce_result = nullptr; // fused consumer: 3 u64 on the host; then out == NULL and nothing else is copied back
  // 2-bit packed input instead of `bases` (then bases == NULL): 4 bases per byte + optional invalid-base bitmap
  const uint8_t* packed = nullptr;
  const uint32_t* invalid_bits = nullptr;
  uint64_t uniform_len = 0; // > 0: n_reads reads of this length back to back, read_off may be NULL (and is not scanned)
  bool scratch_rows = false; // rows and validity bitmap are produced on the device but not copied back (two-pass consumers)
};

template<class Launch>
static int host_pipeline(const HostBatch& hb, Launch&& launch)
{
  constexpr int NS = 3;                         // chunks in flight
  uint64_t CH_VALUES = 48ull << 20;             // ~384 MB of hashes per chunk
  const uint64_t CH_BASES = 256ull << 20;
  if (const char* e = getenv("NTHASH_B200_HOST_CHUNK_VALUES")) CH_VALUES = std::max<uint64_t>(1, strtoull(e, nullptr, 10)); // tests
  const uint64_t n = hb.n_reads;
  const bool timing = getenv("NTHASH_B200_HOST_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  // fixed-length batches (the common case) need no per-read bookkeeping at all
  bool uniform = true;
  const uint64_t len0 = hb.uniform_len ? hb.uniform_len : hb.read_off[1] - hb.read_off[0];
  if (!hb.uniform_len) {
    uint64_t diff = 0; // branch-free so that the compiler vectorises the scan of 10^7 offsets
    for (uint64_t r = 1; r < n; ++r) diff |= (hb.read_off[r + 1] - hb.read_off[r]) ^ len0;
    uniform = diff == 0;
  }
  if (len0 > 0xffffffffull) uniform = false;
  const uint64_t base0 = hb.uniform_len ? 0 : hb.read_off[0];
  const uint64_t nk0 = len0 >= hb.k ? len0 - hb.k + 1 : 0;
  std::vector<uint64_t> koff_v;
  uint64_t rows;
  if (uniform) {
    rows = n * nk0;
  } else {
    koff_v.resize(n + 1);
    rows = nthash_window_rows(hb.read_off, n, hb.k, koff_v.data());
  }
  if (rows == 0) return NTHASH_OK;
  auto koff = [&](uint64_t r) { return uniform ? r * nk0 : koff_v[r]; };

  // chunk boundaries at read granularity
  std::vector<uint64_t> cut(1, 0);
  uint64_t max_reads = 0, max_bases = 0, max_rows = 0;
  if (uniform) {
    const uint64_t per = std::max<uint64_t>(1, std::min(CH_VALUES / std::max<uint64_t>(1, nk0 * hb.H), CH_BASES / std::max<uint64_t>(1, len0)));
    for (uint64_t r = 0; r < n; r += per) cut.push_back(std::min(n, r + per));
    max_reads = std::min(per, n);
    max_bases = max_reads * len0;
    max_rows = max_reads * nk0;
  } else {
    for (uint64_t r = 0; r < n;) {
      uint64_t e = r;
      while (e < n && (e == r || ((koff_v[e + 1] - koff_v[r]) * hb.H <= CH_VALUES && hb.read_off[e + 1] - hb.read_off[r] <= CH_BASES))) ++e;
      cut.push_back(e);
      max_reads = std::max(max_reads, e - r);
      max_bases = std::max(max_bases, hb.read_off[e] - hb.read_off[r]);
      max_rows = std::max(max_rows, koff_v[e] - koff_v[r]);
      r = e;
    }
  }
  const size_t n_chunks = cut.size() - 1;
  const int ns = (int)std::min<size_t>(NS, n_chunks);

  struct Slot
  {
    cudaStream_t st = nullptr;
    uint8_t *d_bases = nullptr, *d_packed = nullptr;
    uint32_t* d_inv = nullptr;
    uint64_t *d_off = nullptr, *d_out = nullptr, *d_fwd = nullptr, *d_rev = nullptr;
    std::vector<uint64_t> h_off;
  } slot[NS];
  uint32_t* d_valid = nullptr;
  uint64_t* d_reduce = nullptr;
  cudaEvent_t ev_valid = nullptr;
  auto cleanup = [&]() {
    for (int i = 0; i < NS; ++i) {
      if (slot[i].st) cudaStreamSynchronize(slot[i].st);
      for (void* q : { (void*)slot[i].d_bases, (void*)slot[i].d_packed, (void*)slot[i].d_inv, (void*)slot[i].d_off, (void*)slot[i].d_out,
                       (void*)slot[i].d_fwd, (void*)slot[i].d_rev })
        if (q) cudaFreeAsync(q, slot[i].st);
      if (i == 0 && d_valid) cudaFreeAsync(d_valid, slot[0].st);
      if (i == 0 && d_reduce) cudaFreeAsync(d_reduce, slot[0].st);
      if (slot[i].st) cudaStreamDestroy(slot[i].st);
    }
    if (ev_valid) cudaEventDestroy(ev_valid);
  };
#define NTH_TRY(expr)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      const int rc__ = fail(NTHASH_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__));       \
      cleanup();                                                                               \
      return rc__;                                                                             \
    }                                                                                          \
  } while (0)
  const uint64_t vwords = (rows + 31) / 32;
  { // staging buffers come from the device's stream-ordered pool, told to keep freed memory for the next call
    int dev = 0;
    cudaMemPool_t pool = nullptr;
    uint64_t keep = ~0ull;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  for (int i = 0; i < ns; ++i) {
    NTH_TRY(cudaStreamCreateWithFlags(&slot[i].st, cudaStreamNonBlocking));
    NTH_TRY(cudaMallocAsync(&slot[i].d_bases, max_bases + 96, slot[i].st));
    if (hb.packed) NTH_TRY(cudaMallocAsync(&slot[i].d_packed, max_bases / 4 + 128, slot[i].st)); // slack: whole 16-byte chunks are read
    if (hb.packed && hb.invalid_bits) NTH_TRY(cudaMallocAsync(&slot[i].d_inv, (max_bases / 32 + 16) * 4, slot[i].st));
    if (!uniform) NTH_TRY(cudaMallocAsync(&slot[i].d_off, 2 * (max_reads + 1) * sizeof(uint64_t), slot[i].st));
    if (hb.out || hb.scratch_rows) NTH_TRY(cudaMallocAsync(&slot[i].d_out, max_rows * hb.H * sizeof(uint64_t), slot[i].st));
    if (hb.strand_cols) {
      NTH_TRY(cudaMallocAsync(&slot[i].d_fwd, max_rows * hb.strand_cols * sizeof(uint64_t), slot[i].st));
      NTH_TRY(cudaMallocAsync(&slot[i].d_rev, max_rows * hb.strand_cols * sizeof(uint64_t), slot[i].st));
    }
  }
  if (hb.reduce_result) {
    NTH_TRY(cudaMallocAsync(&d_reduce, 3 * sizeof(uint64_t), slot[0].st));
    NTH_TRY(cudaMemsetAsync(d_reduce, 0, 3 * sizeof(uint64_t), slot[0].st));
    NTH_TRY(cudaStreamSynchronize(slot[0].st)); // chunks on the other streams accumulate into it
  }
  if (hb.valid_bits || hb.scratch_rows) {
    NTH_TRY(cudaMallocAsync(&d_valid, vwords * 4, slot[0].st));
    NTH_TRY(cudaEventCreateWithFlags(&ev_valid, cudaEventDisableTiming));
    NTH_TRY(cudaMemsetAsync(d_valid, 0xFF, vwords * 4, slot[0].st));
    NTH_TRY(cudaEventRecord(ev_valid, slot[0].st));
    for (int i = 1; i < ns; ++i) NTH_TRY(cudaStreamWaitEvent(slot[i].st, ev_valid, 0));
  }
  const double t1 = now();
  for (size_t c = 0; c < n_chunks; ++c) {
    Slot& s = slot[c % ns];
    if (c >= (size_t)ns) NTH_TRY(cudaStreamSynchronize(s.st)); // the slot's previous chunk has been copied out
    const uint64_t r0 = cut[c], r1 = cut[c + 1], nr = r1 - r0;
    const uint64_t b0 = uniform ? base0 + r0 * len0 : hb.read_off[r0], nbytes = uniform ? (r1 - r0) * len0 : hb.read_off[r1] - b0;
    const uint64_t row0 = koff(r0), nrows = koff(r1) - row0;
    if (nrows == 0) continue;
    bool packed_direct = false;
    uint32_t packed_first = 0;
    // chunk bytes sit 16 bytes into the slot so that "the base before the first one" is addressable
    if (hb.packed) { // a quarter of the bytes cross PCIe; a small kernel expands them to the ASCII the hash kernels read
      // both slices start at the bitmap word holding base b0 (base 32*v0), so one offset addresses them
      const uint64_t v0 = b0 / 32, v1 = (b0 + nbytes + 31) / 32, p0 = 8 * v0, p1 = (b0 + nbytes + 3) / 4;
      NTH_TRY(cudaMemcpyAsync(s.d_packed, hb.packed + p0, p1 - p0, cudaMemcpyHostToDevice, s.st));
      if (hb.invalid_bits) NTH_TRY(cudaMemcpyAsync(s.d_inv, hb.invalid_bits + v0, (v1 - v0) * 4, cudaMemcpyHostToDevice, s.st));
      // the kernels index the staged slices with bases counted from their first byte / word.  Fixed-length reads the
      // nibble-strip kernel takes are hashed straight from the packed bytes; everything else is expanded to ASCII first
      packed_direct = uniform && !hb.reduce_result && !hb.out_fwd && kmer_packed_direct_ok(nr, (uint32_t)len0, hb.k, (uint32_t)hb.H);
      if (!packed_direct)
        NTH_TRY(launch_unpack2bit(s.d_packed, hb.invalid_bits ? s.d_inv : nullptr, b0 - 32 * v0, nbytes, s.d_bases + 16, s.st));
      packed_first = (uint32_t)(b0 - 32 * v0);
    } else {
      NTH_TRY(cudaMemcpyAsync(s.d_bases + 16, hb.bases + b0, nbytes, cudaMemcpyHostToDevice, s.st));
    }
    DevBatch B;
    B.n_reads = nr;
    B.d_out = s.d_out;
    B.d_valid = d_valid;
    B.valid_row0 = row0;
    B.d_fwd = s.d_fwd;
    B.d_rev = s.d_rev;
    B.d_reduce = d_reduce;
    B.rows = nrows;
    if (uniform) {
      B.d_bases = s.d_bases + 16;
      B.n_bases = nbytes + 64;
      B.uniform_len = (uint32_t)len0;
      if (packed_direct) {
        B.d_packed = s.d_packed;
        B.d_inv = hb.invalid_bits ? s.d_inv : nullptr;
        B.packed_first = packed_first;
      }
    } else {
      s.h_off.resize(2 * (nr + 1));
      uint64_t mx = 0;
      for (uint64_t i = 0; i <= nr; ++i) {
        s.h_off[i] = hb.read_off[r0 + i] - b0 + 16;
        s.h_off[nr + 1 + i] = koff_v[r0 + i] - row0;
        if (i < nr) mx = std::max(mx, hb.read_off[r0 + i + 1] - hb.read_off[r0 + i]);
      }
      NTH_TRY(cudaMemcpyAsync(s.d_off, s.h_off.data(), 2 * (nr + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s.st));
      B.d_bases = s.d_bases;
      B.n_bases = 16 + nbytes + 64;
      B.d_read_off = s.d_off;
      B.d_koff = s.d_off + nr + 1;
      B.max_len = mx;
    }
    if (const int rc = launch(B, s.st)) {
      cleanup();
      return rc;
    }
    if (hb.out) NTH_TRY(cudaMemcpyAsync(hb.out + row0 * hb.H, s.d_out, nrows * hb.H * sizeof(uint64_t), cudaMemcpyDeviceToHost, s.st));
    if (hb.strand_cols) {
      NTH_TRY(cudaMemcpyAsync(hb.out_fwd + row0 * hb.strand_cols, s.d_fwd, nrows * hb.strand_cols * sizeof(uint64_t), cudaMemcpyDeviceToHost, s.st));
      NTH_TRY(cudaMemcpyAsync(hb.out_rev + row0 * hb.strand_cols, s.d_rev, nrows * hb.strand_cols * sizeof(uint64_t), cudaMemcpyDeviceToHost, s.st));
    }
  }
  const double t2 = now();
  for (int i = 0; i < ns; ++i) NTH_TRY(cudaStreamSynchronize(slot[i].st));
  if (hb.valid_bits) NTH_TRY(cudaMemcpy(hb.valid_bits, d_valid, vwords * 4, cudaMemcpyDeviceToHost));
  if (hb.reduce_result) NTH_TRY(cudaMemcpy(hb.reduce_result, d_reduce, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  const double t3 = now();
  cleanup();
  if (timing)
    fprintf(stderr, "[nthash_b200 host pipeline] %zu chunks: plan+alloc %.1f ms, enqueue %.1f ms, drain %.1f ms, free %.1f ms\n", n_chunks,
            t1 - t0, t2 - t1, t3 - t2, now() - t3);
#undef NTH_TRY
  return NTHASH_OK;
}

} // namespace nthb

using namespace nthb;

extern "C" {

const char* nthash_fn_name(void) { return "ntHash_v2"; }
const char* nthash_last_error(void) { return g_err.c_str(); }
int nthash_b200_abi_version(void) { return 2; } // 2: consumers, packed input, FASTQ staging, compaction, multi-GPU entry

int nthash_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  int ok = 0;
  for (int d = 0; d < n; ++d) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
  }
  return ok;
}

uint64_t nthash_window_rows(const uint64_t* read_off, uint64_t n_reads, uint32_t k, uint64_t* koff)
{
  uint64_t acc = 0;
  for (uint64_t r = 0; r < n_reads; ++r) {
    if (koff) koff[r] = acc;
    const uint64_t len = read_off[r + 1] - read_off[r];
    if (len >= k) acc += len - k + 1;
  }
  if (koff) koff[n_reads] = acc;
  return acc;
}

uint64_t nthash_valid_words(uint64_t rows) { return (rows + 31) / 32; }

int nthash_kmer_batch_uniform_dev(const uint8_t* d_bases, uint64_t n_bases_readable, uint64_t n_reads,
                                  uint32_t read_len, uint32_t k, uint32_t num_hashes, uint64_t* d_out,
                                  uint32_t* d_valid_bits, uint64_t* d_out_fwd, uint64_t* d_out_rev, void* stream)
{
  if (int rc = check_kh(k, num_hashes)) return rc;
  if (n_reads == 0 || read_len < k) return NTHASH_OK; // no windows anywhere
  if (int rc = check_outputs(d_out, d_out_fwd, d_out_rev)) return rc;
  if (!d_bases || ((uintptr_t)d_bases & 15)) return fail(NTHASH_ERR_INVALID_ARG, "d_bases must be 16-byte aligned");
  if (n_bases_readable < n_reads * (uint64_t)read_len)
    return fail(NTHASH_ERR_INVALID_ARG, "n_bases_readable smaller than n_reads*read_len");
  if (int rc = check_device_ready()) return rc;
  DevBatch B;
  B.d_bases = d_bases;
  B.n_bases = n_bases_readable;
  B.n_reads = n_reads;
  B.uniform_len = read_len;
  B.d_out = d_out;
  B.d_valid = d_valid_bits;
  B.memset_rows = n_reads * (uint64_t)(read_len - k + 1);
  B.d_fwd = d_out_fwd;
  B.d_rev = d_out_rev;
  return kmer_dev_run(B, k, num_hashes, (cudaStream_t)stream);
}

int nthash_kmer_plan_dev(const uint64_t* d_read_off, uint64_t n_reads, uint32_t k, uint64_t* d_koff,
                         uint64_t* total_rows, uint64_t* max_read_len, void* stream)
{
  if (int rc = check_kh(k, 1)) return rc;
  if (!d_read_off || !d_koff) return fail(NTHASH_ERR_INVALID_ARG, "d_read_off and d_koff must not be NULL");
  if (int rc = check_device_ready()) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  uint64_t* d_stats = nullptr;
  NTH_CUDA(cudaMallocAsync(&d_stats, 2 * sizeof(uint64_t), st));
  cudaError_t e = launch_koff_scan(d_read_off, n_reads, k, 0, d_koff, d_stats, st);
  uint64_t h_stats[2] = { 0, 0 };
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_stats, d_stats, sizeof h_stats, cudaMemcpyDeviceToHost, st);
  cudaFreeAsync(d_stats, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  NTH_CUDA(e);
  if (total_rows) *total_rows = h_stats[0];
  if (max_read_len) *max_read_len = h_stats[1];
  return NTHASH_OK;
}

int nthash_kmer_batch_dev(const uint8_t* d_bases, uint64_t n_bases_readable, const uint64_t* d_read_off,
                          const uint64_t* d_koff, uint64_t n_reads, uint64_t max_read_len, uint32_t k,
                          uint32_t num_hashes, uint64_t* d_out, uint32_t* d_valid_bits, uint64_t* d_out_fwd,
                          uint64_t* d_out_rev, void* stream)
{
  if (int rc = check_kh(k, num_hashes)) return rc;
  if (n_reads == 0 || max_read_len < k) return NTHASH_OK;
  if (int rc = check_outputs(d_out, d_out_fwd, d_out_rev)) return rc;
  if (!d_bases || ((uintptr_t)d_bases & 15)) return fail(NTHASH_ERR_INVALID_ARG, "d_bases must be 16-byte aligned");
  if (!d_read_off || !d_koff) return fail(NTHASH_ERR_INVALID_ARG, "d_read_off and d_koff must not be NULL");
  if (int rc = check_device_ready()) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  DevBatch B;
  B.d_rows = d_koff + n_reads; // the row total stays on the device: the bitmap is pre-set by a kernel that reads it there
  B.rows_bound = n_bases_readable;
  B.d_bases = d_bases;
  B.n_bases = n_bases_readable;
  B.d_read_off = d_read_off;
  B.d_koff = d_koff;
  B.n_reads = n_reads;
  B.max_len = max_read_len;
  B.d_out = d_out;
  B.d_valid = d_valid_bits;
  B.d_fwd = d_out_fwd;
  B.d_rev = d_out_rev;
  return kmer_dev_run(B, k, num_hashes, st);
}

// ---- ragged layout planned once (nthash_ragged_plan): steady-state calls neither re-plan nor synchronise ----------

int nthash_ragged_plan_create(const uint64_t* d_read_off, uint64_t n_reads, uint32_t k, void* stream, nthash_ragged_plan** plan_out)
{
  if (!plan_out) return fail(NTHASH_ERR_INVALID_ARG, "plan_out must not be NULL");
  *plan_out = nullptr;
  if (int rc = check_kh(k, 1)) return rc;
  if (!d_read_off) return fail(NTHASH_ERR_INVALID_ARG, "d_read_off must not be NULL");
  if (int rc = check_device_ready()) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  nthash_ragged_plan* pl = new nthash_ragged_plan();
  cudaGetDevice(&pl->device);
  pl->k = k;
  pl->n_reads = n_reads;
  pl->d_read_off = d_read_off;
  uint64_t* d_stats = nullptr;
  uint64_t h_stats[3] = { 0, 0, 0 }, ends[2] = { 0, 0 };
  cudaError_t e = cudaMalloc(&pl->d_koff, (n_reads + 1) * sizeof(uint64_t));
  if (e == cudaSuccess) e = cudaMallocAsync(&d_stats, 3 * sizeof(uint64_t), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_stats, 0, 3 * sizeof(uint64_t), st);
  if (e == cudaSuccess) e = launch_koff_scan(d_read_off, n_reads, k, 0, pl->d_koff, d_stats, st);
  if (e == cudaSuccess) e = launch_block_span_max(d_read_off, n_reads, d_stats + 2, st); // longest 256-read block, in bytes
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_stats, d_stats, sizeof h_stats, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&ends[0], d_read_off, sizeof(uint64_t), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&ends[1], d_read_off + n_reads, sizeof(uint64_t), cudaMemcpyDeviceToHost, st);
  if (d_stats) cudaFreeAsync(d_stats, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st); // the one synchronisation of a plan's life
  int rc = NTHASH_OK;
  if (e == cudaSuccess) {
    pl->rows = h_stats[0];
    pl->max_len = h_stats[1];
    pl->n_bases = ends[1];
    if (pl->rows) {
      rc = plan_ragged(d_read_off, pl->d_koff, n_reads, pl->max_len, ends[1] - ends[0], k, TILE_BUDGET, st, pl->items);
      // reads are the items: the fast kernel's CTAs of 256 threads stage exactly these blocks, so the measured span bounds them
      if (rc == NTHASH_OK && !pl->items.d_items && h_stats[2] + 64 < pl->items.tile_cap && !getenv("NTHASH_B200_NO_EXACT_TILE"))
        pl->items.tile_cap_256 = (uint32_t)h_stats[2] + 64;
      const uint4* t4 = nullptr; // the 4 KB warm-up table of this k is created on first use: do it now, not inside a graph capture
      if (rc == NTHASH_OK) e = get_t4_table(k, &t4);
      if (rc == NTHASH_OK && e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
  }
  if (e != cudaSuccess && rc == NTHASH_OK) rc = fail(NTHASH_ERR_CUDA, "ragged plan: %s", cudaGetErrorString(e));
  if (rc != NTHASH_OK) {
    nthash_ragged_plan_destroy(pl);
    return rc;
  }
  *plan_out = pl;
  return NTHASH_OK;
}

void nthash_ragged_plan_destroy(nthash_ragged_plan* plan)
{
  if (!plan) return;
  if (plan->items.d_items) cudaFree(plan->items.d_items);
  if (plan->items.d_perm) cudaFree(plan->items.d_perm);
  cudaFree(plan->d_koff);
  delete plan;
}

uint64_t nthash_ragged_plan_rows(const nthash_ragged_plan* plan) { return plan ? plan->rows : 0; }
uint64_t nthash_ragged_plan_max_read_len(const nthash_ragged_plan* plan) { return plan ? plan->max_len : 0; }
const uint64_t* nthash_ragged_plan_koff(const nthash_ragged_plan* plan) { return plan ? plan->d_koff : nullptr; }

static int planned_batch(const nthash_ragged_plan* plan, const uint8_t* d_bases, uint64_t n_bases_readable, DevBatch& B)
{
  if (!plan) return fail(NTHASH_ERR_INVALID_ARG, "plan must not be NULL");
  if (!d_bases || ((uintptr_t)d_bases & 15)) return fail(NTHASH_ERR_INVALID_ARG, "d_bases must be 16-byte aligned");
  if (n_bases_readable < plan->n_bases) return fail(NTHASH_ERR_INVALID_ARG, "n_bases_readable smaller than the plan's last read end");
  B.d_bases = d_bases;
  B.n_bases = n_bases_readable;
  B.d_read_off = plan->d_read_off;
  B.d_koff = plan->d_koff;
  B.n_reads = plan->n_reads;
  B.max_len = plan->max_len;
  B.memset_rows = plan->rows;
  B.items = &plan->items;
  return NTHASH_OK;
}

int nthash_kmer_batch_planned_dev(const nthash_ragged_plan* plan, const uint8_t* d_bases, uint64_t n_bases_readable, uint32_t num_hashes,
                                  uint64_t* d_out, uint32_t* d_valid_bits, uint64_t* d_out_fwd, uint64_t* d_out_rev, void* stream)
{
  if (!plan) return fail(NTHASH_ERR_INVALID_ARG, "plan must not be NULL");
  if (int rc = check_kh(plan->k, num_hashes)) return rc;
  if (plan->rows == 0) return NTHASH_OK;
  if (int rc = check_outputs(d_out, d_out_fwd, d_out_rev)) return rc;
  DevBatch B;
  if (int rc = planned_batch(plan, d_bases, n_bases_readable, B)) return rc;
  B.d_out = d_out;
  B.d_valid = d_valid_bits;
  B.d_fwd = d_out_fwd;
  B.d_rev = d_out_rev;
  return kmer_dev_run(B, plan->k, num_hashes, (cudaStream_t)stream);
}

int nthash_kmer_reduce_planned_dev(const nthash_ragged_plan* plan, const uint8_t* d_bases, uint64_t n_bases_readable, uint32_t num_hashes,
                                   uint64_t* d_result, void* stream)
{
  if (!plan) return fail(NTHASH_ERR_INVALID_ARG, "plan must not be NULL");
  if (int rc = check_kh(plan->k, num_hashes)) return rc;
  if (!d_result) return fail(NTHASH_ERR_INVALID_ARG, "d_result must not be NULL");
  NTH_CUDA(cudaMemsetAsync(d_result, 0, 3 * sizeof(uint64_t), (cudaStream_t)stream));
  if (plan->rows == 0) return NTHASH_OK;
  DevBatch B;
  if (int rc = planned_batch(plan, d_bases, n_bases_readable, B)) return rc;
  B.memset_rows = 0;
  B.d_reduce = d_result;
  return kmer_dev_run(B, plan->k, num_hashes, (cudaStream_t)stream);
}

int nthash_seed_batch_planned_dev(const nthash_seed_plan* seeds, const nthash_ragged_plan* plan, const uint8_t* d_bases,
                                  uint64_t n_bases_readable, uint64_t* d_out, uint32_t* d_valid_bits, uint64_t* d_out_fwd,
                                  uint64_t* d_out_rev, void* stream)
{
  if (!seeds || !plan) return fail(NTHASH_ERR_INVALID_ARG, "the seed plan and the ragged plan must not be NULL");
  if (seeds->host.k != plan->k) return fail(NTHASH_ERR_INVALID_ARG, "the ragged plan was made for k=%u, the seeds have length %u", plan->k, seeds->host.k);
  if (plan->rows == 0) return NTHASH_OK;
  if (int rc = check_outputs(d_out, d_out_fwd, d_out_rev)) return rc;
  DevBatch B;
  if (int rc = planned_batch(plan, d_bases, n_bases_readable, B)) return rc;
  B.d_out = d_out;
  B.d_valid = d_valid_bits;
  B.d_fwd = d_out_fwd;
  B.d_rev = d_out_rev;
  return seed_dev_run(seeds, B, (cudaStream_t)stream);
}

int nthash_kmer_batch(const char* bases, const uint64_t* read_off, uint64_t n_reads, uint32_t k,
                      uint32_t num_hashes, uint64_t* out, uint32_t* valid_bits, uint64_t* out_fwd,
                      uint64_t* out_rev, int device)
{
  if (int rc = check_kh(k, num_hashes)) return rc;
  if (n_reads == 0) return NTHASH_OK;
  if (!bases || !read_off) return fail(NTHASH_ERR_INVALID_ARG, "bases and read_off must not be NULL");
  if (int rc = check_outputs(out, out_fwd, out_rev)) return rc;
  if (cudaSetDevice(device) != cudaSuccess) return fail(NTHASH_ERR_NO_DEVICE, "cannot select CUDA device %d", device);
  if (int rc = check_device_ready()) return rc;
  const HostBatch hb = { bases, read_off, n_reads, k, num_hashes, out_fwd ? 1ull : 0ull, out, valid_bits, out_fwd, out_rev };
  return host_pipeline(hb, [&](const DevBatch& B, cudaStream_t st) { return kmer_dev_run(B, k, num_hashes, st); });
}


// One host batch over several GPUs of the box: contiguous read ranges with about equal numbers of bases, one host
// thread + the chunked pipeline per device, no exchange between devices.
int nthash_kmer_batch_multi(const char* bases, const uint64_t* read_off, uint64_t n_reads, uint32_t k, uint32_t num_hashes,
                            uint64_t* out, uint32_t* valid_bits, uint64_t* out_fwd, uint64_t* out_rev, const int* devices,
                            int n_devices)
{
  if (int rc = check_kh(k, num_hashes)) return rc;
  if (n_devices < 1 || !devices) return fail(NTHASH_ERR_INVALID_ARG, "give at least one device");
  if (n_reads == 0) return NTHASH_OK;
  if (!bases || !read_off) return fail(NTHASH_ERR_INVALID_ARG, "bases and read_off must not be NULL");
  if (int rc = check_outputs(out, out_fwd, out_rev)) return rc;
  const uint64_t G = std::min<uint64_t>((uint64_t)n_devices, n_reads);
  // shard boundaries: the first read starting at or after j/G of the bases
  std::vector<uint64_t> cut(G + 1, n_reads);
  cut[0] = 0;
  const uint64_t total = read_off[n_reads] - read_off[0];
  for (uint64_t j = 1; j < G; ++j)
    cut[j] = std::max<uint64_t>(cut[j - 1], std::lower_bound(read_off, read_off + n_reads, read_off[0] + total / G * j) - read_off);
  // dense row of every boundary (the shards' outputs are slices of the caller's arrays)
  std::vector<uint64_t> row(G + 1, 0);
  for (uint64_t j = 0; j < G; ++j) row[j + 1] = row[j] + nthash_window_rows(read_off + cut[j], cut[j + 1] - cut[j], k, nullptr);
  std::vector<int> rcs(G, NTHASH_OK);
  std::vector<std::string> errs(G);
  std::vector<std::vector<uint32_t>> vtmp(G); // per-shard bitmaps start at bit 0; merged below
  std::vector<std::thread> threads;
  for (uint64_t j = 0; j < G; ++j) {
    threads.emplace_back([&, j]() {
      const uint64_t nr = cut[j + 1] - cut[j], rows = row[j + 1] - row[j];
      if (nr == 0 || rows == 0) return;
      if (valid_bits) vtmp[j].assign((rows + 31) / 32, 0u);
      rcs[j] = nthash_kmer_batch(bases, read_off + cut[j], nr, k, num_hashes, out + row[j] * num_hashes,
                                 valid_bits ? vtmp[j].data() : nullptr, out_fwd ? out_fwd + row[j] : nullptr,
                                 out_rev ? out_rev + row[j] : nullptr, devices[j]);
      if (rcs[j] != NTHASH_OK) errs[j] = g_err; // the message lives in this thread's slot
    });
  }
  for (std::thread& t : threads) t.join();
  for (uint64_t j = 0; j < G; ++j)
    if (rcs[j] != NTHASH_OK) return fail(rcs[j], "device %d: %s", devices[j], errs[j].c_str());
  if (valid_bits) {
    std::fill(valid_bits, valid_bits + (row[G] + 31) / 32, 0u);
    for (uint64_t j = 0; j < G; ++j) {
      const uint64_t rows = row[j + 1] - row[j], b0 = row[j];
      const uint32_t sh = (uint32_t)(b0 & 31);
      for (uint64_t w = 0; w < vtmp[j].size(); ++w) {
        uint32_t v = vtmp[j][w];
        if (w + 1 == vtmp[j].size() && (rows & 31)) v &= (1u << (rows & 31)) - 1u; // bits past the shard's last row
        valid_bits[(b0 >> 5) + w] |= v << sh;
        if (sh && ((b0 >> 5) + w + 1) < (row[G] + 31) / 32) valid_bits[(b0 >> 5) + w + 1] |= v >> (32 - sh);
      }
    }
  }
  return NTHASH_OK;
}

// Fixed-length batches from host memory: no offsets array to build, scan or keep (10^7 reads = 80 MB of offsets).
int nthash_kmer_batch_uniform(const char* bases, uint64_t n_reads, uint32_t read_len, uint32_t k, uint32_t num_hashes, uint64_t* out,
                              uint32_t* valid_bits, uint64_t* out_fwd, uint64_t* out_rev, int device)
{
  if (int rc = check_kh(k, num_hashes)) return rc;
  if (int rc = check_outputs(out, out_fwd, out_rev)) return rc;
  if (n_reads == 0 || read_len < k) return NTHASH_OK;
  if (!bases) return fail(NTHASH_ERR_INVALID_ARG, "bases must not be NULL");
  if (cudaSetDevice(device) != cudaSuccess) return fail(NTHASH_ERR_NO_DEVICE, "cannot select CUDA device %d", device);
  if (int rc = check_device_ready()) return rc;
  HostBatch hb = { bases, nullptr, n_reads, k, num_hashes, out_fwd ? 1ull : 0ull, out, valid_bits, out_fwd, out_rev };
  hb.uniform_len = read_len;
  return host_pipeline(hb, [&](const DevBatch& B, cudaStream_t st) { return kmer_dev_run(B, k, num_hashes, st); });
}

// ---- fused consumer: count / sum / xor of every hash value ------------------------------------

int nthash_kmer_reduce_uniform_dev(const uint8_t* d_bases, uint64_t n_bases_readable, uint64_t n_reads,
                                   uint32_t read_len, uint32_t k, uint32_t num_hashes, uint64_t* d_result, void* stream)
{
  if (int rc = check_kh(k, num_hashes)) return rc;
  if (!d_result) return fail(NTHASH_ERR_INVALID_ARG, "d_result must not be NULL");
  if (int rc = check_device_ready()) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  NTH_CUDA(cudaMemsetAsync(d_result, 0, 3 * sizeof(uint64_t), st));
  if (n_reads == 0 || read_len < k) return NTHASH_OK;
  if (!d_bases || ((uintptr_t)d_bases & 15)) return fail(NTHASH_ERR_INVALID_ARG, "d_bases must be 16-byte aligned");
  if (n_bases_readable < n_reads * (uint64_t)read_len)
    return fail(NTHASH_ERR_INVALID_ARG, "n_bases_readable smaller than n_reads*read_len");
  DevBatch B;
  B.d_bases = d_bases;
  B.n_bases = n_bases_readable;
  B.n_reads = n_reads;
  B.uniform_len = read_len;
  B.d_reduce = d_result;
  return kmer_dev_run(B, k, num_hashes, st);
}

int nthash_kmer_reduce_dev(const uint8_t* d_bases, uint64_t n_bases_readable, const uint64_t* d_read_off,
                           const uint64_t* d_koff, uint64_t n_reads, uint64_t max_read_len, uint32_t k,
                           uint32_t num_hashes, uint64_t* d_result, void* stream)
{
  if (int rc = check_kh(k, num_hashes)) return rc;
  if (!d_result) return fail(NTHASH_ERR_INVALID_ARG, "d_result must not be NULL");
  if (int rc = check_device_ready()) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  NTH_CUDA(cudaMemsetAsync(d_result, 0, 3 * sizeof(uint64_t), st));
  if (n_reads == 0 || max_read_len < k) return NTHASH_OK;
  if (!d_bases || ((uintptr_t)d_bases & 15)) return fail(NTHASH_ERR_INVALID_ARG, "d_bases must be 16-byte aligned");
  if (!d_read_off || !d_koff) return fail(NTHASH_ERR_INVALID_ARG, "d_read_off and d_koff must not be NULL");
  DevBatch B;
  B.d_bases = d_bases;
  B.n_bases = n_bases_readable;
  B.d_read_off = d_read_off;
  B.d_koff = d_koff;
  B.n_reads = n_reads;
  B.max_len = max_read_len;
  B.d_reduce = d_result;
  return kmer_dev_run(B, k, num_hashes, st);
}

int nthash_kmer_reduce(const char* bases, const uint64_t* read_off, uint64_t n_reads, uint32_t k, uint32_t num_hashes,
                       uint64_t* result, int device)
{
  if (int rc = check_kh(k, num_hashes)) return rc;
  if (!result) return fail(NTHASH_ERR_INVALID_ARG, "result must not be NULL");
  result[0] = result[1] = result[2] = 0;
  if (n_reads == 0) return NTHASH_OK;
  if (!bases || !read_off) return fail(NTHASH_ERR_INVALID_ARG, "bases and read_off must not be NULL");
  if (cudaSetDevice(device) != cudaSuccess) return fail(NTHASH_ERR_NO_DEVICE, "cannot select CUDA device %d", device);
  if (int rc = check_device_ready()) return rc;
  HostBatch hb = { bases, read_off, n_reads, k, num_hashes, 0ull, nullptr, nullptr, nullptr, nullptr };
  hb.reduce_result = result;
  return host_pipeline(hb, [&](const DevBatch& B, cudaStream_t st) { return kmer_dev_run(B, k, num_hashes, st); });
}

// ---- 2-bit packed input (host entries): a quarter of the bytes cross PCIe --------------------------

static int packed_args(const uint8_t* packed, const uint64_t* read_off, uint64_t n_reads, uint32_t uniform_read_len, int device)
{
  if (n_reads == 0) return NTHASH_OK;
  if (!packed) return fail(NTHASH_ERR_INVALID_ARG, "packed must not be NULL");
  if (!read_off && !uniform_read_len) return fail(NTHASH_ERR_INVALID_ARG, "give read_off or a uniform_read_len > 0");
  if (cudaSetDevice(device) != cudaSuccess) return fail(NTHASH_ERR_NO_DEVICE, "cannot select CUDA device %d", device);
  return check_device_ready();
}

int nthash_kmer_batch_packed2bit(const uint8_t* packed, const uint32_t* invalid_bits, const uint64_t* read_off, uint64_t n_reads,
                                 uint32_t uniform_read_len, uint32_t k, uint32_t num_hashes, uint64_t* out, uint32_t* valid_bits,
                                 int device)
{
  if (int rc = check_kh(k, num_hashes)) return rc;
  if (int rc = check_outputs(out, nullptr, nullptr)) return rc;
  if (int rc = packed_args(packed, read_off, n_reads, uniform_read_len, device)) return rc;
  if (n_reads == 0) return NTHASH_OK;
  HostBatch hb = { nullptr, read_off, n_reads, k, num_hashes, 0ull, out, valid_bits, nullptr, nullptr };
  hb.uniform_len = read_off ? 0 : uniform_read_len;
  hb.packed = packed;
  hb.invalid_bits = invalid_bits;
  return host_pipeline(hb, [&](const DevBatch& B, cudaStream_t st) { return kmer_dev_run(B, k, num_hashes, st); });
}

int nthash_kmer_reduce_packed2bit(const uint8_t* packed, const uint32_t* invalid_bits, const uint64_t* read_off, uint64_t n_reads,
                                  uint32_t uniform_read_len, uint32_t k, uint32_t num_hashes, uint64_t* result, int device)
{
  if (int rc = check_kh(k, num_hashes)) return rc;
  if (!result) return fail(NTHASH_ERR_INVALID_ARG, "result must not be NULL");
  result[0] = result[1] = result[2] = 0;
  if (int rc = packed_args(packed, read_off, n_reads, uniform_read_len, device)) return rc;
  if (n_reads == 0) return NTHASH_OK;
  HostBatch hb = { nullptr, read_off, n_reads, k, num_hashes, 0ull, nullptr, nullptr, nullptr, nullptr };
  hb.uniform_len = read_off ? 0 : uniform_read_len;
  hb.reduce_result = result;
  hb.packed = packed;
  hb.invalid_bits = invalid_bits;
  return host_pipeline(hb, [&](const DevBatch& B, cudaStream_t st) { return kmer_dev_run(B, k, num_hashes, st); });
}

int nthash_kmer_batch_packed2bit_uniform_dev(const uint8_t* d_packed, const uint32_t* d_invalid_bits, uint64_t first_base,
                                             uint64_t n_reads, uint32_t read_len, uint32_t k, uint32_t num_hashes,
                                             uint64_t* d_out, uint32_t* d_valid_bits, void* stream)
{
  if (int rc = check_kh(k, num_hashes)) return rc;
  if (n_reads == 0 || read_len < k) return NTHASH_OK;
  if (int rc = check_outputs(d_out, nullptr, nullptr)) return rc;
  if (!d_packed || ((uintptr_t)d_packed & 15)) return fail(NTHASH_ERR_INVALID_ARG, "d_packed must be 16-byte aligned");
  if ((uintptr_t)d_invalid_bits & 7) return fail(NTHASH_ERR_INVALID_ARG, "d_invalid_bits must be 8-byte aligned");
  if (first_base > 0xffffffffull) return fail(NTHASH_ERR_INVALID_ARG, "first_base must be below 2^32");
  if (!kmer_packed_direct_ok(n_reads, read_len, k, num_hashes))
    return fail(NTHASH_ERR_UNSUPPORTED, "this shape is not hashed from packed bytes directly: expand with nthash_unpack2bit_dev");
  if (int rc = check_device_ready()) return rc;
  DevBatch B;
  B.d_packed = d_packed;
  B.d_inv = d_invalid_bits;
  B.packed_first = (uint32_t)first_base;
  B.n_reads = n_reads;
  B.uniform_len = read_len;
  B.d_out = d_out;
  B.d_valid = d_valid_bits;
  B.memset_rows = n_reads * (uint64_t)(read_len - k + 1);
  return kmer_dev_run(B, k, num_hashes, (cudaStream_t)stream);
}

int nthash_unpack2bit_dev(const uint8_t* d_packed, const uint32_t* d_invalid_bits, uint64_t first_base, uint64_t n_bases,
                          uint8_t* d_bases_out, void* stream)
{
  if (n_bases == 0) return NTHASH_OK;
  if (!d_packed || !d_bases_out) return fail(NTHASH_ERR_INVALID_ARG, "d_packed and d_bases_out must not be NULL");
  if ((uintptr_t)d_bases_out & 15) return fail(NTHASH_ERR_INVALID_ARG, "d_bases_out must be 16-byte aligned");
  if (int rc = check_device_ready()) return rc;
  NTH_CUDA(launch_unpack2bit(d_packed, d_invalid_bits, first_base, n_bases, d_bases_out, (cudaStream_t)stream));
  return NTHASH_OK;
}

// ---- FASTQ staging on the device --------------------------------------------------------------------

int nthash_fastq_extract_dev(const uint8_t* d_text, uint64_t n_bytes, uint8_t* d_bases, uint64_t bases_capacity,
                             uint64_t* d_read_off, uint64_t reads_capacity, uint64_t* n_reads, uint64_t* n_bases, void* stream)
{
  if (!n_reads || !n_bases) return fail(NTHASH_ERR_INVALID_ARG, "n_reads and n_bases must not be NULL");
  *n_reads = *n_bases = 0;
  if (!d_read_off) return fail(NTHASH_ERR_INVALID_ARG, "d_read_off must not be NULL");
  if (n_bytes && (!d_text || !d_bases)) return fail(NTHASH_ERR_INVALID_ARG, "d_text and d_bases must not be NULL");
  if (int rc = check_device_ready()) return rc;
  const cudaError_t e = fastq_extract(d_text, n_bytes, d_bases, bases_capacity, d_read_off, reads_capacity, n_reads, n_bases, (cudaStream_t)stream);
  if (e == cudaErrorInvalidValue) {
    cudaGetLastError();
    return fail(NTHASH_ERR_INVALID_ARG, "the text holds more reads or bases than the output buffers' capacities");
  }
  NTH_CUDA(e);
  return NTHASH_OK;
}

// ---- compacted output: only the windows the reference's loop visits, in its order ------------------

int nthash_compact_rows_dev(const uint64_t* d_out, const uint32_t* d_valid_bits, uint64_t rows, uint32_t values_per_row,
                            uint64_t* d_compact, uint64_t* d_row_index, uint64_t* d_count, void* stream)
{
  if (!d_count) return fail(NTHASH_ERR_INVALID_ARG, "d_count must not be NULL");
  if (rows && (!d_out || !d_valid_bits || !d_compact || !values_per_row))
    return fail(NTHASH_ERR_INVALID_ARG, "d_out, d_valid_bits and d_compact must not be NULL, values_per_row > 0");
  if (int rc = check_device_ready()) return rc;
  NTH_CUDA(launch_compact_rows(d_out, d_valid_bits, rows, values_per_row, d_compact, d_row_index, d_count, (cudaStream_t)stream));
  return NTHASH_OK;
}

// ---- fused consumer: Bloom filter insert / query ------------------------------------------------

static int bloom_args(uint32_t k, uint32_t h, const uint32_t* d_filter, uint64_t bits, const uint64_t* d_result)
{
  if (int rc = check_kh(k, h)) return rc;
  if (!d_filter || bits == 0) return fail(NTHASH_ERR_INVALID_ARG, "the filter must not be NULL or empty");
  if ((uintptr_t)d_filter & 3) return fail(NTHASH_ERR_INVALID_ARG, "d_filter_words must be 4-byte aligned");
  if (!d_result) return fail(NTHASH_ERR_INVALID_ARG, "d_result must not be NULL");
  return check_device_ready();
}

int nthash_kmer_bloom_uniform_dev(const uint8_t* d_bases, uint64_t n_bases_readable, uint64_t n_reads, uint32_t read_len,
                                  uint32_t k, uint32_t num_hashes, uint32_t* d_filter_words, uint64_t filter_bits,
                                  int query, uint64_t* d_result, void* stream)
{
  if (int rc = bloom_args(k, num_hashes, d_filter_words, filter_bits, d_result)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  NTH_CUDA(cudaMemsetAsync(d_result, 0, 3 * sizeof(uint64_t), st));
  if (n_reads == 0 || read_len < k) return NTHASH_OK;
  if (!d_bases || ((uintptr_t)d_bases & 15)) return fail(NTHASH_ERR_INVALID_ARG, "d_bases must be 16-byte aligned");
  if (n_bases_readable < n_reads * (uint64_t)read_len)
    return fail(NTHASH_ERR_INVALID_ARG, "n_bases_readable smaller than n_reads*read_len");
  DevBatch B;
  B.d_bases = d_bases;
  B.n_bases = n_bases_readable;
  B.n_reads = n_reads;
  B.uniform_len = read_len;
  B.d_reduce = d_result;
  B.d_bloom = d_filter_words;
  B.bloom_bits = filter_bits;
  B.bloom_mode = query ? 2 : 1;
  return kmer_dev_run(B, k, num_hashes, st);
}

int nthash_kmer_bloom_dev(const uint8_t* d_bases, uint64_t n_bases_readable, const uint64_t* d_read_off,
                          const uint64_t* d_koff, uint64_t n_reads, uint64_t max_read_len, uint32_t k,
                          uint32_t num_hashes, uint32_t* d_filter_words, uint64_t filter_bits, int query,
                          uint64_t* d_result, void* stream)
{
  if (int rc = bloom_args(k, num_hashes, d_filter_words, filter_bits, d_result)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  NTH_CUDA(cudaMemsetAsync(d_result, 0, 3 * sizeof(uint64_t), st));
  if (n_reads == 0 || max_read_len < k) return NTHASH_OK;
  if (!d_bases || ((uintptr_t)d_bases & 15)) return fail(NTHASH_ERR_INVALID_ARG, "d_bases must be 16-byte aligned");
  if (!d_read_off || !d_koff) return fail(NTHASH_ERR_INVALID_ARG, "d_read_off and d_koff must not be NULL");
  DevBatch B;
  B.d_bases = d_bases;
  B.n_bases = n_bases_readable;
  B.d_read_off = d_read_off;
  B.d_koff = d_koff;
  B.n_reads = n_reads;
  B.max_len = max_read_len;
  B.d_reduce = d_result;
  B.d_bloom = d_filter_words;
  B.bloom_bits = filter_bits;
  B.bloom_mode = query ? 2 : 1;
  return kmer_dev_run(B, k, num_hashes, st);
}

// ---- consumer: minimizer selection ----------------------------------------------------------------------------
// Chunks of reads whose hash rows stay L2-resident between the hash kernel and the selection kernel; the rows themselves
// never leave the device (and, for the host entry, never cross PCIe): out come a bitmap over the dense rows and, on
// request, the selected hashes and their row numbers in row order.
struct MinimizerOut
{
  uint32_t* d_bits;
  uint64_t *d_hash, *d_row, capacity, *d_count;
};

static int minimizer_chunk(const DevBatch& geom, uint32_t k, uint32_t w, uint64_t rows, uint64_t row0, const uint64_t* d_koff_local,
                           uint64_t* d_rows, uint32_t* d_valid, uint32_t* d_end, const MinimizerOut& O, cudaStream_t st)
{
  DevBatch B = geom;
  B.d_out = d_rows;
  B.d_valid = d_valid;
  B.memset_rows = rows;
  if (int rc = kmer_dev_run(B, k, 1, st)) return rc;
  NTH_CUDA(launch_mark_read_ends(d_end, rows, d_koff_local, geom.n_reads, geom.uniform_len ? geom.uniform_len - k + 1 : 0, st));
  NTH_CUDA(launch_minimizer_select(d_rows, d_valid, d_end, rows, w, O.d_bits, row0, st));
  if (O.d_hash || O.d_row)
    NTH_CUDA(launch_compact_rows_at(d_rows, O.d_bits + row0 / 32, rows, 1, O.d_hash, O.d_row, O.d_count, true, row0, O.capacity, st));
  return NTHASH_OK;
}

static int minimizer_args(uint32_t k, uint32_t w, const void* bits, const void* count)
{
  if (int rc = check_kh(k, 1)) return rc;
  if (w < 1 || w > 64) return fail(NTHASH_ERR_INVALID_ARG, "window=%u k-mers outside [1, 64]", w);
  if (!bits || !count) return fail(NTHASH_ERR_INVALID_ARG, "the minimizer bitmap and the count must not be NULL");
  return NTHASH_OK;
}

int nthash_kmer_minimizer_uniform_dev(const uint8_t* d_bases, uint64_t n_bases_readable, uint64_t n_reads, uint32_t read_len, uint32_t k,
                                      uint32_t window, uint32_t* d_min_bits, uint64_t* d_min_hash, uint64_t* d_min_row, uint64_t capacity,
                                      uint64_t* d_count, void* stream)
{
  if (int rc = minimizer_args(k, window, d_min_bits, d_count)) return rc;
  if (int rc = check_device_ready()) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  NTH_CUDA(cudaMemsetAsync(d_count, 0, sizeof(uint64_t), st));
  if (n_reads == 0 || read_len < k) return NTHASH_OK;
  if (!d_bases || ((uintptr_t)d_bases & 15)) return fail(NTHASH_ERR_INVALID_ARG, "d_bases must be 16-byte aligned");
  if (n_bases_readable < n_reads * (uint64_t)read_len) return fail(NTHASH_ERR_INVALID_ARG, "n_bases_readable smaller than n_reads*read_len");
  const uint64_t nk = read_len - k + 1, total = n_reads * nk;
  NTH_CUDA(cudaMemsetAsync(d_min_bits, 0, ((total + 31) / 32) * 4, st));
  // reads per chunk: a multiple of 32 (chunk rows start on a word of the bitmap, chunk bases on a 16-byte boundary),
  // about 8 M rows = 64 MB of hashes, which the 126 MB L2 keeps for the selection pass
  uint64_t per = std::max<uint64_t>(32, ((8ull << 20) / nk) & ~31ull);
  if (const char* e = getenv("NTHASH_B200_MINIMIZER_CHUNK_READS")) per = std::max<uint64_t>(32, strtoull(e, nullptr, 10) & ~31ull);
  per = std::min<uint64_t>(per, (n_reads + 31) & ~31ull);
  uint64_t* d_rows = nullptr;
  uint32_t *d_valid = nullptr, *d_end = nullptr;
  const uint64_t words = (per * nk + 31) / 32 + 4;
  NTH_CUDA(cudaMallocAsync(&d_rows, (per * nk + 64) * sizeof(uint64_t), st));
  cudaError_t e = cudaMallocAsync(&d_valid, 2 * words * 4, st);
  int rc = e == cudaSuccess ? NTHASH_OK : fail(NTHASH_ERR_CUDA, "cudaMallocAsync: %s", cudaGetErrorString(e));
  d_end = d_valid + words;
  const MinimizerOut O = { d_min_bits, d_min_hash, d_min_row, capacity, d_count };
  for (uint64_t r0 = 0; r0 < n_reads && rc == NTHASH_OK; r0 += per) {
    DevBatch G;
    G.d_bases = d_bases + r0 * read_len;
    G.n_bases = n_bases_readable - r0 * read_len;
    G.n_reads = std::min(per, n_reads - r0);
    G.uniform_len = read_len;
    rc = minimizer_chunk(G, k, window, G.n_reads * nk, r0 * nk, nullptr, d_rows, d_valid, d_end, O, st);
  }
  if (rc == NTHASH_OK && !(d_min_hash || d_min_row)) {
    e = launch_popcount(d_min_bits, total, d_count, st);
    if (e != cudaSuccess) rc = fail(NTHASH_ERR_CUDA, "launch_popcount: %s", cudaGetErrorString(e));
  }
  if (d_valid) cudaFreeAsync(d_valid, st);
  cudaFreeAsync(d_rows, st);
  return rc;
}

// Host buffers: the bases go up (ASCII), the bitmap and the selected (hash, row) pairs come back; fixed-length reads when
// read_off is NULL.  One stream, chunk after chunk (the selection needs each chunk's rows while they are cache-resident).
int nthash_kmer_minimizers(const char* bases, const uint64_t* read_off, uint64_t n_reads, uint32_t uniform_read_len, uint32_t k,
                           uint32_t window, uint32_t* min_bits, uint64_t* min_hash, uint64_t* min_row, uint64_t capacity, uint64_t* count,
                           int device)
{
  if (int rc = minimizer_args(k, window, min_bits, count)) return rc;
  *count = 0;
  if (n_reads == 0) return NTHASH_OK;
  if (!bases) return fail(NTHASH_ERR_INVALID_ARG, "bases must not be NULL");
  if (!read_off && !uniform_read_len) return fail(NTHASH_ERR_INVALID_ARG, "give read_off or a uniform_read_len > 0");
  if ((min_hash || min_row) && !capacity) return fail(NTHASH_ERR_INVALID_ARG, "capacity must be > 0 when the lists are requested");
  if (cudaSetDevice(device) != cudaSuccess) return fail(NTHASH_ERR_NO_DEVICE, "cannot select CUDA device %d", device);
  if (int rc = check_device_ready()) return rc;
  std::vector<uint64_t> koff;
  uint64_t total, n_bases, base0 = 0;
  if (read_off) {
    koff.resize(n_reads + 1);
    total = nthash_window_rows(read_off, n_reads, k, koff.data());
    base0 = read_off[0];
    n_bases = read_off[n_reads] - base0;
  } else {
    total = uniform_read_len >= k ? n_reads * (uint64_t)(uniform_read_len - k + 1) : 0;
    n_bases = n_reads * (uint64_t)uniform_read_len;
  }
  std::fill(min_bits, min_bits + (total + 31) / 32, 0u);
  if (total == 0) return NTHASH_OK;
  cudaStream_t st = nullptr;
  uint8_t* d_bases = nullptr;
  uint32_t* d_bits = nullptr;
  uint64_t *d_hash = nullptr, *d_row = nullptr, *d_count = nullptr, *d_off = nullptr, *d_rows = nullptr;
  uint32_t* d_valid = nullptr;
  std::vector<uint64_t> h_off;
  int rc = NTHASH_OK;
  auto cu = [&](cudaError_t e, const char* what) {
    if (e != cudaSuccess && rc == NTHASH_OK) rc = fail(NTHASH_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return e == cudaSuccess;
  };
  const uint64_t bwords = (total + 31) / 32;
  const uint64_t cap = (min_hash || min_row) ? capacity : 0;
  if (cu(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), "cudaStreamCreate") &&
      cu(cudaMallocAsync(&d_bases, n_bases + 96, st), "cudaMallocAsync") && cu(cudaMallocAsync(&d_bits, bwords * 4 + 16, st), "cudaMallocAsync") &&
      cu(cudaMallocAsync(&d_count, 8, st), "cudaMallocAsync") && (!min_hash || cu(cudaMallocAsync(&d_hash, cap * 8, st), "cudaMallocAsync")) &&
      (!min_row || cu(cudaMallocAsync(&d_row, cap * 8, st), "cudaMallocAsync"))) {
    cu(cudaMemcpyAsync(d_bases + 16, bases + base0, n_bases, cudaMemcpyHostToDevice, st), "H2D bases");
    cu(cudaMemsetAsync(d_bits, 0, bwords * 4 + 16, st), "memset");
    cu(cudaMemsetAsync(d_count, 0, 8, st), "memset");
    const MinimizerOut O = { d_bits, d_hash, d_row, cap, d_count };
    if (!read_off) {
      // the device entry does its own chunking; its bitmap / count arguments are ours (it zeroes them again: harmless)
      if (rc == NTHASH_OK)
        rc = nthash_kmer_minimizer_uniform_dev(d_bases + 16, n_bases + 64, n_reads, uniform_read_len, k, window, d_bits, d_hash, d_row, cap, d_count, st);
    } else {
      // ragged: chunk boundaries where the dense row number is a multiple of 32 (whole bitmap words per chunk)
      const uint64_t CH_ROWS = 8ull << 20;
      uint64_t max_reads = 0, max_rows = 0;
      std::vector<uint64_t> cut(1, 0);
      for (uint64_t r = 0; r < n_reads;) {
        uint64_t e = r + 1;
        while (e < n_reads && (koff[e] - koff[r] < CH_ROWS || (koff[e] & 31))) ++e;
        if (e < n_reads && (koff[e] & 31)) e = n_reads;
        cut.push_back(e);
        max_reads = std::max(max_reads, e - r);
        max_rows = std::max(max_rows, koff[e] - koff[r]);
        r = e;
      }
      const uint64_t words = (max_rows + 31) / 32 + 4;
      if (cu(cudaMallocAsync(&d_off, 2 * (max_reads + 1) * 8, st), "cudaMallocAsync") && cu(cudaMallocAsync(&d_rows, (max_rows + 64) * 8, st), "cudaMallocAsync") &&
          cu(cudaMallocAsync(&d_valid, 2 * words * 4, st), "cudaMallocAsync")) {
        for (size_t c = 0; c + 1 < cut.size() && rc == NTHASH_OK; ++c) {
          const uint64_t r0 = cut[c], r1 = cut[c + 1], nr = r1 - r0, rows = koff[r1] - koff[r0];
          if (rows == 0) continue;
          cu(cudaStreamSynchronize(st), "sync"); // h_off is reused
          h_off.resize(2 * (nr + 1));
          uint64_t mx = 0;
          const uint64_t b0 = read_off[r0] & ~15ull; // chunk bases start on a 16-byte boundary of the device copy (base0 sits at +16)
          const uint64_t shift = (read_off[r0] - base0) - ((read_off[r0] - base0) & ~15ull);
          (void)b0;
          for (uint64_t i = 0; i <= nr; ++i) {
            h_off[i] = read_off[r0 + i] - read_off[r0] + shift;
            h_off[nr + 1 + i] = koff[r0 + i] - koff[r0];
            if (i < nr) mx = std::max(mx, read_off[r0 + i + 1] - read_off[r0 + i]);
          }
          cu(cudaMemcpyAsync(d_off, h_off.data(), 2 * (nr + 1) * 8, cudaMemcpyHostToDevice, st), "H2D offsets");
          DevBatch G;
          G.d_bases = d_bases + 16 + ((read_off[r0] - base0) & ~15ull);
          G.n_bases = n_bases + 64 - ((read_off[r0] - base0) & ~15ull);
          G.d_read_off = d_off;
          G.d_koff = d_off + nr + 1;
          G.n_reads = nr;
          G.max_len = mx;
          if (rc == NTHASH_OK) rc = minimizer_chunk(G, k, window, rows, koff[r0], d_off + nr + 1, d_rows, d_valid, d_valid + words, O, st);
        }
      }
    }
    if (rc == NTHASH_OK && !cap) cu(launch_popcount(d_bits, total, d_count, st), "launch_popcount");
    cu(cudaMemcpyAsync(min_bits, d_bits, bwords * 4, cudaMemcpyDeviceToHost, st), "D2H bitmap");
    cu(cudaMemcpyAsync(count, d_count, 8, cudaMemcpyDeviceToHost, st), "D2H count");
    cu(cudaStreamSynchronize(st), "sync");
    if (rc == NTHASH_OK && cap) {
      const uint64_t n = std::min<uint64_t>(*count, cap);
      if (min_hash) cu(cudaMemcpyAsync(min_hash, d_hash, n * 8, cudaMemcpyDeviceToHost, st), "D2H hashes");
      if (min_row) cu(cudaMemcpyAsync(min_row, d_row, n * 8, cudaMemcpyDeviceToHost, st), "D2H rows");
      cu(cudaStreamSynchronize(st), "sync");
    }
  }
  for (void* q : { (void*)d_bases, (void*)d_bits, (void*)d_hash, (void*)d_row, (void*)d_count, (void*)d_off, (void*)d_rows, (void*)d_valid })
    if (q) cudaFreeAsync(q, st);
  if (st) {
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
  }
  return rc;
}

// ---- fused consumer: ntCard-style cardinality sketch ---------------------------------------------------

static int sketch_args(uint32_t k, uint32_t sample_bits, uint32_t index_bits, const uint32_t* d_counters, const uint64_t* d_result)
{
  if (int rc = check_kh(k, 1)) return rc;
  if (sample_bits < 1 || index_bits < 1 || index_bits > 31 || sample_bits + index_bits > 63)
    return fail(NTHASH_ERR_INVALID_ARG, "sample_bits=%u, index_bits=%u: need 1 <= sample_bits, 1 <= index_bits <= 31, sum <= 63", sample_bits, index_bits);
  if (!d_counters || ((uintptr_t)d_counters & 3)) return fail(NTHASH_ERR_INVALID_ARG, "d_counters must not be NULL (4-byte aligned)");
  if (!d_result) return fail(NTHASH_ERR_INVALID_ARG, "d_result must not be NULL");
  return check_device_ready();
}

int nthash_kmer_sketch_uniform_dev(const uint8_t* d_bases, uint64_t n_bases_readable, uint64_t n_reads, uint32_t read_len, uint32_t k,
                                   uint32_t sample_bits, uint32_t index_bits, uint32_t* d_counters, uint64_t* d_result, void* stream)
{
  if (int rc = sketch_args(k, sample_bits, index_bits, d_counters, d_result)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  NTH_CUDA(cudaMemsetAsync(d_result, 0, 3 * sizeof(uint64_t), st));
  if (n_reads == 0 || read_len < k) return NTHASH_OK;
  if (!d_bases || ((uintptr_t)d_bases & 15)) return fail(NTHASH_ERR_INVALID_ARG, "d_bases must be 16-byte aligned");
  if (n_bases_readable < n_reads * (uint64_t)read_len) return fail(NTHASH_ERR_INVALID_ARG, "n_bases_readable smaller than n_reads*read_len");
  DevBatch B;
  B.d_bases = d_bases;
  B.n_bases = n_bases_readable;
  B.n_reads = n_reads;
  B.uniform_len = read_len;
  B.d_reduce = d_result;
  B.d_bloom = d_counters;
  B.bloom_bits = ((uint64_t)sample_bits << 8) | index_bits;
  B.bloom_mode = 3;
  return kmer_dev_run(B, k, 1, st);
}

int nthash_kmer_sketch_dev(const uint8_t* d_bases, uint64_t n_bases_readable, const uint64_t* d_read_off, const uint64_t* d_koff,
                           uint64_t n_reads, uint64_t max_read_len, uint32_t k, uint32_t sample_bits, uint32_t index_bits,
                           uint32_t* d_counters, uint64_t* d_result, void* stream)
{
  if (int rc = sketch_args(k, sample_bits, index_bits, d_counters, d_result)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  NTH_CUDA(cudaMemsetAsync(d_result, 0, 3 * sizeof(uint64_t), st));
  if (n_reads == 0 || max_read_len < k) return NTHASH_OK;
  if (!d_bases || ((uintptr_t)d_bases & 15)) return fail(NTHASH_ERR_INVALID_ARG, "d_bases must be 16-byte aligned");
  if (!d_read_off || !d_koff) return fail(NTHASH_ERR_INVALID_ARG, "d_read_off and d_koff must not be NULL");
  DevBatch B;
  B.d_bases = d_bases;
  B.n_bases = n_bases_readable;
  B.d_read_off = d_read_off;
  B.d_koff = d_koff;
  B.n_reads = n_reads;
  B.max_len = max_read_len;
  B.d_reduce = d_result;
  B.d_bloom = d_counters;
  B.bloom_bits = ((uint64_t)sample_bits << 8) | index_bits;
  B.bloom_mode = 3;
  return kmer_dev_run(B, k, 1, st);
}

// ---- SeedNtHash -------------------------------------------------------------------------------

int nthash_seed_plan_create(const char* const* seeds, uint32_t n_seeds, uint32_t k, uint32_t num_hashes_per_seed,
                            nthash_seed_plan** plan_out)
{
  if (!plan_out) return fail(NTHASH_ERR_INVALID_ARG, "plan_out must not be NULL");
  *plan_out = nullptr;
  if (int rc = check_kh(k, num_hashes_per_seed)) return rc;
  if (int rc = check_device_ready()) return rc;
  nthash_seed_plan* plan = new nthash_seed_plan();
  const std::string err = build_seed_plan(seeds, n_seeds, k, num_hashes_per_seed, plan->host);
  if (!err.empty()) {
    delete plan;
    return fail(NTHASH_ERR_INVALID_ARG, "%s", err.c_str());
  }
  cudaGetDevice(&plan->device);
  cudaError_t e = cudaMalloc(&plan->d_blob, plan->host.blob.size());
  if (e == cudaSuccess) e = cudaMemcpy(plan->d_blob, plan->host.blob.data(), plan->host.blob.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(plan->d_blob);
    delete plan;
    return fail(NTHASH_ERR_CUDA, "seed plan upload: %s", cudaGetErrorString(e));
  }
  plan->jit = seed_jit_build(plan->host, plan->jit_note, true);
  if (plan->jit) plan->jit_note = "specialised kernel compiled with NVRTC";
  *plan_out = plan;
  return NTHASH_OK;
}

void nthash_seed_plan_destroy(nthash_seed_plan* plan)
{
  if (!plan) return;
  seed_jit_destroy(plan->jit);
  cudaFree(plan->d_blob);
  delete plan;
}

const char* nthash_seed_plan_kernel_note(const nthash_seed_plan* plan) { return plan ? plan->jit_note.c_str() : ""; }

int nthash_seed_jit_selftest(const char* const* seeds, uint32_t n_seeds, uint32_t k, uint32_t num_hashes_per_seed)
{
  SeedPlanHost host;
  const std::string err = build_seed_plan(seeds, n_seeds, k, num_hashes_per_seed, host);
  if (!err.empty()) return fail(NTHASH_ERR_INVALID_ARG, "%s", err.c_str());
  std::string why;
  if (!seed_jit_compile_all(host, why)) return fail(NTHASH_ERR_UNSUPPORTED, "%s", why.c_str());
  return NTHASH_OK;
}

int nthash_seed_plan_symmetric(const nthash_seed_plan* plan) { return plan && plan->host.all_symmetric ? 1 : 0; }

int nthash_seed_batch_uniform_dev(const nthash_seed_plan* plan, const uint8_t* d_bases, uint64_t n_bases_readable,
                                  uint64_t n_reads, uint32_t read_len, uint64_t* d_out, uint32_t* d_valid_bits,
                                  uint64_t* d_out_fwd, uint64_t* d_out_rev, void* stream)
{
  if (!plan) return fail(NTHASH_ERR_INVALID_ARG, "plan must not be NULL");
  const uint32_t k = plan->host.k;
  if (n_reads == 0 || read_len < k) return NTHASH_OK;
  if (int rc = check_outputs(d_out, d_out_fwd, d_out_rev)) return rc;
  if (!d_bases || ((uintptr_t)d_bases & 15)) return fail(NTHASH_ERR_INVALID_ARG, "d_bases must be 16-byte aligned");
  if (n_bases_readable < n_reads * (uint64_t)read_len)
    return fail(NTHASH_ERR_INVALID_ARG, "n_bases_readable smaller than n_reads*read_len");
  if (int rc = check_device_ready()) return rc;
  DevBatch B;
  B.d_bases = d_bases;
  B.n_bases = n_bases_readable;
  B.n_reads = n_reads;
  B.uniform_len = read_len;
  B.d_out = d_out;
  B.d_valid = d_valid_bits;
  B.memset_rows = n_reads * (uint64_t)(read_len - k + 1);
  B.d_fwd = d_out_fwd;
  B.d_rev = d_out_rev;
  return seed_dev_run(plan, B, (cudaStream_t)stream);
}

int nthash_blind_seed_roll_batch_dev(const nthash_seed_plan* plan, uint8_t* d_kmers, uint64_t n_bytes_readable,
                                     const uint8_t* d_in_base, uint64_t n, uint64_t* d_out, uint64_t* d_out_fwd,
                                     uint64_t* d_out_rev, void* stream)
{
  if (!plan) return fail(NTHASH_ERR_INVALID_ARG, "plan must not be NULL");
  if (n == 0) return NTHASH_OK;
  if (!d_kmers || !d_in_base) return fail(NTHASH_ERR_INVALID_ARG, "d_kmers and d_in_base must not be NULL");
  if (int rc = check_device_ready()) return rc;
  NTH_CUDA(launch_blind_seed_shift(d_kmers, d_in_base, n, plan->host.k, (cudaStream_t)stream));
  // the new windows are n reads of exactly k bases: one row each (SeedNtHash's init hashes whatever bytes it is given,
  // like BlindSeedNtHash; no validity bitmap is produced)
  return nthash_seed_batch_uniform_dev(plan, d_kmers, n_bytes_readable, n, plan->host.k, d_out, nullptr, d_out_fwd, d_out_rev, stream);
}

int nthash_seed_batch_dev(const nthash_seed_plan* plan, const uint8_t* d_bases, uint64_t n_bases_readable,
                          const uint64_t* d_read_off, const uint64_t* d_koff, uint64_t n_reads, uint64_t max_read_len,
                          uint64_t* d_out, uint32_t* d_valid_bits, uint64_t* d_out_fwd, uint64_t* d_out_rev,
                          void* stream)
{
  if (!plan) return fail(NTHASH_ERR_INVALID_ARG, "plan must not be NULL");
  const uint32_t k = plan->host.k;
  if (n_reads == 0 || max_read_len < k) return NTHASH_OK;
  if (int rc = check_outputs(d_out, d_out_fwd, d_out_rev)) return rc;
  if (!d_bases || ((uintptr_t)d_bases & 15)) return fail(NTHASH_ERR_INVALID_ARG, "d_bases must be 16-byte aligned");
  if (!d_read_off || !d_koff) return fail(NTHASH_ERR_INVALID_ARG, "d_read_off and d_koff must not be NULL");
  if (int rc = check_device_ready()) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  DevBatch B;
  B.d_rows = d_koff + n_reads;
  B.rows_bound = n_bases_readable;
  B.d_bases = d_bases;
  B.n_bases = n_bases_readable;
  B.d_read_off = d_read_off;
  B.d_koff = d_koff;
  B.n_reads = n_reads;
  B.max_len = max_read_len;
  B.d_out = d_out;
  B.d_valid = d_valid_bits;
  B.d_fwd = d_out_fwd;
  B.d_rev = d_out_rev;
  return seed_dev_run(plan, B, st);
}

// The host entries take seed strings, not a plan: compiled plans (an NVRTC build, ~0.3 s) are kept per (device, seeds, h)
// for the life of the process so that repeated calls do not recompile.
static int cached_seed_plan(const char* const* seeds, uint32_t n_seeds, uint32_t k, uint32_t h, const nthash_seed_plan** out)
{
  static std::mutex mu;
  static std::map<std::string, nthash_seed_plan*> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  std::string key = std::to_string(dev) + "/" + std::to_string(k) + "/" + std::to_string(h);
  for (uint32_t i = 0; i < n_seeds; ++i) key += std::string("/") + (seeds && seeds[i] ? seeds[i] : "");
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it == cache.end()) {
    nthash_seed_plan* plan = nullptr;
    if (int rc = nthash_seed_plan_create(seeds, n_seeds, k, h, &plan)) return rc;
    it = cache.emplace(key, plan).first;
  }
  *out = it->second;
  return NTHASH_OK;
}

int nthash_seed_batch(const char* bases, const uint64_t* read_off, uint64_t n_reads, const char* const* seeds,
                      uint32_t n_seeds, uint32_t k, uint32_t num_hashes_per_seed, uint64_t* out,
                      uint32_t* valid_bits, uint64_t* out_fwd, uint64_t* out_rev, int device)
{
  if (int rc = check_kh(k, num_hashes_per_seed)) return rc;
  if (n_reads == 0) return NTHASH_OK;
  if (!bases || !read_off) return fail(NTHASH_ERR_INVALID_ARG, "bases and read_off must not be NULL");
  if (int rc = check_outputs(out, out_fwd, out_rev)) return rc;
  if (cudaSetDevice(device) != cudaSuccess) return fail(NTHASH_ERR_NO_DEVICE, "cannot select CUDA device %d", device);
  const nthash_seed_plan* plan = nullptr;
  if (int rc = cached_seed_plan(seeds, n_seeds, k, num_hashes_per_seed, &plan)) return rc;
  const HostBatch hb = { bases, read_off, n_reads, k, (uint64_t)n_seeds * num_hashes_per_seed, out_fwd ? (uint64_t)n_seeds : 0ull,
                         out, valid_bits, out_fwd, out_rev };
  return host_pipeline(hb, [&](const DevBatch& B, cudaStream_t st) { return seed_dev_run(plan, B, st); });
}

// Fused SeedNtHash consumer for a uniform device-resident batch: the specialised kernel accumulates count / sum / xor of
// every item without a byte for the exact path (no hash is stored), a second small launch redoes the flagged items window
// by window.  Adds into d_result (not cleared here).  Returns 1 when the fused form does not apply to this plan / geometry.
static int seed_reduce_fused(const nthash_seed_plan* plan, const uint8_t* d_bases, uint64_t n_bases, uint64_t n_reads, uint32_t read_len,
                             uint64_t* d_result, cudaStream_t st)
{
  if (getenv("NTHASH_B200_DISABLE_SEED_JIT")) return 1;
  SeedParams P;
  fill_seed_params(plan, P);
  P.bases = d_bases;
  P.n_bases = n_bases;
  P.reduce_out = d_result;
  if (!plan_uniform(n_reads, read_len, P.k, P.g, P.tile_cap)) return 1;
  if (!seed_jit_reduce_applies(plan->jit, P)) return 1;
  uint8_t* d_dirty = nullptr;
  NTH_CUDA(cudaMallocAsync(&d_dirty, n_reads + P.g.n_items, st));
  cudaError_t e = cudaMemsetAsync(d_dirty, 0, n_reads + P.g.n_items, st);
  P.read_dirty = d_dirty;
  P.item_dirty = d_dirty + n_reads;
  if (e == cudaSuccess) e = launch_seed_jit(plan->jit, P, st);
  if (e == cudaSuccess) e = launch_seed_reduce_dirty(P, n_reads, st);
  cudaFreeAsync(d_dirty, st);
  NTH_CUDA(e);
  return NTHASH_OK;
}

// SeedNtHash consumer: fused into the specialised kernel for fixed-length reads (seed_reduce_fused); otherwise two passes on
// the device (the seed kernels write a chunk of rows, a reduction reads them back).  No hash crosses PCIe either way.
// result = {windows visited, sum, xor} over all n_seeds * num_hashes_per_seed values.
int nthash_seed_reduce(const char* bases, const uint64_t* read_off, uint64_t n_reads, const char* const* seeds, uint32_t n_seeds,
                       uint32_t k, uint32_t num_hashes_per_seed, uint64_t* result, int device)
{
  if (int rc = check_kh(k, num_hashes_per_seed)) return rc;
  if (!result) return fail(NTHASH_ERR_INVALID_ARG, "result must not be NULL");
  result[0] = result[1] = result[2] = 0;
  if (n_reads == 0) return NTHASH_OK;
  if (!bases || !read_off) return fail(NTHASH_ERR_INVALID_ARG, "bases and read_off must not be NULL");
  if (cudaSetDevice(device) != cudaSuccess) return fail(NTHASH_ERR_NO_DEVICE, "cannot select CUDA device %d", device);
  const nthash_seed_plan* plan = nullptr;
  if (int rc = cached_seed_plan(seeds, n_seeds, k, num_hashes_per_seed, &plan)) return rc;
  const uint32_t H = n_seeds * num_hashes_per_seed;
  HostBatch hb = { bases, read_off, n_reads, k, (uint64_t)H, 0ull, nullptr, nullptr, nullptr, nullptr };
  hb.reduce_result = result;
  hb.scratch_rows = true;
  const int rc = host_pipeline(hb, [&](const DevBatch& B, cudaStream_t st) {
    if (B.uniform_len) {
      const int f = seed_reduce_fused(plan, B.d_bases, B.n_bases, B.n_reads, B.uniform_len, B.d_reduce, st);
      if (f != 1) return f;
    }
    if (int r = seed_dev_run(plan, B, st)) return r;
    NTH_CUDA(launch_reduce_rows(B.d_out, B.d_valid, B.valid_row0, B.rows, H, B.d_reduce, st));
    return (int)NTHASH_OK;
  });
  return rc;
}

int nthash_seed_reduce_uniform_dev(const nthash_seed_plan* plan, const uint8_t* d_bases, uint64_t n_bases_readable, uint64_t n_reads,
                                   uint32_t read_len, uint64_t* d_result, void* stream)
{
  if (!plan) return fail(NTHASH_ERR_INVALID_ARG, "plan must not be NULL");
  if (!d_result) return fail(NTHASH_ERR_INVALID_ARG, "d_result must not be NULL");
  if (int rc = check_device_ready()) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  NTH_CUDA(cudaMemsetAsync(d_result, 0, 3 * sizeof(uint64_t), st));
  const uint32_t k = plan->host.k, H = plan->host.n_seeds * plan->host.h;
  if (n_reads == 0 || read_len < k) return NTHASH_OK;
  if (!d_bases || ((uintptr_t)d_bases & 15)) return fail(NTHASH_ERR_INVALID_ARG, "d_bases must be 16-byte aligned");
  if (n_bases_readable < n_reads * (uint64_t)read_len)
    return fail(NTHASH_ERR_INVALID_ARG, "n_bases_readable smaller than n_reads*read_len");
  {
    const int f = seed_reduce_fused(plan, d_bases, n_bases_readable, n_reads, read_len, d_result, st);
    if (f != 1) return f;
  }
  // two passes: chunks of reads whose rows fit a fixed scratch buffer (~512 MB of hashes); chunk starts stay 16-byte aligned
  const uint64_t nk = read_len - k + 1;
  uint64_t per = std::max<uint64_t>(16, ((64ull << 20) / (nk * H)) & ~15ull);
  per = std::min<uint64_t>(per, (n_reads + 15) & ~(uint64_t)15);
  uint64_t* d_out = nullptr;
  uint32_t* d_valid = nullptr;
  NTH_CUDA(cudaMallocAsync(&d_out, per * nk * H * sizeof(uint64_t), st));
  cudaError_t e = cudaMallocAsync(&d_valid, ((per * nk + 31) / 32) * 4, st);
  int rc = e == cudaSuccess ? NTHASH_OK : fail(NTHASH_ERR_CUDA, "cudaMallocAsync: %s", cudaGetErrorString(e));
  for (uint64_t r0 = 0; r0 < n_reads && rc == NTHASH_OK; r0 += per) {
    const uint64_t nr = std::min(per, n_reads - r0);
    DevBatch B;
    B.d_bases = d_bases + r0 * read_len;
    B.n_bases = n_bases_readable - r0 * read_len;
    B.n_reads = nr;
    B.uniform_len = read_len;
    B.d_out = d_out;
    B.d_valid = d_valid;
    B.memset_rows = nr * nk;
    rc = seed_dev_run(plan, B, st);
    if (rc == NTHASH_OK) {
      e = launch_reduce_rows(d_out, d_valid, 0, nr * nk, H, d_result, st);
      if (e != cudaSuccess) rc = fail(NTHASH_ERR_CUDA, "launch_reduce_rows: %s", cudaGetErrorString(e));
    }
  }
  if (d_valid) cudaFreeAsync(d_valid, st);
  cudaFreeAsync(d_out, st);
  return rc;
}

int nthash_blind_roll_batch_dev(uint64_t* d_fwd, uint64_t* d_rev, const uint8_t* d_out_base,
                                const uint8_t* d_in_base, uint64_t n, uint32_t k, uint32_t num_hashes,
                                uint64_t* d_out, void* stream)
{
  if (int rc = check_kh(k, num_hashes)) return rc;
  if (n == 0) return NTHASH_OK;
  if (!d_fwd || !d_rev || !d_out_base || !d_in_base || !d_out)
    return fail(NTHASH_ERR_INVALID_ARG, "blind roll: all pointers are required");
  if (int rc = check_device_ready()) return rc;
  NTH_CUDA(launch_blind(d_fwd, d_rev, d_out_base, d_in_base, n, k, num_hashes, d_out, false, (cudaStream_t)stream));
  return NTHASH_OK;
}

int nthash_blind_peek4_batch_dev(const uint64_t* d_fwd, const uint64_t* d_rev, const uint8_t* d_out_base,
                                 uint64_t n, uint32_t k, uint32_t num_hashes, uint64_t* d_out, void* stream)
{
  if (int rc = check_kh(k, num_hashes)) return rc;
  if (n == 0) return NTHASH_OK;
  if (!d_fwd || !d_rev || !d_out_base || !d_out)
    return fail(NTHASH_ERR_INVALID_ARG, "blind peek4: all pointers are required");
  if (int rc = check_device_ready()) return rc;
  NTH_CUDA(launch_blind(const_cast<uint64_t*>(d_fwd), const_cast<uint64_t*>(d_rev), d_out_base, nullptr, n, k,
                        num_hashes, d_out, true, (cudaStream_t)stream));
  return NTHASH_OK;
}

int nthash_blind_roll_batch(uint64_t* fwd, uint64_t* rev, const char* out_base, const char* in_base,
                            uint64_t n, uint32_t k, uint32_t num_hashes, uint64_t* out, int device)
{
  if (int rc = check_kh(k, num_hashes)) return rc;
  if (n == 0) return NTHASH_OK;
  if (!fwd || !rev || !out_base || !in_base || !out)
    return fail(NTHASH_ERR_INVALID_ARG, "blind roll: all pointers are required");
  if (cudaSetDevice(device) != cudaSuccess) return fail(NTHASH_ERR_NO_DEVICE, "cannot select CUDA device %d", device);
  if (int rc = check_device_ready()) return rc;
  uint64_t* d_state = nullptr; // fwd | rev | out
  uint8_t* d_chars = nullptr;  // out_base | in_base
  const size_t sb = n * sizeof(uint64_t);
  int rc = NTHASH_OK;
  cudaError_t e = cudaMalloc(&d_state, sb * (2 + (size_t)num_hashes));
  if (e == cudaSuccess) e = cudaMalloc(&d_chars, 2 * n);
  if (e == cudaSuccess) e = cudaMemcpy(d_state, fwd, sb, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_state + n, rev, sb, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_chars, out_base, n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_chars + n, in_base, n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    rc = nthash_blind_roll_batch_dev(d_state, d_state + n, d_chars, d_chars + n, n, k, num_hashes, d_state + 2 * n, nullptr);
    if (rc == NTHASH_OK) {
      e = cudaMemcpy(fwd, d_state, sb, cudaMemcpyDeviceToHost);
      if (e == cudaSuccess) e = cudaMemcpy(rev, d_state + n, sb, cudaMemcpyDeviceToHost);
      if (e == cudaSuccess) e = cudaMemcpy(out, d_state + 2 * n, sb * num_hashes, cudaMemcpyDeviceToHost);
    }
  }
  cudaFree(d_state);
  cudaFree(d_chars);
  if (e != cudaSuccess) return fail(NTHASH_ERR_CUDA, "blind roll (host): %s", cudaGetErrorString(e));
  return rc;
}

} // extern "C"
