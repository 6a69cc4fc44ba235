// engine.hpp — internal interfaces between the C ABI (capi.cu) and the kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <string>

namespace nthb {

constexpr int KMER_NT = 256;      // items (threads) per CTA of the k-mer / seed kernels
constexpr uint32_t SEG_LONG = 252; // windows per item when a read is cut up (63 words: odd => conflict-free LDS)
constexpr uint32_t TILE_BUDGET = 72 * 1024; // staged bytes per CTA we are willing to spend on whole-read items

// How work items map onto the base stream and the dense output.
//   ragged : item i starts at byte item_byte[i]; its windows are dense indices
//            [item_out[i], item_out[i+1]).  When every read is one item these arrays are
//            read_off / koff themselves.
//   uniform: item i = (read r = i / segs, segment s = i % segs) of fixed-length reads.
struct KmerGeom
{
  const uint64_t* item_byte = nullptr;
  const uint64_t* item_out = nullptr;
  // ragged batches, optional: item_perm[256 * b + t] = which of block b's 256 items thread t of the fast kernel's CTA b takes
  // (the block's items ordered by length class, longest first: a warp runs as long as its longest item).  Computed once per
  // layout by launch_item_perm; without it the kernel sorts its items itself (six CTA barriers per launch and CTA).
  const uint8_t* item_perm = nullptr;
  uint64_t n_items = 0;
  uint32_t read_len = 0, nk = 0, seg = 0, segs = 1;
  // flat (fast kernel, uniform long reads): item i = dense windows [i*seg, (i+1)*seg) of the whole batch, whichever
  // read(s) they fall in; `total` = n_reads * nk.  An item that crosses a read boundary is finished by the fix-up kernel.
  uint32_t flat = 0;
  uint64_t total = 0;
};

struct KmerParams
{
  const uint8_t* bases = nullptr;
  uint64_t n_bases = 0; // readable extent of `bases`
  KmerGeom g;
  uint32_t k = 0, h = 0;
  uint64_t* out = nullptr;
  uint32_t* valid_bits = nullptr;
  uint64_t valid_row0 = 0; // bit index of this batch's row 0 inside valid_bits (chunked host pipeline)
  uint64_t* out_fwd = nullptr;
  uint64_t* out_rev = nullptr;
  uint64_t* reduce_out = nullptr; // fused consumer: {windows visited, sum, xor} instead of out (all other outputs NULL)
  // fused Bloom-filter consumer (reduce_out = {windows visited, windows whose bits were all set already, 0}):
  uint32_t* bloom_words = nullptr; // bit b of the filter = bit (b & 31) of word b >> 5
  uint64_t bloom_bits = 0;         // filter size in bits; bit index = hash % bloom_bits
  uint32_t bloom_mode = 0;         // 0: none, 1: insert (atomicOr), 2: query, 3: cardinality sketch (bloom_words = counters, bloom_bits = s << 8 | r)
  uint32_t tile_cap = 0; // bytes of base tile a CTA may stage
  // ragged batches planned once (reads are items): the longest byte span of any block of 256 consecutive reads, measured on
  // the device at plan time (+ slack), instead of the 256 x longest-read bound — smaller tiles, more resident CTAs (0: unknown)
  uint32_t tile_cap_256 = 0;
  bool use_tma = true;   // allow the fast kernel (kmer_fast_kernel.cu) when the request permits
  bool general_fits = true; // false: the general kernel's CTA tile would not fit shared memory (huge k) - fast kernel or nothing
  uint64_t s[4], sk[4], mult[4]; // filled by launch_kmer
  uint32_t prefetch_ctas = 0;    // fast kernel: L2-prefetch the base tile this many CTAs ahead (0 = off)
  const uint4* t4 = nullptr;     // tetramer warm-up table (fast kernel only), filled by launch_kmer_fast
  uint32_t two = 2;              // the constant 2 as a kernel parameter (roll_step: keeps a multiply on the FMA pipe)
  // 2-bit packed input read directly by kmer_pack_kernel (then `bases` is not used): 4 bases per byte, A C G T = 0 1 2 3; bit j of
  // inv_bits (nullable) marks base j of the slice as invalid; the batch's base 0 is base `packed_first` of the slice
  const uint8_t* packed = nullptr;
  const uint32_t* inv_bits = nullptr;
  uint32_t packed_first = 0;
};

uint32_t kmer_smem_bytes(uint32_t tile_cap);
// Fast path (kmer_fast_kernel.cu): uniform batch, all items full, rows 16-byte multiples, h in {1,2,4}.
bool kmer_fast_ok(const KmerParams& P);
cudaError_t launch_kmer_fast(const KmerParams& P, cudaStream_t st);
// can a fixed-length batch of 2-bit packed reads be hashed straight from the packed bytes (no ASCII expansion)?
bool kmer_packed_direct_ok(uint64_t n_reads, uint32_t read_len, uint32_t k, uint32_t h);
cudaError_t launch_kmer(KmerParams P, cudaStream_t st);
// the per-(device, k) tetramer warm-up table of the fast kernel (created on first use, kept for the life of the process)
cudaError_t get_t4_table(uint32_t k, const uint4** out);

// ---- SeedNtHash (seed_kernel.cu) -------------------------------------------------------------
struct SeedParams
{
  const uint8_t* bases = nullptr;
  uint64_t n_bases = 0;
  KmerGeom g;
  const uint64_t* read_off = nullptr; // ragged batches (NULL: uniform); per READ, for the emission replay
  const uint64_t* koff = nullptr;
  const uint64_t* item_read = nullptr; // item -> read when reads are cut into several items
  uint32_t k = 0, h = 0, n_seeds = 0;
  uint64_t* out = nullptr;
  uint32_t* valid_bits = nullptr;
  uint64_t valid_row0 = 0; // bit index of this batch's row 0 inside valid_bits (chunked host pipeline)
  uint64_t* out_fwd = nullptr;
  uint64_t* out_rev = nullptr;
  uint8_t* read_dirty = nullptr; // one byte per read, zeroed by the caller
  uint64_t* reduce_out = nullptr; // fused consumer (uniform batches, specialised kernel): {windows visited, sum, xor}; no outputs then
  uint8_t* item_dirty = nullptr;  // fused consumer: one byte per item, zeroed by the caller
  uint32_t tile_cap = 0;
  const uint8_t* plan_blob = nullptr; // device copy of SeedPlanHost::blob
  uint32_t plan_smem_bytes = 0, groups_off = 0, tables_off = 0, care_off = 0, refblk_off = 0, care_words = 0;
  uint32_t any_ignore = 0;
  uint64_t s[4], sk[4];
};
uint32_t seed_smem_bytes(uint32_t plan_smem, uint32_t tile_cap);
cudaError_t launch_seed(SeedParams P, uint64_t n_reads, cudaStream_t st);      // generic hash kernel + emission replay
cudaError_t launch_seed_emit(const SeedParams& P, uint64_t n_reads, cudaStream_t st); // emission replay only
cudaError_t launch_seed_reduce_dirty(const SeedParams& P, uint64_t n_reads, cudaStream_t st); // fused consumer: the flagged items, exactly

// Seed kernel specialised per seed set at run time (seed_jit.cu).
struct SeedJit;
struct SeedPlanHost;
SeedJit* seed_jit_build(const SeedPlanHost& plan, std::string& why, bool load);
void seed_jit_destroy(SeedJit* j);
bool seed_jit_compile_all(const SeedPlanHost& plan, std::string& why); // every variant, compile only (no GPU needed)
const char* seed_jit_source(const SeedJit* j);
bool seed_jit_applies(const SeedJit* j, const SeedParams& P);
bool seed_jit_reduce_applies(const SeedJit* j, const SeedParams& P); // fused consumer variant (P.reduce_out)
cudaError_t launch_seed_jit(const SeedJit* j, const SeedParams& P, cudaStream_t st);

// BlindNtHash::roll / peek over n independent (fwd, rev) states (blind_kernel.cu).
cudaError_t launch_blind(uint64_t* fwd, uint64_t* rev, const uint8_t* out_base, const uint8_t* in_base, uint64_t n,
                         uint32_t k, uint32_t h, uint64_t* out, bool peek4, cudaStream_t st);

// FASTQ text in device memory -> concatenated bases + read_off (fastq_stage.cu); counts come back through host pointers.
cudaError_t fastq_extract(const uint8_t* d_text, uint64_t n_bytes, uint8_t* d_bases, uint64_t bases_capacity, uint64_t* d_read_off,
                          uint64_t reads_capacity, uint64_t* n_reads_out, uint64_t* n_bases_out, cudaStream_t st);

// BlindSeedNtHash::roll: shift every state's k-mer by one base and append in_base[i] (hashing = the seed batch path).
cudaError_t launch_blind_seed_shift(uint8_t* kmers, const uint8_t* in_base, uint64_t n, uint32_t k, cudaStream_t st);

// Uniform geometry for n_reads reads of read_len bases; returns false if unsupported.
bool plan_uniform(uint64_t n_reads, uint32_t read_len, uint32_t k, KmerGeom& g, uint32_t& tile_cap);
// Worst-case staged span of one CTA for a geometry cut into `seg`-window items.
uint32_t span_bound(uint32_t seg, uint32_t segs, uint32_t k);

// ---- ragged-batch preparation (prep_kernels.cu) ---------------------------------------
// koff[r] = sum_{r'<r} max(0, len_r' - k + 1), koff[n] = total; stats[0]=total, stats[1]=max len,
// stats[2] = number of items when reads are cut into seg-window items.  All device pointers.
cudaError_t launch_koff_scan(const uint64_t* read_off, uint64_t n_reads, uint32_t k, uint32_t seg,
                             uint64_t* koff, uint64_t* stats, cudaStream_t st);
// {rows with their validity bit set, sum, xor of all their values} accumulated into d_result (3 x u64, not zeroed here).
cudaError_t launch_reduce_rows(const uint64_t* d_out, const uint32_t* d_valid, uint64_t valid_row0, uint64_t rows, uint32_t H,
                               uint64_t* d_result, cudaStream_t st);
// Dense rows -> only the rows whose validity bit is set, in order (what the reference's `while (roll())` loop yields).
cudaError_t launch_compact_rows(const uint64_t* d_out, const uint32_t* d_valid, uint64_t rows, uint32_t H, uint64_t* d_compact,
                                uint64_t* d_row_index, uint64_t* d_count, cudaStream_t st);
cudaError_t launch_compact_rows_at(const uint64_t* d_out, const uint32_t* d_valid, uint64_t rows, uint32_t H, uint64_t* d_compact,
                                   uint64_t* d_row_index, uint64_t* d_count, bool append, uint64_t row_offset, uint64_t capacity,
                                   cudaStream_t st);
// 2-bit packed bases (+ optional invalid-base bitmap) -> ASCII; bases [first_base, first_base + n_bases) of the packed
// stream go to d_out[0 .. n_bases) (16-byte aligned, padded to a 16-byte multiple).
cudaError_t launch_unpack2bit(const uint8_t* d_packed, const uint32_t* d_invalid, uint64_t first_base, uint64_t n_bases, uint8_t* d_out,
                              cudaStream_t st);
// ---- minimizer selection (minimizer.cu): chunk-local hash rows + validity bitmap -> bits of the caller's bitmap ----
// d_end_bits: one bit per chunk row, set on the last row of every read (memset + marked here; (rows+31)/32 + 3 words).
cudaError_t launch_mark_read_ends(uint32_t* d_end_bits, uint64_t rows, const uint64_t* d_koff, uint64_t n_reads, uint32_t uniform_nk,
                                  cudaStream_t st);
// one thread per window of w consecutive k-mers; d_valid / d_end_bits need two readable words past their last one
cudaError_t launch_minimizer_select(const uint64_t* d_rows, const uint32_t* d_valid, const uint32_t* d_end_bits, uint64_t n_rows, uint32_t w,
                                    uint32_t* d_min_bits, uint64_t row0, cudaStream_t st);
cudaError_t launch_popcount(const uint32_t* d_bits, uint64_t n_bits, uint64_t* d_count, cudaStream_t st);
// Expands reads into items (only needed when some read exceeds the whole-read tile budget).  The tables hold `cap` + 1
// entries, `cap` >= the true item count (a host-side bound); the surplus is padded with empty items.
cudaError_t launch_item_fill(const uint64_t* read_off, const uint64_t* koff, uint64_t n_reads, uint32_t k,
                             uint32_t seg, uint64_t* item_byte, uint64_t* item_out, uint64_t* item_read,
                             uint64_t cap, cudaStream_t st);
// perm[256 * b + t] for every block b of 256 consecutive items (see KmerGeom::item_perm); perm holds ceil(n_items / 256) * 256 bytes
cudaError_t launch_item_perm(const uint64_t* item_out, uint64_t n_items, uint8_t* perm, cudaStream_t st);
// *d_max = max(*d_max, longest byte span of a block of 256 consecutive reads) (d_max zeroed by the caller)
cudaError_t launch_block_span_max(const uint64_t* read_off, uint64_t n_reads, uint64_t* d_max, cudaStream_t st);
// valid_bits <- ones for *d_rows rows (<= rows_bound, which only sizes the grid); nothing is read back.
cudaError_t launch_fill_valid(uint32_t* d_valid, const uint64_t* d_rows, uint64_t rows_bound, cudaStream_t st);

} // namespace nthb
