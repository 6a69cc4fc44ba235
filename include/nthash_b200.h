/*
 * nthash_b200.h — C ABI of the B200-native ntHash v2 batch engine (libnthash_b200.so).
 *
 * The reference (bcgsc/ntHash 2.4.0) has no FFI: its only interface is the C++ iterator
 * classes of include/nthash/nthash.hpp, whose hot loop `while (obj.roll())` lives in user
 * code and costs one library call per k-mer.  Each entry point below replaces that loop
 * for a whole batch of reads; the reference interface it stands in for is cited per
 * function as file:line in the reference tree.  include/nthash/nthash.hpp of THIS repo
 * re-exposes the reference's class API on top of these calls.
 *
 * Conventions
 *  - Plain C: pointers and sizes only.  Return 0 (NTHASH_OK) or a negative error code;
 *    the reference's raise_error() -> exit(1) (src/internal.hpp:16-22) never happens here.
 *    nthash_last_error() returns the thread-local message of the last failure.
 *  - Hash values are bit-identical to the reference ("ntHash_v2", nthash.hpp:18).
 *  - Input: reads laid back to back in `bases` (ASCII, one byte per base, as the reference
 *    takes them: nthash.hpp:74-78); read r is bases[read_off[r] .. read_off[r+1]).
 *  - Output: DENSE WINDOW ROWS.  koff[r] = sum_{r'<r} max(0, len_r' - k + 1) (reads shorter
 *    than k own no windows; the reference refuses to construct on them, kmer.cpp:215-220).
 *    The window of read r starting at base p (what the reference reports as get_pos()==p,
 *    nthash.hpp:170) is row w = koff[r] + p:
 *        out[w*H + j]      j-th hash of the window, H = num_hashes (k-mers) or
 *                          n_seeds*num_hashes_per_seed, seed-major (seed.cpp:167-171)
 *        valid_bits        bit (w & 31) of word (w >> 5) is 1 iff the reference's
 *                          while(roll()) loop visits that window (it skips windows by the
 *                          rules of kmer.cpp:228-264 / seed.cpp:493-544); rows whose bit is 0
 *                          read back as 0.  Bits past the last row are unspecified.
 *        out_fwd/out_rev   get_forward_hash()/get_reverse_hash() per window (k-mers) or per
 *                          window per seed (spaced seeds); optional.
 *    `valid_bits`, `out_fwd`, `out_rev` may be NULL (both of fwd/rev or neither).
 *  - `*_dev` entry points take DEVICE pointers on the current CUDA device and only enqueue
 *    work on `stream` (a cudaStream_t, NULL = default stream); nothing is copied and the
 *    host is not synchronised unless stated (nthash_kmer_plan_dev, nthash_ragged_plan_create and
 *    nthash_fastq_extract_dev return counts to the host and therefore do synchronise; the hashing
 *    entries never do: scratch comes from the stream-ordered pool, totals stay on the device).  `bases` must be 16-byte aligned (cudaMalloc
 *    memory is).  The un-suffixed entry points take HOST pointers and do H2D, the kernel
 *    and D2H themselves on `device`.
 *  - Supported domain: 3 <= k <= 65535 (the reference segfaults for k < 3, kmer.cpp:47),
 *    1 <= num_hashes <= 255 (uint8_t in the reference), any read lengths.
 *  - Byte semantics: every byte value is handled as the reference handles it (SEED_TAB,
 *    src/internal.hpp:132-165: ACGTU in either case hash, anything else is an invalid base
 *    for NtHash and a zero forward seed / `c & 7` complement seed for SeedNtHash), with ONE
 *    documented deviation: the raw control bytes 0x01, 0x03, 0x04, 0x05, 0x07.  They are the
 *    reference's complement slots (SEED_TAB[c & 7], internal.hpp:133) and make NtHash disagree
 *    with itself there (init() maps them through CONVERT_TAB = 255, roll() hashes them).  This
 *    engine — and oracle/, which restates the same choice — treats them as invalid bases in
 *    NtHash (no window containing one is emitted).  Text input never contains them.
 *  - Thread-safe: no mutable global state besides per-thread error text.
 */
#ifndef NTHASH_B200_H
#define NTHASH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NTHASH_OK 0
#define NTHASH_ERR_INVALID_ARG (-1) /* what the reference answers with raise_error()/exit(1) */
#define NTHASH_ERR_CUDA (-2)        /* a CUDA runtime call failed; see nthash_last_error() */
#define NTHASH_ERR_NO_DEVICE (-3)   /* no usable sm_100 device: the engine has no CPU fallback */
#define NTHASH_ERR_UNSUPPORTED (-4)

/* NTHASH_FN_NAME, include/nthash/nthash.hpp:18 */
const char* nthash_fn_name(void);
const char* nthash_last_error(void);
int nthash_b200_abi_version(void);
/* Number of CUDA devices the engine can run on (compute capability 10.x). */
int nthash_device_count(void);

/* ---- layout helpers (host arithmetic only) --------------------------------------- */
/* Fills koff[0..n_reads] (nullable) and returns the total number of window rows. */
uint64_t nthash_window_rows(const uint64_t* read_off, uint64_t n_reads, uint32_t k, uint64_t* koff);
/* Number of 32-bit words of a valid_bits array for `rows` rows. */
uint64_t nthash_valid_words(uint64_t rows);

/* ---- NtHash: contiguous k-mers ------------------------------------------------------
 * Replaces `nthash::NtHash it(seq, len, h, k); while (it.roll()) use(it.hashes())`
 * (nthash.hpp:62-211; NtHash::init/roll src/kmer.cpp:228-264; extend_hashes
 * src/internal.hpp:104-118) run over every read of the batch.                           */

/* Fixed-length reads: read r is bases[r*read_len .. (r+1)*read_len); koff[r] = r*(read_len-k+1).
 * `n_bases_readable` = readable extent of d_bases (>= n_reads*read_len). */
int nthash_kmer_batch_uniform_dev(const uint8_t* d_bases, uint64_t n_bases_readable, uint64_t n_reads,
                                  uint32_t read_len, uint32_t k, uint32_t num_hashes, uint64_t* d_out,
                                  uint32_t* d_valid_bits, uint64_t* d_out_fwd, uint64_t* d_out_rev,
                                  void* stream);

/* Ragged reads, step 1: d_koff[0..n_reads] from d_read_off on the device.  Synchronises
 * `stream` to return the totals the caller needs to size its buffers.                    */
int nthash_kmer_plan_dev(const uint64_t* d_read_off, uint64_t n_reads, uint32_t k, uint64_t* d_koff,
                         uint64_t* total_rows, uint64_t* max_read_len, void* stream);
/* Ragged reads, step 2.  d_koff/max_read_len as produced by nthash_kmer_plan_dev. */
int nthash_kmer_batch_dev(const uint8_t* d_bases, uint64_t n_bases_readable, const uint64_t* d_read_off,
                          const uint64_t* d_koff, uint64_t n_reads, uint64_t max_read_len, uint32_t k,
                          uint32_t num_hashes, uint64_t* d_out, uint32_t* d_valid_bits, uint64_t* d_out_fwd,
                          uint64_t* d_out_rev, void* stream);

/* Ragged reads, planned once.  nthash_ragged_plan_create() computes the layout of a batch — koff, the row total, the
 * longest read, the item tables the kernels use when reads are too long to be one work item, and the order in which
 * every CTA hands its items to its threads (by length class, so that a warp's lanes finish together) — and synchronises
 * the host ONCE; the plan then serves any number of calls on batches with the same read_off (the bases may change: same
 * layout, new sequence).  The *_planned_dev entries only enqueue kernels: no allocation that blocks, no read-back, no
 * host synchronisation, so they can be captured into a CUDA graph.  d_read_off is borrowed (keep it alive and
 * unchanged while the plan lives); the plan owns koff (nthash_ragged_plan_koff, device pointer, n_reads + 1 entries).
 * The reference has no counterpart: its get_pos() (nthash.hpp:170) is this arithmetic done one k-mer at a time.   */
typedef struct nthash_ragged_plan nthash_ragged_plan;
int nthash_ragged_plan_create(const uint64_t* d_read_off, uint64_t n_reads, uint32_t k, void* stream,
                              nthash_ragged_plan** plan_out);
void nthash_ragged_plan_destroy(nthash_ragged_plan* plan);
uint64_t nthash_ragged_plan_rows(const nthash_ragged_plan* plan);
uint64_t nthash_ragged_plan_max_read_len(const nthash_ragged_plan* plan);
const uint64_t* nthash_ragged_plan_koff(const nthash_ragged_plan* plan);
/* NtHash over the planned batch: same outputs and optional arguments as nthash_kmer_batch_dev. */
int nthash_kmer_batch_planned_dev(const nthash_ragged_plan* plan, const uint8_t* d_bases, uint64_t n_bases_readable,
                                  uint32_t num_hashes, uint64_t* d_out, uint32_t* d_valid_bits, uint64_t* d_out_fwd,
                                  uint64_t* d_out_rev, void* stream);
/* The fused count / sum / xor consumer over the planned batch (see nthash_kmer_reduce_dev). */
int nthash_kmer_reduce_planned_dev(const nthash_ragged_plan* plan, const uint8_t* d_bases, uint64_t n_bases_readable,
                                   uint32_t num_hashes, uint64_t* d_result, void* stream);

/* Host buffers in, host buffers out (H2D + kernel + D2H on `device`). */
int nthash_kmer_batch(const char* bases, const uint64_t* read_off, uint64_t n_reads, uint32_t k,
                      uint32_t num_hashes, uint64_t* out, uint32_t* valid_bits, uint64_t* out_fwd,
                      uint64_t* out_rev, int device);

/* The same batch over several GPUs of one box: contiguous read ranges with about equal
 * numbers of bases, one host thread and one chunked pipeline per listed device, no exchange between devices.
 * Outputs are the same dense arrays as nthash_kmer_batch's.                                                  */
int nthash_kmer_batch_multi(const char* bases, const uint64_t* read_off, uint64_t n_reads, uint32_t k,
                            uint32_t num_hashes, uint64_t* out, uint32_t* valid_bits, uint64_t* out_fwd,
                            uint64_t* out_rev, const int* devices, int n_devices);

/* Fixed-length batch from host memory (n_reads reads of read_len bases back to back): nthash_kmer_batch without
 * the offsets array.  Rows are n_reads * (read_len - k + 1); same outputs and optional arguments.            */
int nthash_kmer_batch_uniform(const char* bases, uint64_t n_reads, uint32_t read_len, uint32_t k, uint32_t num_hashes,
                              uint64_t* out, uint32_t* valid_bits, uint64_t* out_fwd, uint64_t* out_rev, int device);

/* ---- fused consumer (first of the "next" rows: the step after the hash path) --------------------
 * What the reference's own benchmark does with the hashes (examples/benchmark.cpp:34-39: a running
 * sum over `while (h.roll())`), done on the device so that no hash ever leaves the SM:
 *   result[0] = number of windows the reference's loop visits over the whole batch
 *   result[1] = 64-bit wrap-around sum, result[2] = xor, of ALL num_hashes values of those windows.
 * The device forms zero d_result themselves; the host form only uploads the bases.              */
int nthash_kmer_reduce_uniform_dev(const uint8_t* d_bases, uint64_t n_bases_readable, uint64_t n_reads,
                                   uint32_t read_len, uint32_t k, uint32_t num_hashes, uint64_t* d_result,
                                   void* stream);
int nthash_kmer_reduce_dev(const uint8_t* d_bases, uint64_t n_bases_readable, const uint64_t* d_read_off,
                           const uint64_t* d_koff, uint64_t n_reads, uint64_t max_read_len, uint32_t k,
                           uint32_t num_hashes, uint64_t* d_result, void* stream);
int nthash_kmer_reduce(const char* bases, const uint64_t* read_off, uint64_t n_reads, uint32_t k,
                       uint32_t num_hashes, uint64_t* result, int device);

/* ---- input format next to the path: 2-bit packed bases ------------------------------------------
 * The reference takes ASCII; pipelines that already hold 2-bit sequence can hand it over as is and move a quarter of
 * the bytes across PCIe.  Base g of the batch is bits 2*(g & 3).. of packed[g >> 2], 0 = A, 1 = C, 2 = G, 3 = T;
 * bit g of invalid_bits (nullable; 32-bit words, bit g & 31 of word g >> 5) marks a base that is not ACGT: it is
 * expanded to 'N', so every byte-level rule of the reference still applies (NtHash skips such windows,
 * kmer.cpp:232-235).  read_off indexes BASES exactly as in nthash_kmer_batch.  The device helper expands bases
 * [first_base, first_base + n_bases) of a packed stream into d_bases_out (16-byte aligned, writable up to the next
 * 16-byte multiple), ready for the *_dev entry points.  read_off may be NULL for n_reads reads of uniform_read_len
 * bases back to back (no offsets array to build or scan); uniform_read_len is ignored otherwise.               */
int nthash_kmer_batch_packed2bit(const uint8_t* packed, const uint32_t* invalid_bits, const uint64_t* read_off,
                                 uint64_t n_reads, uint32_t uniform_read_len, uint32_t k, uint32_t num_hashes,
                                 uint64_t* out, uint32_t* valid_bits, int device);
int nthash_kmer_reduce_packed2bit(const uint8_t* packed, const uint32_t* invalid_bits, const uint64_t* read_off,
                                  uint64_t n_reads, uint32_t uniform_read_len, uint32_t k, uint32_t num_hashes,
                                  uint64_t* result, int device);
/* Device-resident packed input hashed DIRECTLY (no ASCII copy in HBM): n_reads reads of read_len bases back to back, the
 * first one starting at base `first_base` (< 2^32) of the packed stream / bitmap.  d_packed must be 16-byte aligned and
 * readable up to the next multiple of 16 bytes past the last base, d_invalid_bits (nullable) up to the next multiple of
 * 8 bytes.  Served by the nibble-strip kernel for the shapes it takes (num_hashes <= 4, reads of at most ~250 bases whose
 * (read_len - k + 1) * num_hashes is a multiple of 8 — Illumina-like batches); returns NTHASH_ERR_UNSUPPORTED otherwise:
 * then expand with nthash_unpack2bit_dev and use nthash_kmer_batch_uniform_dev.  The host entries above make this choice
 * themselves per pipeline chunk.                                                                                      */
int nthash_kmer_batch_packed2bit_uniform_dev(const uint8_t* d_packed, const uint32_t* d_invalid_bits, uint64_t first_base,
                                             uint64_t n_reads, uint32_t read_len, uint32_t k, uint32_t num_hashes,
                                             uint64_t* d_out, uint32_t* d_valid_bits, void* stream);
int nthash_unpack2bit_dev(const uint8_t* d_packed, const uint32_t* d_invalid_bits, uint64_t first_base, uint64_t n_bases,
                          uint8_t* d_bases_out, void* stream);

/* ---- FASTQ staging on the device (the caller's side of the path) --------------------------------
 * FASTQ text already in device memory (plain four-line records, LF or CRLF) -> the concatenated layout the batch
 * entry points take: d_bases (sequence lines back to back) and d_read_off[n_reads + 1].  No host parsing: newline
 * positions come from a device-wide select, one warp copies each sequence line.  *n_reads / *n_bases (host) receive
 * the counts; the call synchronises the stream twice to size its passes.  Capacities: d_bases needs room for every
 * sequence byte (n_bytes / 2 always suffices), d_read_off for the number of records + 1.  Feed the result to
 * nthash_kmer_plan_dev + nthash_kmer_batch_dev (or the seed / consumer entries).                              */
int nthash_fastq_extract_dev(const uint8_t* d_text, uint64_t n_bytes, uint8_t* d_bases, uint64_t bases_capacity,
                             uint64_t* d_read_off, uint64_t reads_capacity, uint64_t* n_reads, uint64_t* n_bases,
                             void* stream);

/* ---- compacted output -------------------------------------------------------------------------------
 * The dense layout keeps a (zero) row for every window; the reference's loop only ever sees the windows it
 * visits.  This keeps exactly those rows (validity bit set), in order: d_compact[n][values_per_row], and
 * optionally their dense row numbers d_row_index[n] (row = koff[read] + get_pos(), so (read, position) follow
 * from koff); n goes to *d_count (device memory).  d_compact / d_row_index need room for `rows` entries at most.
 * Works on the output of the k-mer and the seed entry points alike.                                           */
int nthash_compact_rows_dev(const uint64_t* d_out, const uint32_t* d_valid_bits, uint64_t rows, uint32_t values_per_row,
                            uint64_t* d_compact, uint64_t* d_row_index, uint64_t* d_count, void* stream);

/* ---- fused consumer: Bloom filter (the caller the reference's header names, nthash.hpp:14-17: k-mer
 * hashes feeding Bloom filters) ------------------------------------------------------------------
 * For every window the reference's `while (h.roll())` loop visits, the num_hashes values of h.hashes()
 * (any 1..255, extend_hashes src/internal.hpp:104-118) address bits `hash % filter_bits` of a device-resident
 * filter; bit b is bit (b & 31) of 32-bit word b >> 5, i.e. bit (b & 7) of byte b >> 3.  query == 0 sets the
 * bits (atomically; the filter is the result and is order-independent), query != 0 only tests them.
 *   d_result[0] = windows visited
 *   d_result[1] = windows whose num_hashes bits were all set already (query: exact; insert: depends on the
 *                 order in which colliding k-mers of the same batch arrive)
 * d_result (3 x uint64, the third is 0) is zeroed by the call.  d_filter_words holds (filter_bits+31)/32 words. */
int nthash_kmer_bloom_uniform_dev(const uint8_t* d_bases, uint64_t n_bases_readable, uint64_t n_reads,
                                  uint32_t read_len, uint32_t k, uint32_t num_hashes, uint32_t* d_filter_words,
                                  uint64_t filter_bits, int query, uint64_t* d_result, void* stream);
int nthash_kmer_bloom_dev(const uint8_t* d_bases, uint64_t n_bases_readable, const uint64_t* d_read_off,
                          const uint64_t* d_koff, uint64_t n_reads, uint64_t max_read_len, uint32_t k,
                          uint32_t num_hashes, uint32_t* d_filter_words, uint64_t filter_bits, int query,
                          uint64_t* d_result, void* stream);

/* ---- consumer: minimizer selection ---------------------------------------------------------------------
 * The sketch most k-mer pipelines keep instead of all hashes.  A window is `window` consecutive k-mers of one read
 * (dense rows j .. j+window-1, 1 <= window <= 64); its minimizer is the k-mer with the smallest canonical hash
 * hashes()[0] among those the reference's loop visits, the leftmost one on ties; a window without a visited k-mer has
 * none; reads with fewer than `window` k-mers have no windows.  Bit (w & 31) of word (w >> 5) of min_bits is set iff
 * k-mer row w is the minimizer of at least one window ((rows + 31) / 32 words, zeroed here).  min_hash / min_row
 * (nullable, `capacity` entries each) receive the selected rows' hashes and dense row numbers in increasing row order;
 * *count = number of selected rows (entries past `capacity` are counted, not written).  About 2 / (window + 1) of the
 * rows are selected on random sequence, so that fraction of 8 bytes per k-mer is all that has to leave the device.  */
int nthash_kmer_minimizer_uniform_dev(const uint8_t* d_bases, uint64_t n_bases_readable, uint64_t n_reads,
                                      uint32_t read_len, uint32_t k, uint32_t window, uint32_t* d_min_bits,
                                      uint64_t* d_min_hash, uint64_t* d_min_row, uint64_t capacity, uint64_t* d_count,
                                      void* stream);
/* Host buffers; fixed-length reads when read_off == NULL (then uniform_read_len > 0), ragged otherwise. */
int nthash_kmer_minimizers(const char* bases, const uint64_t* read_off, uint64_t n_reads, uint32_t uniform_read_len,
                           uint32_t k, uint32_t window, uint32_t* min_bits, uint64_t* min_hash, uint64_t* min_row,
                           uint64_t capacity, uint64_t* count, int device);

/* ---- fused consumer: cardinality sketch in the style of ntCard (the k-mer counting tool built on ntHash that
 * nthash.hpp:56-57 points at).  Every window the reference's loop visits contributes its canonical hash h
 * (hashes()[0]): if the top `sample_bits` bits of h are zero (a 2^-sample_bits sample of the distinct k-mers), the
 * counter indexed by the next `index_bits` bits, d_counters[(h >> (64 - sample_bits - index_bits)) & (2^index_bits - 1)],
 * is incremented (uint32, atomically; NOT zeroed here so that batches accumulate).  The multiplicity histogram of that
 * table is what ntCard's estimator consumes (F0, f1, f2, ...).  d_result = {windows visited, windows sampled, 0}.
 * Nothing but the counters leaves the SM: the kernel runs at the speed of the reduce consumer.                     */
int nthash_kmer_sketch_uniform_dev(const uint8_t* d_bases, uint64_t n_bases_readable, uint64_t n_reads,
                                   uint32_t read_len, uint32_t k, uint32_t sample_bits, uint32_t index_bits,
                                   uint32_t* d_counters, uint64_t* d_result, void* stream);
int nthash_kmer_sketch_dev(const uint8_t* d_bases, uint64_t n_bases_readable, const uint64_t* d_read_off,
                           const uint64_t* d_koff, uint64_t n_reads, uint64_t max_read_len, uint32_t k,
                           uint32_t sample_bits, uint32_t index_bits, uint32_t* d_counters, uint64_t* d_result,
                           void* stream);

/* ---- SeedNtHash: spaced seeds ----------------------------------------------------------------
 * Replaces `nthash::SeedNtHash it(seq, len, seeds, h, k); while (it.roll()) use(it.hashes())`
 * (nthash.hpp:313-521; SeedNtHash::init/roll src/seed.cpp:493-544; ntmsm64 :130-270).
 * Rows hold n_seeds * num_hashes_per_seed values, seed-major (seed.cpp:167-171); out_fwd/out_rev
 * hold n_seeds values per row (get_forward_hash()/get_reverse_hash(), nthash.hpp:489-500).
 * SeedNtHash's own rules are reproduced exactly: non-ACGTU bytes inside a window are hashed, not
 * skipped (forward seed 0, reverse seed SEED_TAB[c & 7]); a non-ACGTU byte arriving as the incoming
 * base makes the iterator jump k positions (seed.cpp:527-530); init only rejects NUL bytes at
 * block positions (seed.cpp:151).                                                              */

/* Compiled seed set, resident on the current device.  Replaces the seed handling of the
 * SeedNtHash constructors (seed.cpp:449-491: check_seeds :85-104, get_blocks :19-66).  Seeds are
 * k-character strings of '1' (care) and '0' (don't care).  A length mismatch is an error
 * (seed.cpp:90-95); an asymmetric seed is accepted, as in the reference, which only warns
 * (seed.cpp:96-102) — query it with nthash_seed_plan_symmetric().                              */
typedef struct nthash_seed_plan nthash_seed_plan;
int nthash_seed_plan_create(const char* const* seeds, uint32_t n_seeds, uint32_t k,
                            uint32_t num_hashes_per_seed, nthash_seed_plan** plan_out);
void nthash_seed_plan_destroy(nthash_seed_plan* plan);
int nthash_seed_plan_symmetric(const nthash_seed_plan* plan);
/* Which kernel serves uniform batches for this plan: the one specialised for the seed set at plan
 * creation (NVRTC), or — with the reason — the generic one.  Both are CUDA; results are identical. */
const char* nthash_seed_plan_kernel_note(const nthash_seed_plan* plan);
/* Build check usable without a GPU: generates and compiles (sm_100a) the specialised kernel. */
int nthash_seed_jit_selftest(const char* const* seeds, uint32_t n_seeds, uint32_t k, uint32_t num_hashes_per_seed);

int nthash_seed_batch_uniform_dev(const nthash_seed_plan* plan, const uint8_t* d_bases,
                                  uint64_t n_bases_readable, uint64_t n_reads, uint32_t read_len,
                                  uint64_t* d_out, uint32_t* d_valid_bits, uint64_t* d_out_fwd,
                                  uint64_t* d_out_rev, void* stream);
/* SeedNtHash over a planned ragged batch (plan made with k = the seed length); enqueue only, like the k-mer form. */
int nthash_seed_batch_planned_dev(const nthash_seed_plan* seeds, const nthash_ragged_plan* plan, const uint8_t* d_bases,
                                  uint64_t n_bases_readable, uint64_t* d_out, uint32_t* d_valid_bits,
                                  uint64_t* d_out_fwd, uint64_t* d_out_rev, void* stream);
/* Ragged reads: d_koff / max_read_len from nthash_kmer_plan_dev(..., k = seed length, ...). */
int nthash_seed_batch_dev(const nthash_seed_plan* plan, const uint8_t* d_bases, uint64_t n_bases_readable,
                          const uint64_t* d_read_off, const uint64_t* d_koff, uint64_t n_reads,
                          uint64_t max_read_len, uint64_t* d_out, uint32_t* d_valid_bits,
                          uint64_t* d_out_fwd, uint64_t* d_out_rev, void* stream);
/* Host buffers in/out; compiles the seeds, runs, and frees the plan. */
int nthash_seed_batch(const char* bases, const uint64_t* read_off, uint64_t n_reads,
                      const char* const* seeds, uint32_t n_seeds, uint32_t k, uint32_t num_hashes_per_seed,
                      uint64_t* out, uint32_t* valid_bits, uint64_t* out_fwd, uint64_t* out_rev, int device);

/* SeedNtHash consumer: {windows the reference's loop visits, sum, xor of all their n_seeds * num_hashes_per_seed values}
 * without any hash crossing PCIe.  Two passes on the device (the seed kernels write a chunk of rows to scratch memory,
 * a reduction reads them back) - not yet fused like the k-mer consumers.  The device form zeroes d_result itself.   */
int nthash_seed_reduce(const char* bases, const uint64_t* read_off, uint64_t n_reads, const char* const* seeds,
                       uint32_t n_seeds, uint32_t k, uint32_t num_hashes_per_seed, uint64_t* result, int device);
int nthash_seed_reduce_uniform_dev(const nthash_seed_plan* plan, const uint8_t* d_bases, uint64_t n_bases_readable,
                                   uint64_t n_reads, uint32_t read_len, uint64_t* d_result, void* stream);

/* ---- BlindNtHash: caller-fed rolling over many independent states -----------------------
 * Replaces `blind.roll(char_in)` / `blind.peek(char_in)` (nthash.hpp:213-311; BlindNtHash::roll
 * src/kmer.cpp:355-364, ::peek :382-393) applied to n states at once, e.g. all frontier nodes of
 * a de Bruijn graph traversal.  The reference object keeps a deque of its k-mer only to know the
 * outgoing (first) base; here a state is (fwd[i], rev[i]) and the caller supplies out_base[i].
 * Initial states: hash the n k-mers with nthash_kmer_batch_uniform_dev(read_len = k) asking for
 * out_fwd/out_rev (the BlindNtHash constructor, kmer.cpp:338-353).  As in the reference there is
 * no validity check on the incoming byte.                                                     */

/* roll(in_base[i]): fwd/rev are updated in place, d_out[i*num_hashes + j] receives hashes(). */
int nthash_blind_roll_batch_dev(uint64_t* d_fwd, uint64_t* d_rev, const uint8_t* d_out_base,
                                const uint8_t* d_in_base, uint64_t n, uint32_t k, uint32_t num_hashes,
                                uint64_t* d_out, void* stream);
/* peek('A'),('C'),('G'),('T') of every state: d_out[(i*4 + e)*num_hashes + j]; states untouched. */
int nthash_blind_peek4_batch_dev(const uint64_t* d_fwd, const uint64_t* d_rev, const uint8_t* d_out_base,
                                 uint64_t n, uint32_t k, uint32_t num_hashes, uint64_t* d_out, void* stream);
/* BlindSeedNtHash::roll(char_in) on n states (nthash.hpp:537-632, src/seed.cpp:701-718).  A state is its k-mer, as in
 * the reference object (which keeps it in a deque): d_kmers holds n windows of k bytes back to back (16-byte aligned,
 * n_bytes_readable >= n*k); every window drops its first base and takes d_in_base[i], then d_out[i][n_seeds*h] gets
 * hashes() (seed-major) and the optional d_out_fwd/d_out_rev [i][n_seeds] the per-seed strand hashes.  Initial
 * hashes of the windows: nthash_seed_batch_uniform_dev with read_len = k.                                        */
int nthash_blind_seed_roll_batch_dev(const nthash_seed_plan* plan, uint8_t* d_kmers, uint64_t n_bytes_readable,
                                     const uint8_t* d_in_base, uint64_t n, uint64_t* d_out, uint64_t* d_out_fwd,
                                     uint64_t* d_out_rev, void* stream);
/* Host buffers in/out. */
int nthash_blind_roll_batch(uint64_t* fwd, uint64_t* rev, const char* out_base, const char* in_base,
                            uint64_t n, uint32_t k, uint32_t num_hashes, uint64_t* out, int device);

#ifdef __cplusplus
}
#endif
#endif
