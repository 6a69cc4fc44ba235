/*
 * nthash_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the ntHash v2 rolling-hash path of the reference
 * (bcgsc/ntHash 2.4.0: src/internal.hpp, src/kmer.cpp, src/seed.cpp).  It is
 * the checker for the CUDA engine in nthash_b200/; only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may call
 * it.  The product path (nthash_b200/) never links or loads it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   - every golden vector the reference's own tests hold for the path
 *     (tests/tests.cpp:54-57, :193-200, :236-240 of the reference), and
 *   - the unmodified reference compiled from /root/reference into
 *     oracle/_ref/libnthash_ref.so (oracle/Makefile), on randomized inputs.
 *
 * The `nto_*` functions here and the `ntr_*` functions of oracle/ref_driver.cpp
 * share signatures so tests can swap one for the other.
 */
#ifndef NTHASH_ORACLE_H
#define NTHASH_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- primitives (exposed so tests can pin them one by one) ------------- */
uint64_t nto_srol(uint64_t x);                 /* internal.hpp:41-47  */
uint64_t nto_sror(uint64_t x);                 /* internal.hpp:83-88  */
uint64_t nto_srol_n(uint64_t x, unsigned d);   /* d-fold srol, any d  */
uint64_t nto_seed(unsigned char c);            /* SEED_TAB, internal.hpp:132-165 */
uint64_t nto_srol_table(unsigned char c, unsigned d); /* internal.hpp:343-348 */
void nto_extend_hashes(uint64_t fwd, uint64_t rev, unsigned k, unsigned h,
                       uint64_t* out);         /* internal.hpp:104-118 */

/* ---- per-read iterators, run to exhaustion ------------------------------
 * Each emits what `while (obj.roll())` would visit, in order:
 *   pos_out[i]            = get_pos()
 *   hash_out[i*H .. +H)   = hashes()         (H = h, or n_seeds*h seed-major)
 *   fwd_out / rev_out     = get_forward_hash()/get_reverse_hash()
 *                           (one per emission for k-mers, n_seeds per emission
 *                           for spaced seeds).  Any output pointer may be NULL.
 * Returns the number of emissions (never writes more than `cap`).
 * Returns (size_t)-1 where the reference constructor would raise_error().   */
size_t nto_kmer_read(const char* seq, size_t len, unsigned k, unsigned h,
                     size_t pos0, uint64_t* pos_out, uint64_t* hash_out,
                     uint64_t* fwd_out, uint64_t* rev_out, size_t cap);

size_t nto_seed_read(const char* seq, size_t len, const char* const* seeds,
                     unsigned n_seeds, unsigned h_per_seed, unsigned k,
                     size_t pos0, uint64_t* pos_out, uint64_t* hash_out,
                     uint64_t* fwd_out, uint64_t* rev_out, size_t cap);

/* BlindNtHash: kmer.cpp:338-364.  Starts from seq[0..k), feeds `n_in`
 * incoming characters, records hashes()/fwd/rev after every roll(char).
 * hash_out is [n_in][h]; hash0_out[h] (nullable) gets the constructor's.    */
void nto_blind_read(const char* kmer, unsigned k, unsigned h,
                    const char* chars_in, size_t n_in, uint64_t* hash0_out,
                    uint64_t* hash_out, uint64_t* fwd_out, uint64_t* rev_out);

/* Seed pattern -> (blocks, monomers) exactly as seed.cpp:19-66 picks them.
 * blocks_out holds pairs [start,end); returns 0 on success.                 */
int nto_get_blocks(const char* seed, unsigned* blocks_out, unsigned* n_blocks,
                   unsigned* monos_out, unsigned* n_monos, unsigned cap);

/* ---- batch forms in the engine's dense layout ---------------------------
 * Reads are back to back in `bases`; read r is bases[read_off[r]..read_off[r+1]).
 * kmer row offset koff[r] = sum_{r'<r} max(0, len_r' - k + 1).
 * out[(koff[r]+p)*H + j] = hash j of the window starting at p (0 where the
 * reference would not emit p); valid[koff[r]+p] = 1/0 (one byte per window);
 * out_fwd/out_rev per window (per window per seed for spaced seeds).
 * Every pointer except bases/read_off may be NULL.  n_threads >= 1 splits the
 * read range over that many pthreads (the reference itself is single-threaded).
 * sum_out/xor_out (nullable): 64-bit sum / xor over every emitted hash value.
 * Returns the number of emitted windows.                                    */
uint64_t nto_kmer_batch(const char* bases, const uint64_t* read_off,
                        uint64_t n_reads, unsigned k, unsigned h,
                        uint64_t* out, uint8_t* valid, uint64_t* out_fwd,
                        uint64_t* out_rev, int n_threads, uint64_t* sum_out,
                        uint64_t* xor_out);

uint64_t nto_seed_batch(const char* bases, const uint64_t* read_off,
                        uint64_t n_reads, const char* const* seeds,
                        unsigned n_seeds, unsigned h_per_seed, unsigned k,
                        uint64_t* out, uint8_t* valid, uint64_t* out_fwd,
                        uint64_t* out_rev, int n_threads, uint64_t* sum_out,
                        uint64_t* xor_out);

/* Deterministic synthetic reads: splitmix64
 * stream seeded with `seed`, 32 bases per draw, "ACGT"[r & 3], low bits first. */
void nto_gen_bases(char* dst, uint64_t n, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
