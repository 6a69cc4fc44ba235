/*
 * ref_driver.cpp — CPU ORACLE, strong form (test infrastructure, NOT product code).
 *
 * extern "C" harness around the UNMODIFIED reference (bcgsc/ntHash 2.4.0).
 * oracle/Makefile compiles the reference's own src/kmer.cpp and src/seed.cpp
 * where they lie under /root/reference and links them with this file into
 * oracle/_ref/libnthash_ref.so.  No reference source is copied into the repo;
 * this file only calls the reference's public classes
 * (include/nthash/nthash.hpp:62-211 NtHash, :213-311 BlindNtHash, :313-521 SeedNtHash).
 *
 * The `ntr_*` entry points mirror the `nto_*` ones of nthash_oracle.h, so tests
 * can check the plain-C restatement against the real thing, and bench.py can
 * time the reference's own roll() loop (loop shape of examples/benchmark.cpp:34-39)
 * on the host cores.
 */
#include <nthash/nthash.hpp>

#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Job
{
  uint64_t r0 = 0, r1 = 0, n_emit = 0, sum = 0, xr = 0;
};

template<class PerRead>
uint64_t
run_threads(uint64_t n_reads, int n_threads, uint64_t* sum_out, uint64_t* xor_out, PerRead&& body)
{
  if (n_threads < 1) n_threads = 1;
  if ((uint64_t)n_threads > n_reads) n_threads = n_reads ? (int)n_reads : 1;
  std::vector<Job> jobs(n_threads);
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; t++) {
    jobs[t].r0 = n_reads * (uint64_t)t / (uint64_t)n_threads;
    jobs[t].r1 = n_reads * (uint64_t)(t + 1) / (uint64_t)n_threads;
    auto fn = [&body, &jobs, t]() {
      Job& j = jobs[t];
      for (uint64_t r = j.r0; r < j.r1; r++) body(r, j);
    };
    if (n_threads == 1) fn();
    else th.emplace_back(fn);
  }
  for (auto& x : th) x.join();
  uint64_t n = 0, sum = 0, xr = 0;
  for (auto& j : jobs) { n += j.n_emit; sum += j.sum; xr ^= j.xr; }
  if (sum_out) *sum_out = sum;
  if (xor_out) *xor_out = xr;
  return n;
}

std::vector<uint64_t>
make_koff(const uint64_t* read_off, uint64_t n_reads, unsigned k)
{
  std::vector<uint64_t> koff(n_reads + 1);
  uint64_t acc = 0;
  for (uint64_t r = 0; r < n_reads; r++) {
    koff[r] = acc;
    uint64_t len = read_off[r + 1] - read_off[r];
    if (len >= k) acc += len - k + 1;
  }
  koff[n_reads] = acc;
  return koff;
}

} // namespace

extern "C" {

uint64_t ntr_kmer_strand(const char* kmer, unsigned k, int reverse)
{
  /* fwd / rev hash of one k-mer through the reference's init path */
  nthash::NtHash h(kmer, k, 1, (nthash::typedefs::K_TYPE)k);
  h.roll();
  return reverse ? h.get_reverse_hash() : h.get_forward_hash();
}

size_t ntr_kmer_read(const char* seq, size_t len, unsigned k, unsigned h, size_t pos0,
                     uint64_t* pos_out, uint64_t* hash_out, uint64_t* fwd_out, uint64_t* rev_out,
                     size_t cap)
{
  if (k == 0 || h == 0 || len < k || pos0 > len - k) return (size_t)-1; /* would exit(1) */
  nthash::NtHash it(seq, len, (uint8_t)h, (uint16_t)k, pos0);
  size_t n = 0;
  while (it.roll()) {
    if (n < cap) {
      if (pos_out) pos_out[n] = it.get_pos();
      if (hash_out) std::memcpy(hash_out + n * h, it.hashes(), sizeof(uint64_t) * h);
      if (fwd_out) fwd_out[n] = it.get_forward_hash();
      if (rev_out) rev_out[n] = it.get_reverse_hash();
    }
    n++;
  }
  return n;
}

size_t ntr_seed_read(const char* seq, size_t len, const char* const* seeds, unsigned n_seeds,
                     unsigned h, unsigned k, size_t pos0, uint64_t* pos_out, uint64_t* hash_out,
                     uint64_t* fwd_out, uint64_t* rev_out, size_t cap)
{
  if (k == 0 || h == 0 || n_seeds == 0 || len < k || pos0 > len - k) return (size_t)-1;
  std::vector<std::string> sv(seeds, seeds + n_seeds);
  for (auto& s : sv)
    if (s.size() != k) return (size_t)-1; /* would exit(1), seed.cpp:90-95 */
  nthash::SeedNtHash it(seq, len, sv, (uint8_t)h, (uint16_t)k, pos0);
  size_t n = 0, H = (size_t)n_seeds * h;
  while (it.roll()) {
    if (n < cap) {
      if (pos_out) pos_out[n] = it.get_pos();
      if (hash_out) std::memcpy(hash_out + n * H, it.hashes(), sizeof(uint64_t) * H);
      if (fwd_out) std::memcpy(fwd_out + n * n_seeds, it.get_forward_hash(), sizeof(uint64_t) * n_seeds);
      if (rev_out) std::memcpy(rev_out + n * n_seeds, it.get_reverse_hash(), sizeof(uint64_t) * n_seeds);
    }
    n++;
  }
  return n;
}

void ntr_blind_read(const char* kmer, unsigned k, unsigned h, const char* chars_in, size_t n_in,
                    uint64_t* hash0_out, uint64_t* hash_out, uint64_t* fwd_out, uint64_t* rev_out)
{
  nthash::BlindNtHash it(kmer, (uint8_t)h, (uint16_t)k, 0);
  if (hash0_out) std::memcpy(hash0_out, it.hashes(), sizeof(uint64_t) * h);
  for (size_t i = 0; i < n_in; i++) {
    it.roll(chars_in[i]);
    if (hash_out) std::memcpy(hash_out + i * h, it.hashes(), sizeof(uint64_t) * h);
    if (fwd_out) fwd_out[i] = it.get_forward_hash();
    if (rev_out) rev_out[i] = it.get_reverse_hash();
  }
}

uint64_t ntr_kmer_batch(const char* bases, const uint64_t* read_off, uint64_t n_reads, unsigned k,
                        unsigned h, uint64_t* out, uint8_t* valid, uint64_t* out_fwd,
                        uint64_t* out_rev, int n_threads, uint64_t* sum_out, uint64_t* xor_out)
{
  if (k == 0 || h == 0) return (uint64_t)-1;
  std::vector<uint64_t> koff;
  if (out || valid || out_fwd || out_rev) {
    koff = make_koff(read_off, n_reads, k);
    uint64_t tot = koff[n_reads];
    if (out) std::memset(out, 0, sizeof(uint64_t) * tot * h);
    if (valid) std::memset(valid, 0, tot);
    if (out_fwd) std::memset(out_fwd, 0, sizeof(uint64_t) * tot);
    if (out_rev) std::memset(out_rev, 0, sizeof(uint64_t) * tot);
  }
  const uint64_t* kp = koff.empty() ? nullptr : koff.data();
  return run_threads(n_reads, n_threads, sum_out, xor_out, [&](uint64_t r, Job& j) {
    size_t len = (size_t)(read_off[r + 1] - read_off[r]);
    if (len < k) return;
    nthash::NtHash it(bases + read_off[r], len, (uint8_t)h, (uint16_t)k);
    while (it.roll()) {
      const uint64_t* hv = it.hashes();
      for (unsigned q = 0; q < h; q++) { j.sum += hv[q]; j.xr ^= hv[q]; }
      j.n_emit++;
      if (kp) {
        uint64_t o = kp[r] + it.get_pos();
        if (out) std::memcpy(out + o * h, hv, sizeof(uint64_t) * h);
        if (valid) valid[o] = 1;
        if (out_fwd) out_fwd[o] = it.get_forward_hash();
        if (out_rev) out_rev[o] = it.get_reverse_hash();
      }
    }
  });
}

uint64_t ntr_seed_batch(const char* bases, const uint64_t* read_off, uint64_t n_reads,
                        const char* const* seeds, unsigned n_seeds, unsigned h, unsigned k,
                        uint64_t* out, uint8_t* valid, uint64_t* out_fwd, uint64_t* out_rev,
                        int n_threads, uint64_t* sum_out, uint64_t* xor_out)
{
  if (k == 0 || h == 0 || n_seeds == 0) return (uint64_t)-1;
  std::vector<std::string> sv(seeds, seeds + n_seeds);
  for (auto& s : sv)
    if (s.size() != k) return (uint64_t)-1;
  const size_t H = (size_t)n_seeds * h;
  std::vector<uint64_t> koff;
  if (out || valid || out_fwd || out_rev) {
    koff = make_koff(read_off, n_reads, k);
    uint64_t tot = koff[n_reads];
    if (out) std::memset(out, 0, sizeof(uint64_t) * tot * H);
    if (valid) std::memset(valid, 0, tot);
    if (out_fwd) std::memset(out_fwd, 0, sizeof(uint64_t) * tot * n_seeds);
    if (out_rev) std::memset(out_rev, 0, sizeof(uint64_t) * tot * n_seeds);
  }
  const uint64_t* kp = koff.empty() ? nullptr : koff.data();
  return run_threads(n_reads, n_threads, sum_out, xor_out, [&](uint64_t r, Job& j) {
    size_t len = (size_t)(read_off[r + 1] - read_off[r]);
    if (len < k) return;
    /* one object per read, seeds re-parsed each time: the reference's own usage */
    nthash::SeedNtHash it(bases + read_off[r], len, sv, (uint8_t)h, (uint16_t)k);
    while (it.roll()) {
      const uint64_t* hv = it.hashes();
      for (size_t q = 0; q < H; q++) { j.sum += hv[q]; j.xr ^= hv[q]; }
      j.n_emit++;
      if (kp) {
        uint64_t o = kp[r] + it.get_pos();
        if (out) std::memcpy(out + o * H, hv, sizeof(uint64_t) * H);
        if (valid) valid[o] = 1;
        if (out_fwd) std::memcpy(out_fwd + o * n_seeds, it.get_forward_hash(), sizeof(uint64_t) * n_seeds);
        if (out_rev) std::memcpy(out_rev + o * n_seeds, it.get_reverse_hash(), sizeof(uint64_t) * n_seeds);
      }
    }
  });
}

const char* ntr_fn_name() { return nthash::NTHASH_FN_NAME; }

} // extern "C"
