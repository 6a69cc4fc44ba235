// <nthash/nthash.hpp> of the B200 engine: the reference's C++ iterator API re-exposed on top of the batch
// C ABI (include/nthash_b200.h).  Link with -lnthash_b200; nothing else changes for a caller of
// bcgsc/ntHash 2.4.0 (namespace, class names, constructor and method signatures, error behaviour).
//
// How it maps (reference file:line in /root/reference):
//   NtHash::roll()        include/nthash/nthash.hpp:123, src/kmer.cpp:246-264
//       first call hashes the whole remaining sequence on the GPU (nthash_kmer_batch) in chunks of
//       CHUNK windows; every roll() then only advances to the next row whose validity bit is set and
//       copies its hashes into the object's own array (hashes() stays a stable pointer, as callers
//       expect: tests/tests.cpp:423).
//   SeedNtHash::roll()    include/nthash/nthash.hpp:428, src/seed.cpp:518-544     -> nthash_seed_batch
//   roll_back()/peek*()   single steps from the current (fwd, rev) state; they are latency-bound
//       by nature (graph traversal) and are evaluated on the host with the same split-rotate
//       arithmetic (detail:: below).  Traversal engines that step many states at once use
//       nthash_blind_roll_batch_dev / nthash_blind_peek4_batch_dev instead.
//   BlindNtHash / BlindSeedNtHash  (nthash.hpp:213-311, :537-632): caller-fed single steps, host side.
//   raise_error()         src/internal.hpp:16-22: message on stderr + exit(1), reproduced by detail::die().
//
// Deviations, all documented in DESIGN.md §7: SeedNtHash::roll_back()/peek_back() return the true
// previous window (the reference's are off by one for seeds with monomers); after
// roll() returns false get_pos() stays on the last visited window.
#pragma once

#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <iostream>
#include <memory>
#include <string>
#include <string_view>
#include <sys/types.h>
#include <vector>

#include "../nthash_b200.h"

namespace nthash {

static const char* const NTHASH_FN_NAME = "ntHash_v2";

namespace typedefs {
using NUM_HASHES_TYPE = uint8_t;
using K_TYPE = uint16_t;
using SpacedSeedBlocks = std::vector<std::array<unsigned, 2>>;
using SpacedSeedMonomers = std::vector<unsigned>;
} // namespace typedefs

namespace detail {

[[noreturn]] inline void die(const char* cls, const std::string& msg)
{
  std::cerr << "[ntHash::" << cls << "] \33[31mERROR: \33[0m" << msg << std::endl;
  std::exit(1);
}
inline void warn(const char* cls, const std::string& msg)
{
  std::cerr << "[ntHash::" << cls << "] \33[33mWARNING: \33[0m" << msg << std::endl;
}
inline void check_abi(int rc, const char* cls)
{
  if (rc != NTHASH_OK) die(cls, std::string("nthash_b200: ") + nthash_last_error());
}

// ---- single-step arithmetic (host side of roll_back / peek / Blind*) ----
constexpr uint64_t kSeed[4] = { 0x3c8bfbb395c60474ULL, 0x3193c18562a02b4cULL, 0x20323ed082572324ULL, 0x295549f54be24456ULL };
inline uint64_t seed_of(unsigned char c)
{
  switch (c) {
    case 'A': case 'a': case 4: case 5: return kSeed[0];
    case 'C': case 'c': case 7: return kSeed[1];
    case 'G': case 'g': case 3: return kSeed[2];
    case 'T': case 't': case 'U': case 'u': case 1: return kSeed[3];
    default: return 0;
  }
}
inline uint64_t rol1(uint64_t x) { return ((x << 1) & 0xFFFFFFFDFFFFFFFFULL) | ((x >> 63) << 33) | ((x >> 32) & 1ULL); }
inline uint64_t ror1(uint64_t x) { return ((x >> 1) & 0x7FFFFFFEFFFFFFFFULL) | (((x >> 33) & 1ULL) << 63) | ((x & 1ULL) << 32); }
inline uint64_t roln(uint64_t x, unsigned d)
{
  const uint64_t m33 = (1ULL << 33) - 1, m31 = (1ULL << 31) - 1;
  uint64_t lo = x & m33, hi = x >> 33;
  const unsigned a = d % 33, b = d % 31;
  if (a) lo = ((lo << a) | (lo >> (33 - a))) & m33;
  if (b) hi = ((hi << b) | (hi >> (31 - b))) & m31;
  return (hi << 33) | lo;
}
inline void extend(uint64_t fwd, uint64_t rev, unsigned k, unsigned h, uint64_t* out)
{
  out[0] = fwd + rev;
  for (unsigned i = 1; i < h; ++i) {
    uint64_t t = out[0] * ((uint64_t)i ^ ((uint64_t)k * 0x90b45d39fb6da1faULL));
    out[i] = t ^ (t >> 27);
  }
}
// one step forward / backward of the contiguous k-mer recurrence
inline void step_fwd(uint64_t& f, uint64_t& r, unsigned k, unsigned char c_out, unsigned char c_in)
{
  f = rol1(f) ^ seed_of(c_in) ^ roln(seed_of(c_out), k);
  r = ror1(r ^ roln(seed_of(c_in & 7), k) ^ seed_of(c_out & 7));
}
inline void step_back(uint64_t& f, uint64_t& r, unsigned k, unsigned char c_out, unsigned char c_in)
{
  f = ror1(f ^ roln(seed_of(c_in), k) ^ seed_of(c_out));
  r = rol1(r) ^ seed_of(c_in & 7) ^ roln(seed_of(c_out & 7), k);
}
// closed form of one window under a care mask (empty mask = all positions)
template<class Window>
inline void closed_form(const Window& w, unsigned k, const std::string* mask, uint64_t& f, uint64_t& r)
{
  f = r = 0;
  for (unsigned q = 0; q < k; ++q) {
    if (mask && (*mask)[q] != '1') continue;
    const unsigned char c = (unsigned char)w[q];
    f ^= roln(seed_of(c), k - 1 - q);
    r ^= roln(seed_of(c & 7), q);
  }
}
inline bool hashable(unsigned char c) { return seed_of(c) != 0 && c > 7; }

inline std::vector<std::string> seeds_from_parsed(const std::vector<std::vector<unsigned>>& parsed, unsigned k)
{
  std::vector<std::string> out;
  for (const auto& s : parsed) {
    std::string m(k, '1');
    for (unsigned i : s) m[i] = '0';
    out.push_back(m);
  }
  return out;
}
inline void check_seed_strings(const std::vector<std::string>& seeds, unsigned k, const char* cls)
{
  for (const auto& s : seeds) {
    if (s.size() != k)
      die(cls, "Spaced seed string length (" + std::to_string(s.size()) + ") not equal to k=" + std::to_string(k) + " in " + s);
    if (!std::equal(s.begin(), s.end(), s.rbegin()))
      warn(cls, "Seed " + s + " is not symmetric, reverse-complement hashing will be inconsistent");
  }
}

// GPU results for windows [w0, w1) of one sequence, H hashes and S strand values per window
struct WindowCache
{
  size_t w0 = 0, w1 = 0;
  std::vector<uint64_t> out, fwd, rev;
  std::vector<uint32_t> valid;
  bool has(size_t w) const { return w >= w0 && w < w1; }
  bool ok(size_t w) const { return valid[(w - w0) >> 5] >> ((w - w0) & 31) & 1u; }
};

constexpr size_t CHUNK = size_t(1) << 22; // windows hashed per GPU call by NtHash

} // namespace detail

inline std::vector<std::vector<unsigned>> parse_seeds(const std::vector<std::string>& seed_strings)
{
  std::vector<std::vector<unsigned>> out;
  for (const auto& s : seed_strings) {
    out.emplace_back();
    for (unsigned i = 0; i < s.size(); ++i)
      if (s[i] != '1') out.back().push_back(i);
  }
  return out;
}

// ------------------------------------------------------------------------------------- NtHash
class NtHash
{
public:
  NtHash(const char* seq, size_t seq_len, typedefs::NUM_HASHES_TYPE num_hashes, typedefs::K_TYPE k, size_t pos = 0)
    : seq(seq, seq_len), num_hashes(num_hashes), k(k), pos(pos), hash_arr(new uint64_t[num_hashes ? num_hashes : 1])
  {
    if (k == 0) detail::die("NtHash", "k must be greater than 0");
    if (seq_len < k)
      detail::die("NtHash", "sequence length (" + std::to_string(seq_len) + ") is smaller than k (" + std::to_string(k) + ")");
    if (pos > seq_len - k)
      detail::die("NtHash", "passed position (" + std::to_string(pos) + ") is larger than sequence length (" + std::to_string(seq_len) + ")");
  }
  NtHash(const std::string& seq, typedefs::NUM_HASHES_TYPE num_hashes, typedefs::K_TYPE k, size_t pos = 0)
    : NtHash(seq.data(), seq.size(), num_hashes, k, pos)
  {
  }
  NtHash(const NtHash& o)
    : seq(o.seq), num_hashes(o.num_hashes), k(o.k), pos(o.pos), initialized(o.initialized), fwd_hash(o.fwd_hash),
      rev_hash(o.rev_hash), hash_arr(new uint64_t[o.num_hashes ? o.num_hashes : 1]), cache(o.cache)
  {
    std::memcpy(hash_arr.get(), o.hash_arr.get(), num_hashes * sizeof(uint64_t));
  }
  NtHash(NtHash&&) = default;

  bool roll() { return seek(initialized ? pos + 1 : pos, +1); }
  bool roll_back()
  {
    if (!initialized) return seek(pos, +1);
    if (pos == 0) return false;
    // across an invalid base the reference re-initialises k positions back and scans forward (kmer.cpp:274-277)
    if (!detail::hashable(seq[pos - 1])) return pos >= k ? seek(pos - k, +1) : false;
    return seek(pos - 1, -1);
  }
  bool peek() { return pos >= seq.size() - k ? false : peek(seq[pos + k]); }
  bool peek_back() { return pos == 0 ? false : peek_back(seq[pos - 1]); }
  bool peek(char char_in)
  {
    if (!initialized) return seek(pos, +1);
    if (!detail::hashable(char_in)) return false;
    uint64_t f = fwd_hash, r = rev_hash;
    detail::step_fwd(f, r, k, seq[pos], char_in);
    detail::extend(f, r, k, num_hashes, hash_arr.get());
    return true;
  }
  bool peek_back(char char_in)
  {
    if (!initialized) return seek(pos, +1);
    if (!detail::hashable(char_in)) return false;
    uint64_t f = fwd_hash, r = rev_hash;
    detail::step_back(f, r, k, seq[pos + k - 1], char_in);
    detail::extend(f, r, k, num_hashes, hash_arr.get());
    return true;
  }

  const uint64_t* hashes() const { return hash_arr.get(); }
  size_t get_pos() const { return pos; }
  typedefs::NUM_HASHES_TYPE get_hash_num() const { return num_hashes; }
  typedefs::K_TYPE get_k() const { return k; }
  uint64_t get_forward_hash() const { return fwd_hash; }
  uint64_t get_reverse_hash() const { return rev_hash; }

private:
  std::string_view seq;
  typedefs::NUM_HASHES_TYPE num_hashes;
  typedefs::K_TYPE k;
  size_t pos;
  bool initialized = false;
  uint64_t fwd_hash = 0, rev_hash = 0;
  std::unique_ptr<uint64_t[]> hash_arr;
  std::shared_ptr<detail::WindowCache> cache; // shared between copies until one of them needs another chunk

  void load_chunk(size_t w)
  {
    const size_t n_win = seq.size() - k + 1;
    auto c = std::make_shared<detail::WindowCache>();
    c->w0 = w / detail::CHUNK * detail::CHUNK;
    c->w1 = std::min(n_win, c->w0 + detail::CHUNK);
    const size_t rows = c->w1 - c->w0;
    c->out.resize(rows * num_hashes);
    c->fwd.resize(rows);
    c->rev.resize(rows);
    c->valid.resize((rows + 31) / 32);
    const uint64_t off[2] = { 0, rows + k - 1 };
    detail::check_abi(nthash_kmer_batch(seq.data() + c->w0, off, 1, k, num_hashes, c->out.data(), c->valid.data(),
                                        c->fwd.data(), c->rev.data(), 0),
                      "NtHash");
    cache = c;
  }
  // move to the nearest hashable window at or beyond `w` in direction dir; false if there is none
  bool seek(size_t w, int dir)
  {
    const size_t n_win = seq.size() - k + 1;
    for (;; w += dir) {
      if (w >= n_win) return false; // also catches w wrapping below zero
      if (!cache || !cache->has(w)) load_chunk(w);
      if (cache->ok(w)) break;
    }
    const size_t i = w - cache->w0;
    std::memcpy(hash_arr.get(), cache->out.data() + i * num_hashes, num_hashes * sizeof(uint64_t));
    fwd_hash = cache->fwd[i];
    rev_hash = cache->rev[i];
    pos = w;
    initialized = true;
    return true;
  }
};

// -------------------------------------------------------------------------------- BlindNtHash
class BlindNtHash
{
public:
  BlindNtHash(const char* seq, typedefs::NUM_HASHES_TYPE num_hashes, typedefs::K_TYPE k, ssize_t pos = 0)
    : seq(seq + pos, seq + pos + k), num_hashes(num_hashes), pos(pos), hash_arr(new uint64_t[num_hashes ? num_hashes : 1])
  {
    if (k == 0) detail::die("BlindNtHash", "k must be greater than 0");
    // the reference hashes seq[0..k) whatever `pos` is (kmer.cpp:350-351); kept for drop-in behaviour
    detail::closed_form(seq, k, nullptr, fwd_hash, rev_hash);
    detail::extend(fwd_hash, rev_hash, k, num_hashes, hash_arr.get());
  }
  BlindNtHash(const BlindNtHash& o)
    : seq(o.seq), num_hashes(o.num_hashes), pos(o.pos), fwd_hash(o.fwd_hash), rev_hash(o.rev_hash),
      hash_arr(new uint64_t[o.num_hashes ? o.num_hashes : 1])
  {
    std::memcpy(hash_arr.get(), o.hash_arr.get(), num_hashes * sizeof(uint64_t));
  }
  BlindNtHash(BlindNtHash&&) = default;

  void roll(char char_in)
  {
    detail::step_fwd(fwd_hash, rev_hash, seq.size(), seq.front(), char_in);
    detail::extend(fwd_hash, rev_hash, seq.size(), num_hashes, hash_arr.get());
    seq.pop_front();
    seq.push_back(char_in);
    ++pos;
  }
  void roll_back(char char_in)
  {
    detail::step_back(fwd_hash, rev_hash, seq.size(), seq.back(), char_in);
    detail::extend(fwd_hash, rev_hash, seq.size(), num_hashes, hash_arr.get());
    seq.pop_back();
    seq.push_front(char_in);
    --pos;
  }
  void peek(char char_in)
  {
    uint64_t f = fwd_hash, r = rev_hash;
    detail::step_fwd(f, r, seq.size(), seq.front(), char_in);
    detail::extend(f, r, seq.size(), num_hashes, hash_arr.get());
  }
  void peek_back(char char_in)
  {
    uint64_t f = fwd_hash, r = rev_hash;
    detail::step_back(f, r, seq.size(), seq.back(), char_in);
    detail::extend(f, r, seq.size(), num_hashes, hash_arr.get());
  }

  const uint64_t* hashes() const { return hash_arr.get(); }
  ssize_t get_pos() const { return pos; }
  typedefs::NUM_HASHES_TYPE get_hash_num() const { return num_hashes; }
  typedefs::K_TYPE get_k() const { return seq.size(); }
  uint64_t get_forward_hash() const { return fwd_hash; }
  uint64_t get_reverse_hash() const { return rev_hash; }

private:
  std::deque<char> seq;
  typedefs::NUM_HASHES_TYPE num_hashes;
  ssize_t pos;
  uint64_t fwd_hash = 0, rev_hash = 0;
  std::unique_ptr<uint64_t[]> hash_arr;
};

// --------------------------------------------------------------------------------- SeedNtHash
class SeedNtHash
{
public:
  SeedNtHash(const char* seq, size_t seq_len, const std::vector<std::string>& seeds,
             typedefs::NUM_HASHES_TYPE num_hashes_per_seed, typedefs::K_TYPE k, size_t pos = 0)
    : seq(seq, seq_len), num_hashes_per_seed(num_hashes_per_seed), k(k), pos(pos), seeds(seeds)
  {
    detail::check_seed_strings(seeds, k, "SeedNtHash");
    alloc();
  }
  SeedNtHash(const std::string& seq, const std::vector<std::string>& seeds, typedefs::NUM_HASHES_TYPE num_hashes_per_seed,
             typedefs::K_TYPE k, size_t pos = 0)
    : SeedNtHash(seq.data(), seq.size(), seeds, num_hashes_per_seed, k, pos)
  {
  }
  SeedNtHash(const char* seq, size_t seq_len, const std::vector<std::vector<unsigned>>& seeds,
             typedefs::NUM_HASHES_TYPE num_hashes_per_seed, typedefs::K_TYPE k, size_t pos = 0)
    : seq(seq, seq_len), num_hashes_per_seed(num_hashes_per_seed), k(k), pos(pos), seeds(detail::seeds_from_parsed(seeds, k))
  {
    alloc();
  }
  SeedNtHash(const std::string& seq, const std::vector<std::vector<unsigned>>& seeds,
             typedefs::NUM_HASHES_TYPE num_hashes_per_seed, typedefs::K_TYPE k, size_t pos = 0)
    : SeedNtHash(seq.data(), seq.size(), seeds, num_hashes_per_seed, k, pos)
  {
  }
  SeedNtHash(const SeedNtHash& o)
    : seq(o.seq), num_hashes_per_seed(o.num_hashes_per_seed), k(o.k), pos(o.pos), initialized(o.initialized), seeds(o.seeds),
      cache(o.cache), origin(o.origin)
  {
    alloc();
    std::memcpy(fwd_hash.get(), o.fwd_hash.get(), seeds.size() * sizeof(uint64_t));
    std::memcpy(rev_hash.get(), o.rev_hash.get(), seeds.size() * sizeof(uint64_t));
    std::memcpy(hash_arr.get(), o.hash_arr.get(), get_hash_num() * sizeof(uint64_t));
  }
  SeedNtHash(SeedNtHash&&) = default;

  bool roll() { return seek(initialized ? pos + 1 : pos, +1); }
  bool roll_back()
  {
    if (!initialized) return seek(pos, +1);
    return pos == 0 ? false : seek(pos - 1, -1);
  }
  bool peek() { return pos >= seq.size() - k ? false : peek(seq[pos + k]); }
  bool peek_back() { return pos == 0 ? false : peek_back(seq[pos - 1]); }
  bool peek(char char_in)
  {
    if (!initialized) return seek(pos, +1);
    std::string w(seq.substr(pos + 1, k - 1));
    w.push_back(char_in);
    fill_from_window(w);
    return true;
  }
  bool peek_back(char char_in)
  {
    if (!initialized) return seek(pos, +1);
    std::string w(1, char_in);
    w.append(seq.substr(pos, k - 1));
    fill_from_window(w);
    return true;
  }

  const uint64_t* hashes() const { return hash_arr.get(); }
  size_t get_pos() const { return pos; }
  unsigned get_hash_num() const { return num_hashes_per_seed * seeds.size(); }
  typedefs::NUM_HASHES_TYPE get_hash_num_per_seed() const { return num_hashes_per_seed; }
  typedefs::K_TYPE get_k() const { return k; }
  uint64_t* get_forward_hash() const { return fwd_hash.get(); }
  uint64_t* get_reverse_hash() const { return rev_hash.get(); }

private:
  std::string_view seq;
  typedefs::NUM_HASHES_TYPE num_hashes_per_seed;
  typedefs::K_TYPE k;
  size_t pos;
  bool initialized = false;
  std::vector<std::string> seeds;
  std::unique_ptr<uint64_t[]> fwd_hash, rev_hash, hash_arr;
  std::shared_ptr<detail::WindowCache> cache; // windows of seq[origin..): the visiting order depends on the start
  size_t origin = 0;

  void alloc()
  {
    fwd_hash.reset(new uint64_t[seeds.size()]());
    rev_hash.reset(new uint64_t[seeds.size()]());
    hash_arr.reset(new uint64_t[std::max<size_t>(1, get_hash_num())]());
  }
  void load()
  {
    if (seq.size() < k) detail::die("SeedNtHash", "sequence length is smaller than k");
    origin = pos;
    auto c = std::make_shared<detail::WindowCache>();
    const size_t rows = seq.size() - origin - k + 1, m = seeds.size(), H = get_hash_num();
    c->w0 = 0;
    c->w1 = rows;
    c->out.resize(rows * H);
    c->fwd.resize(rows * m);
    c->rev.resize(rows * m);
    c->valid.resize((rows + 31) / 32);
    std::vector<const char*> sp;
    for (const auto& s : seeds) sp.push_back(s.c_str());
    const uint64_t off[2] = { 0, seq.size() - origin };
    detail::check_abi(nthash_seed_batch(seq.data() + origin, off, 1, sp.data(), (uint32_t)m, k, num_hashes_per_seed,
                                        c->out.data(), c->valid.data(), c->fwd.data(), c->rev.data(), 0),
                      "SeedNtHash");
    cache = c;
  }
  bool seek(size_t w, int dir)
  {
    if (!cache) load();
    if (w < origin) return false;
    for (size_t i = w - origin;; i += dir) {
      if (i >= cache->w1) return false;
      if (cache->ok(i)) {
        const size_t m = seeds.size(), H = get_hash_num();
        std::memcpy(hash_arr.get(), cache->out.data() + i * H, H * sizeof(uint64_t));
        std::memcpy(fwd_hash.get(), cache->fwd.data() + i * m, m * sizeof(uint64_t));
        std::memcpy(rev_hash.get(), cache->rev.data() + i * m, m * sizeof(uint64_t));
        pos = origin + i;
        initialized = true;
        return true;
      }
    }
  }
  void fill_from_window(const std::string& w)
  {
    for (size_t s = 0; s < seeds.size(); ++s) {
      uint64_t f, r;
      detail::closed_form(w, k, &seeds[s], f, r);
      detail::extend(f, r, k, num_hashes_per_seed, hash_arr.get() + s * num_hashes_per_seed);
    }
  }
};

// ---------------------------------------------------------------------------- BlindSeedNtHash
class BlindSeedNtHash
{
public:
  BlindSeedNtHash(const char* seq, const std::vector<std::string>& seeds, typedefs::NUM_HASHES_TYPE num_hashes_per_seed,
                  typedefs::K_TYPE k, ssize_t pos = 0)
    : seq(seq + pos, seq + pos + k), num_hashes_per_seed(num_hashes_per_seed), k(k), pos(pos), seeds(seeds),
      fwd_hash(new uint64_t[seeds.size()]()), rev_hash(new uint64_t[seeds.size()]()),
      hash_arr(new uint64_t[std::max<size_t>(1, num_hashes_per_seed * seeds.size())]())
  {
    detail::check_seed_strings(seeds, k, "SeedNtHash");
    refresh();
  }
  BlindSeedNtHash(const BlindSeedNtHash& o)
    : seq(o.seq), num_hashes_per_seed(o.num_hashes_per_seed), k(o.k), pos(o.pos), seeds(o.seeds),
      fwd_hash(new uint64_t[o.seeds.size()]()), rev_hash(new uint64_t[o.seeds.size()]()),
      hash_arr(new uint64_t[std::max<size_t>(1, o.num_hashes_per_seed * o.seeds.size())]())
  {
    std::memcpy(fwd_hash.get(), o.fwd_hash.get(), seeds.size() * sizeof(uint64_t));
    std::memcpy(rev_hash.get(), o.rev_hash.get(), seeds.size() * sizeof(uint64_t));
    std::memcpy(hash_arr.get(), o.hash_arr.get(), get_hash_num() * sizeof(uint64_t));
  }
  BlindSeedNtHash(BlindSeedNtHash&&) = default;

  void roll(char char_in)
  {
    seq.pop_front();
    seq.push_back(char_in);
    ++pos;
    refresh();
  }
  void roll_back(char char_in)
  {
    seq.pop_back();
    seq.push_front(char_in);
    --pos;
    refresh();
  }

  const uint64_t* hashes() const { return hash_arr.get(); }
  ssize_t get_pos() const { return pos; }
  unsigned get_hash_num() const { return num_hashes_per_seed * seeds.size(); }
  typedefs::NUM_HASHES_TYPE get_hash_num_per_seed() const { return num_hashes_per_seed; }
  typedefs::K_TYPE get_k() const { return k; }
  uint64_t* get_forward_hash() const { return fwd_hash.get(); }
  uint64_t* get_reverse_hash() const { return rev_hash.get(); }

private:
  std::deque<char> seq;
  typedefs::NUM_HASHES_TYPE num_hashes_per_seed;
  typedefs::K_TYPE k;
  ssize_t pos;
  std::vector<std::string> seeds;
  std::unique_ptr<uint64_t[]> fwd_hash, rev_hash, hash_arr;

  void refresh() // the window's hash depends only on its bases and the care mask
  {
    for (size_t s = 0; s < seeds.size(); ++s) {
      detail::closed_form(seq, k, &seeds[s], fwd_hash[s], rev_hash[s]);
      detail::extend(fwd_hash[s], rev_hash[s], k, num_hashes_per_seed, hash_arr.get() + s * num_hashes_per_seed);
    }
  }
};

} // namespace nthash
