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
// Short sequences never go to the GPU: a sequence with at most NTHASH_B200_HOST_CUTOFF windows (default 32768; a GPU
// round trip costs about as much as hashing that many windows on one core) is rolled on the host with the same
// arithmetic, window by window, exactly like the reference does — so a loop that builds one object per 150 bp read runs
// at the reference's speed instead of paying a kernel launch per read.  Longer sequences are hashed on the GPU in
// bounded chunks (a few million windows; host memory per object stays below ~100 MB whatever the sequence length) on
// device NTHASH_B200_DEVICE (default 0).
//
// Deviations, all documented in DESIGN.md §8: SeedNtHash::roll_back()/peek_back() return the true
// previous window (the reference's are off by one for seeds with monomers); after
// roll() returns false get_pos() stays on the last visited window; k < 3 is rejected (the reference segfaults).
#pragma once

#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <future>
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
  bool strands = false; // fwd / rev filled (only once the caller has asked for get_forward_hash() / get_reverse_hash())
  std::vector<uint64_t> out, fwd, rev;
  std::vector<uint32_t> valid;
  bool has(size_t w) const { return w >= w0 && w < w1; }
  bool ok(size_t w) const { return valid[(w - w0) >> 5] >> ((w - w0) & 31) & 1u; }
};
// The chunk behind the one being iterated is hashed by a helper thread meanwhile (the GPU call and its copies hide behind
// the caller's own loop over the current chunk); `key` says which chunk the pending result is.
struct Prefetch
{
  std::future<std::shared_ptr<WindowCache>> fut;
  size_t key = ~size_t(0);
  bool strands = false;
  std::shared_ptr<WindowCache> take(size_t want, bool want_strands)
  {
    std::shared_ptr<WindowCache> c;
    if (fut.valid()) {
      c = fut.get();
      if (key != want || (want_strands && !strands)) c.reset();
    }
    key = ~size_t(0);
    return c;
  }
};

constexpr size_t CHUNK = size_t(1) << 22; // windows hashed per GPU call by NtHash
// SeedNtHash chunks: a multiple of 244 x 32 windows (the batch engine then cuts a chunk into 244-window items, which its
// specialised kernel needs), sized so that one chunk's hashes + strand hashes stay below ~64 MB of host memory
inline size_t seed_chunk_rows(size_t values_per_window)
{
  const size_t unit = 244u * 32u, want = (size_t(64) << 20) / (8 * values_per_window);
  return std::max<size_t>(unit, want / unit * unit);
}
inline size_t env_size(const char* name, size_t dflt)
{
  const char* e = std::getenv(name);
  return e && *e ? (size_t)std::strtoull(e, nullptr, 10) : dflt;
}
inline size_t host_cutoff()
{
  static const size_t v = env_size("NTHASH_B200_HOST_CUTOFF", size_t(1) << 15);
  return v;
}
inline int device()
{
  static const int v = (int)env_size("NTHASH_B200_DEVICE", 0);
  return v;
}
inline bool prefetch_enabled()
{
  static const bool v = env_size("NTHASH_B200_SHIM_PREFETCH", 1) != 0;
  return v;
}

// byte -> seed (SEED_TAB, src/internal.hpp:132-165) and byte -> base index A,C,G,T = 0..3 (4: not hashable by NtHash)
struct ByteTables
{
  uint64_t seed[256];
  uint8_t idx[256];
};
inline const ByteTables& byte_tables()
{
  static const ByteTables t = [] {
    ByteTables x{};
    for (unsigned c = 0; c < 256; ++c) {
      x.seed[c] = seed_of((unsigned char)c);
      x.idx[c] = 4;
    }
    const char* names = "AaCcGgTtUu";
    const uint8_t ids[10] = { 0, 0, 1, 1, 2, 2, 3, 3, 3, 3 };
    for (int i = 0; i < 10; ++i) x.idx[(unsigned char)names[i]] = ids[i];
    return x;
  }();
  return t;
}

// Host evaluation of spaced-seed windows: per (k, seeds) tables of every care position's rotated seeds, shared by all
// objects of a thread (building them costs more than hashing one short read)
struct SeedHostTables
{
  unsigned k = 0;
  std::vector<std::string> seeds;
  std::vector<std::vector<unsigned>> care;       // care positions per seed
  std::vector<std::vector<uint64_t>> fwd, rev;   // [seed][care index * 256 + byte]
};
inline std::shared_ptr<const SeedHostTables> seed_host_tables(const std::vector<std::string>& seeds, unsigned k)
{
  thread_local std::vector<std::shared_ptr<const SeedHostTables>> memo;
  for (const auto& t : memo)
    if (t->k == k && t->seeds == seeds) return t;
  auto t = std::make_shared<SeedHostTables>();
  t->k = k;
  t->seeds = seeds;
  for (const auto& sd : seeds) {
    t->care.emplace_back();
    for (unsigned q = 0; q < k; ++q)
      if (sd[q] == '1') t->care.back().push_back(q);
    const auto& cp = t->care.back();
    t->fwd.emplace_back(cp.size() * 256);
    t->rev.emplace_back(cp.size() * 256);
    for (size_t j = 0; j < cp.size(); ++j) {
      for (unsigned c = 0; c < 256; ++c) {
        t->fwd.back()[j * 256 + c] = seed_of((unsigned char)c) ? roln(seed_of((unsigned char)c), k - 1 - cp[j]) : 0;
        t->rev.back()[j * 256 + c] = seed_of((unsigned char)(c & 7)) ? roln(seed_of((unsigned char)(c & 7)), cp[j]) : 0;
      }
    }
  }
  if (memo.size() >= 8) memo.erase(memo.begin());
  memo.push_back(t);
  return t;
}

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
    host_mode = seq_len - k + 1 <= detail::host_cutoff();
    if (host_mode) { // srol^k of the four seeds (what srol_table(c, k) returns, src/internal.hpp:343-348); index 4 = no seed
      for (int x = 0; x < 4; ++x) rk[x] = detail::roln(detail::kSeed[x], k);
      rk[4] = 0;
    }
  }
  NtHash(const std::string& seq, typedefs::NUM_HASHES_TYPE num_hashes, typedefs::K_TYPE k, size_t pos = 0)
    : NtHash(seq.data(), seq.size(), num_hashes, k, pos)
  {
  }
  NtHash(const NtHash& o)
    : seq(o.seq), num_hashes(o.num_hashes), k(o.k), pos(o.pos), initialized(o.initialized), host_mode(o.host_mode),
      want_strands(o.want_strands), fwd_hash(o.fwd_hash), rev_hash(o.rev_hash), hash_arr(new uint64_t[o.num_hashes ? o.num_hashes : 1]),
      cache(o.cache)
  {
    std::memcpy(rk, o.rk, sizeof rk);
    std::memcpy(hash_arr.get(), o.hash_arr.get(), num_hashes * sizeof(uint64_t));
  }
  NtHash(NtHash&&) = default;

  bool roll()
  {
    if (!host_mode) {
      if (initialized && cache) { // the common step: the next window of the cached chunk is a hashable one
        const size_t w = pos + 1;
        if (w < cache->w1 && w > cache->w0 && cache->ok(w)) {
          take(w);
          return true;
        }
      }
      return seek(initialized ? pos + 1 : pos, +1);
    }
    // the reference's own control flow (src/kmer.cpp:246-264), on the host
    if (!initialized) return host_init();
    if (pos >= seq.size() - k) return false;
    const unsigned char c_in = seq[pos + k], c_out = seq[pos];
    const detail::ByteTables& t = detail::byte_tables();
    if (t.idx[c_in] == 4) {
      pos += k;
      return host_init();
    }
    fwd_hash = detail::rol1(fwd_hash) ^ detail::kSeed[t.idx[c_in]] ^ rk[t.idx[c_out]];
    rev_hash = detail::ror1(rev_hash ^ rk[3 - t.idx[c_in]] ^ detail::kSeed[3 - t.idx[c_out]]);
    detail::extend(fwd_hash, rev_hash, k, num_hashes, hash_arr.get());
    ++pos;
    return true;
  }
  bool roll_back()
  {
    if (host_mode) { // src/kmer.cpp:266-288
      if (!initialized) return host_init();
      if (pos == 0) return false;
      const detail::ByteTables& t = detail::byte_tables();
      const unsigned char c_in = seq[pos - 1], c_out = seq[pos + k - 1];
      if (t.idx[c_in] == 4) {
        if (pos < k) return false;
        pos -= k;
        return host_init();
      }
      fwd_hash = detail::ror1(fwd_hash ^ rk[t.idx[c_in]] ^ detail::kSeed[t.idx[c_out]]);
      rev_hash = detail::rol1(rev_hash) ^ detail::kSeed[3 - t.idx[c_in]] ^ rk[3 - t.idx[c_out]];
      detail::extend(fwd_hash, rev_hash, k, num_hashes, hash_arr.get());
      --pos;
      return true;
    }
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
    if (!initialized) return first();
    if (!detail::hashable(char_in)) return false;
    need_strands();
    uint64_t f = fwd_hash, r = rev_hash;
    detail::step_fwd(f, r, k, seq[pos], char_in);
    detail::extend(f, r, k, num_hashes, hash_arr.get());
    return true;
  }
  bool peek_back(char char_in)
  {
    if (!initialized) return first();
    if (!detail::hashable(char_in)) return false;
    need_strands();
    uint64_t f = fwd_hash, r = rev_hash;
    detail::step_back(f, r, k, seq[pos + k - 1], char_in);
    detail::extend(f, r, k, num_hashes, hash_arr.get());
    return true;
  }

  const uint64_t* hashes() const { return hash_arr.get(); }
  size_t get_pos() const { return pos; }
  typedefs::NUM_HASHES_TYPE get_hash_num() const { return num_hashes; }
  typedefs::K_TYPE get_k() const { return k; }
  // GPU path: the strand hashes are fetched from the first time somebody asks for them (two more arrays to copy back)
  uint64_t get_forward_hash() const
  {
    const_cast<NtHash*>(this)->need_strands();
    return fwd_hash;
  }
  uint64_t get_reverse_hash() const
  {
    const_cast<NtHash*>(this)->need_strands();
    return rev_hash;
  }

private:
  std::string_view seq;
  typedefs::NUM_HASHES_TYPE num_hashes;
  typedefs::K_TYPE k;
  size_t pos;
  bool initialized = false;
  bool host_mode = false;       // short sequence: rolled on the host, window by window
  bool want_strands = false;    // GPU path: chunks carry the strand hashes too
  uint64_t rk[5] = { 0, 0, 0, 0, 0 };
  uint64_t fwd_hash = 0, rev_hash = 0;
  std::unique_ptr<uint64_t[]> hash_arr;
  std::shared_ptr<detail::WindowCache> cache; // shared between copies until one of them needs another chunk
  detail::Prefetch ahead;                     // the next chunk, being hashed while this one is iterated

  bool first() { return host_mode ? host_init() : seek(pos, +1); }
  // NtHash::init (src/kmer.cpp:228-244): the first window at or after pos without an invalid base; base hashes by
  // Horner's rule over its k bases (the closed form of base_forward_hash / base_reverse_hash, kmer.cpp:43-73, :123-152)
  bool host_init()
  {
    const detail::ByteTables& t = detail::byte_tables();
    const size_t n = seq.size();
    for (;;) {
      if (pos + k > n) return false;
      size_t bad = k;
      for (size_t i = k; i-- > 0;)
        if (t.idx[(unsigned char)seq[pos + i]] == 4) {
          bad = i;
          break;
        }
      if (bad == k) break;
      pos += bad + 1;
    }
    uint64_t f = 0, r = 0;
    for (size_t i = 0; i < k; ++i) {
      f = detail::rol1(f) ^ detail::kSeed[t.idx[(unsigned char)seq[pos + i]]];
      r = detail::rol1(r) ^ detail::kSeed[3 - t.idx[(unsigned char)seq[pos + k - 1 - i]]];
    }
    fwd_hash = f;
    rev_hash = r;
    detail::extend(f, r, k, num_hashes, hash_arr.get());
    initialized = true;
    return true;
  }
  void need_strands()
  {
    if (host_mode || want_strands) return;
    want_strands = true;
    if (initialized) { // the current chunk came without them: fetch it again
      load_chunk(pos, 0);
      take(pos);
    }
  }
  void take(size_t w)
  {
    const size_t i = w - cache->w0;
    const uint64_t* src = cache->out.data() + i * num_hashes;
    uint64_t* dst = hash_arr.get();
    for (unsigned j = 0; j < num_hashes; ++j) dst[j] = src[j];
    if (want_strands) {
      fwd_hash = cache->fwd[i];
      rev_hash = cache->rev[i];
    }
    pos = w;
  }
  // windows [c * CHUNK, (c + 1) * CHUNK) of the sequence through the CUDA engine (one batch call: H2D, kernel, D2H)
  static std::shared_ptr<detail::WindowCache> hash_chunk(std::string_view seq, unsigned k, unsigned h, size_t c_idx, bool strands)
  {
    const size_t n_win = seq.size() - k + 1;
    auto c = std::make_shared<detail::WindowCache>();
    c->w0 = c_idx * detail::CHUNK;
    c->w1 = std::min(n_win, c->w0 + detail::CHUNK);
    c->strands = strands;
    const size_t rows = c->w1 - c->w0;
    c->out.resize(rows * h);
    if (strands) {
      c->fwd.resize(rows);
      c->rev.resize(rows);
    }
    c->valid.resize((rows + 31) / 32);
    const uint64_t off[2] = { 0, rows + k - 1 };
    detail::check_abi(nthash_kmer_batch(seq.data() + c->w0, off, 1, k, h, c->out.data(), c->valid.data(),
                                        strands ? c->fwd.data() : nullptr, strands ? c->rev.data() : nullptr, detail::device()),
                      "NtHash");
    return c;
  }
  void load_chunk(size_t w, int dir)
  {
    const size_t c_idx = w / detail::CHUNK, n_win = seq.size() - k + 1;
    auto c = ahead.take(c_idx, want_strands);
    cache = c ? c : hash_chunk(seq, k, num_hashes, c_idx, want_strands);
    if (dir > 0 && (c_idx + 1) * detail::CHUNK < n_win && detail::prefetch_enabled()) { // rolling forward: start on the next chunk now
      const std::string_view sq = seq;
      const unsigned kk = k, hh = num_hashes;
      const bool st = want_strands;
      ahead.key = c_idx + 1;
      ahead.strands = st;
      ahead.fut = std::async(std::launch::async, [sq, kk, hh, c_idx, st] { return hash_chunk(sq, kk, hh, c_idx + 1, st); });
    }
  }
  // move to the nearest hashable window at or beyond `w` in direction dir; false if there is none
  bool seek(size_t w, int dir)
  {
    const size_t n_win = seq.size() - k + 1;
    for (;; w += dir) {
      if (w >= n_win) return false; // also catches w wrapping below zero
      if (!cache || !cache->has(w)) load_chunk(w, dir);
      if (cache->ok(w)) break;
    }
    take(w);
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
      cache(o.cache), origin(o.origin), on_orbit(o.on_orbit), want_strands(o.want_strands), strands_current(o.strands_current), host(o.host)
  {
    alloc();
    std::memcpy(fwd_hash.get(), o.fwd_hash.get(), seeds.size() * sizeof(uint64_t));
    std::memcpy(rev_hash.get(), o.rev_hash.get(), seeds.size() * sizeof(uint64_t));
    std::memcpy(hash_arr.get(), o.hash_arr.get(), get_hash_num() * sizeof(uint64_t));
  }
  SeedNtHash(SeedNtHash&&) = default;

  // SeedNtHash::init / roll (src/seed.cpp:493-544).  The cache holds the windows of one chunk together with the bitmap
  // of the positions the reference visits when it starts at the chunk's first window (`origin`); as long as this object
  // got where it is by rolling forward from there it only has to walk that bitmap.
  bool roll()
  {
    if (seq.size() < k || pos > seq.size() - k) return false; // the reference's init() finds nothing to hash there (seed.cpp:497-512)
    if (!initialized) return land(pos);
    if (pos >= seq.size() - k) return false;
    const bool step1 = detail::byte_tables().seed[(unsigned char)seq[pos + k]] != 0;
    const size_t nxt = step1 ? pos + 1 : pos + k; // an invalid incoming base: k positions on, re-init
    if (nxt > seq.size() - k) return false;
    if (on_orbit && cache && nxt - origin < cache->w1) {
      if (cache->ok(nxt - origin)) { // the common step
        take(nxt - origin);
        return true;
      }
      return walk(nxt - origin);
    }
    return land(nxt);
  }
  bool roll_back() // src/seed.cpp:546-575, with the true previous window (see the deviations above)
  {
    if (seq.size() < k || pos > seq.size() - k) return false;
    if (!initialized) return land(pos);
    if (pos == 0) return false;
    if (detail::seed_of(seq[pos - 1]) == 0) return pos >= k ? land(pos - k) : false;
    --pos;
    if (cache && pos >= origin && pos - origin < cache->w1 && cache->ok(pos - origin)) {
      take(pos - origin);
    } else { // a window the cached visiting order skips (or none cached): one window, closed form on the host
      fill_from_window(std::string(seq.substr(pos, k)), true);
      on_orbit = false;
    }
    return true;
  }
  bool peek() { return pos >= seq.size() - k ? false : peek(seq[pos + k]); }
  bool peek_back() { return pos == 0 ? false : peek_back(seq[pos - 1]); }
  bool peek(char char_in)
  {
    if (!initialized) return roll();
    std::string w(seq.substr(pos + 1, k - 1));
    w.push_back(char_in);
    fill_from_window(w, false);
    return true;
  }
  bool peek_back(char char_in)
  {
    if (!initialized) return roll();
    std::string w(1, char_in);
    w.append(seq.substr(pos, k - 1));
    fill_from_window(w, false);
    return true;
  }

  const uint64_t* hashes() const { return hash_arr.get(); }
  size_t get_pos() const { return pos; }
  unsigned get_hash_num() const { return num_hashes_per_seed * seeds.size(); }
  typedefs::NUM_HASHES_TYPE get_hash_num_per_seed() const { return num_hashes_per_seed; }
  typedefs::K_TYPE get_k() const { return k; }
  // GPU path: the per-seed strand hashes are fetched from the first time somebody asks for them (the returned pointers
  // stay valid and are refreshed by every later roll, as the reference's are)
  uint64_t* get_forward_hash() const
  {
    const_cast<SeedNtHash*>(this)->need_strands();
    return fwd_hash.get();
  }
  uint64_t* get_reverse_hash() const
  {
    const_cast<SeedNtHash*>(this)->need_strands();
    return rev_hash.get();
  }

private:
  std::string_view seq;
  typedefs::NUM_HASHES_TYPE num_hashes_per_seed;
  typedefs::K_TYPE k;
  size_t pos;
  bool initialized = false;
  std::vector<std::string> seeds;
  std::unique_ptr<uint64_t[]> fwd_hash, rev_hash, hash_arr;
  std::shared_ptr<detail::WindowCache> cache; // windows [origin, origin + w1) and the reference's visiting order from `origin`
  size_t origin = 0;
  bool on_orbit = false;                       // pos was reached by rolling forward from `origin`
  bool want_strands = false;                   // GPU path: chunks carry the strand hashes too
  bool strands_current = false;                // fwd_hash / rev_hash belong to the current window
  std::shared_ptr<const detail::SeedHostTables> host; // short sequences: evaluated on the host
  detail::Prefetch ahead;                      // the chunk behind the cached one, being hashed meanwhile

  void alloc()
  {
    fwd_hash.reset(new uint64_t[seeds.size()]());
    rev_hash.reset(new uint64_t[seeds.size()]());
    hash_arr.reset(new uint64_t[std::max<size_t>(1, get_hash_num())]());
  }
  void take(size_t i)
  {
    const size_t m = seeds.size(), H = get_hash_num();
    const uint64_t* src = cache->out.data() + i * H;
    uint64_t* dst = hash_arr.get();
    for (size_t j = 0; j < H; ++j) dst[j] = src[j];
    strands_current = cache->strands;
    if (cache->strands) {
      std::memcpy(fwd_hash.get(), cache->fwd.data() + i * m, m * sizeof(uint64_t));
      std::memcpy(rev_hash.get(), cache->rev.data() + i * m, m * sizeof(uint64_t));
    }
    pos = origin + i;
    initialized = true;
  }
  void need_strands()
  {
    want_strands = true;
    if (!initialized || strands_current) return;
    // the current window came from a chunk without strand hashes: one window by the closed form (later chunks carry them)
    for (size_t s = 0; s < seeds.size(); ++s) detail::closed_form(seq.substr(pos, k), k, &seeds[s], fwd_hash[s], rev_hash[s]);
    strands_current = true;
    if (cache && !cache->strands && on_orbit && pos >= origin && pos - origin < cache->w1) { // and the rest of this chunk again, with them
      const size_t at = pos;
      load(origin);
      take(at - origin);
    }
  }
  // next visited window at or after cache row i; crosses into the next chunk when the bitmap runs out
  bool walk(size_t i)
  {
    for (; i < cache->w1; ++i)
      if (cache->ok(i)) {
        take(i);
        return true;
      }
    // nothing left in this chunk: init() continues scanning at the first window behind it
    return land(origin + cache->w1);
  }
  // init() at window w: hash a chunk that starts there; its first visited row is where the reference lands
  bool land(size_t w)
  {
    for (;;) {
      if (w > seq.size() - k) return false;
      load(w);
      on_orbit = true;
      for (size_t i = 0; i < cache->w1; ++i)
        if (cache->ok(i)) {
          take(i);
          return true;
        }
      w = origin + cache->w1; // a whole chunk of windows init() skips (NUL bytes): keep scanning
    }
  }
  // `rows` windows from window `w` on through the CUDA engine, with the reference's visiting order when it starts at w
  static std::shared_ptr<detail::WindowCache> hash_chunk(std::string_view seq, const std::vector<std::string>& seeds, unsigned k,
                                                          unsigned hps, size_t w, size_t rows, bool strands)
  {
    auto c = std::make_shared<detail::WindowCache>();
    const size_t m = seeds.size(), H = hps * m;
    c->w0 = 0;
    c->w1 = rows;
    c->strands = strands;
    c->out.resize(rows * H);
    if (strands) {
      c->fwd.resize(rows * m);
      c->rev.resize(rows * m);
    }
    c->valid.assign((rows + 31) / 32, 0u);
    std::vector<const char*> sp;
    for (const auto& s : seeds) sp.push_back(s.c_str());
    const uint64_t off[2] = { 0, rows + k - 1 };
    detail::check_abi(nthash_seed_batch(seq.data() + w, off, 1, sp.data(), (uint32_t)m, k, hps, c->out.data(), c->valid.data(),
                                        strands ? c->fwd.data() : nullptr, strands ? c->rev.data() : nullptr, detail::device()),
                      "SeedNtHash");
    return c;
  }
  void load(size_t w)
  {
    origin = w;
    const size_t m = seeds.size(), H = get_hash_num(), left = seq.size() - origin - k + 1;
    const bool on_host = seq.size() - k + 1 <= detail::host_cutoff() && !std::memchr(seq.data(), 0, seq.size());
    const size_t chunk = detail::seed_chunk_rows(H + 2 * m), rows = on_host ? left : std::min(left, chunk);
    if (!on_host) {
      auto pre = ahead.take(w, want_strands);
      cache = pre ? pre : hash_chunk(seq, seeds, k, num_hashes_per_seed, w, rows, want_strands);
      if (rows < left && detail::prefetch_enabled()) { // more behind this chunk: start on it now
        const std::string_view sq = seq;
        const std::vector<std::string> sd = seeds;
        const unsigned kk = k, hh = num_hashes_per_seed;
        const size_t w2 = w + rows, rows2 = std::min(left - rows, chunk);
        const bool st = want_strands;
        ahead.key = w2;
        ahead.strands = st;
        ahead.fut = std::async(std::launch::async, [sq, sd, kk, hh, w2, rows2, st] { return hash_chunk(sq, sd, kk, hh, w2, rows2, st); });
      }
      return;
    }
    auto c = std::make_shared<detail::WindowCache>();
    c->w0 = 0;
    c->w1 = rows;
    c->strands = true;
    c->out.resize(rows * H);
    c->fwd.resize(rows * m);
    c->rev.resize(rows * m);
    c->valid.assign((rows + 31) / 32, 0u);
    {
      if (!host) host = detail::seed_host_tables(seeds, k);
      // every window's value (seed.cpp:149-171: only care positions contribute, a byte without a seed contributes 0 forward
      // and SEED_TAB[c & 7] reverse) + the visiting order: from a visited window, one step if the incoming base has a seed,
      // else k (seed.cpp:524-530).  Sequences holding a NUL byte take the GPU path, which also knows init()'s NUL rule.
      const unsigned char* sq = (const unsigned char*)seq.data() + origin;
      for (size_t i = 0; i < rows;) {
        for (size_t sd = 0; sd < m; ++sd) {
          const auto& cp = host->care[sd];
          const uint64_t *tf = host->fwd[sd].data(), *tr = host->rev[sd].data();
          uint64_t f = 0, r = 0;
          for (size_t j = 0; j < cp.size(); ++j) {
            const unsigned c8 = sq[i + cp[j]];
            f ^= tf[j * 256 + c8];
            r ^= tr[j * 256 + c8];
          }
          c->fwd[i * m + sd] = f;
          c->rev[i * m + sd] = r;
          detail::extend(f, r, k, num_hashes_per_seed, c->out.data() + i * H + sd * num_hashes_per_seed);
        }
        c->valid[i >> 5] |= 1u << (i & 31);
        if (i + 1 >= rows) break;
        i += detail::seed_of(sq[i + k]) != 0 ? 1 : k;
      }
    }
    cache = c;
  }
  void fill_from_window(const std::string& w, bool with_strands)
  {
    for (size_t s = 0; s < seeds.size(); ++s) {
      uint64_t f, r;
      detail::closed_form(w, k, &seeds[s], f, r);
      detail::extend(f, r, k, num_hashes_per_seed, hash_arr.get() + s * num_hashes_per_seed);
      if (with_strands) {
        fwd_hash[s] = f;
        rev_hash[s] = r;
      }
    }
    if (with_strands) strands_current = true;
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
