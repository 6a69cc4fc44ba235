// class_bench.cpp — the reference's user loop (`Hash h(seq, ...); while (h.roll()) sum += h.hashes()[i];`, the shape of
// examples/benchmark.cpp:31-39) over deterministic reads, compiled twice from this one file: against the reference's own
// header + sources (oracle/Makefile -> oracle/_ref/class_bench_ref) and against this repo's drop-in header + CUDA engine
// (-> oracle/_ref/class_bench_shim).  Both print one JSON line: seconds, windows visited, a checksum over every visited
// window and the peak resident set — so the two builds can be compared for speed, memory and bit-exactness.
//
//   class_bench kmer  <n_reads> <read_len> <k> <h>      NtHash     over n_reads reads (one object per read)
//   class_bench seed  <n_reads> <read_len> <h>          SeedNtHash over n_reads reads, the two 31-wide seeds of BASELINE configs[3]
// NTHASH_BENCH_PASSES=<n>: repeat the whole loop n times in this process; `seconds` is the first pass (which pays one-time
// costs such as creating the CUDA context), `seconds_best` the fastest one.
// NTHASH_BENCH_NO_STRANDS=1: leave get_forward_hash() / get_reverse_hash() out of the checksum (the loop of the reference's
// examples/benchmark.cpp only reads hashes()).
// NTHASH_BENCH_DIRTY=<per million>: that share of the bases is replaced by other bytes (N, n, Y, lower-case acgt, '.').
// Reads: one splitmix64 stream (seed 42), 32 bases per draw, "ACGT"[r & 3].
#include <nthash/nthash.hpp>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static std::string gen_bases(size_t n, uint64_t seed)
{
  std::string s(n, 'A');
  uint64_t state = seed;
  for (size_t i = 0; i < n; i += 32) {
    state += 0x9e3779b97f4a7c15ULL;
    uint64_t z = state;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    z ^= z >> 31;
    for (size_t j = 0; j < 32 && i + j < n; ++j, z >>= 2) s[i + j] = "ACGT"[z & 3];
  }
  return s;
}

int main(int argc, char** argv)
{
  if (argc < 5) {
    std::fprintf(stderr, "usage: class_bench kmer <n_reads> <read_len> <k> <h> | seed <n_reads> <read_len> <h>\n");
    return 2;
  }
  const bool seed_mode = std::strcmp(argv[1], "seed") == 0;
  const size_t n_reads = std::strtoull(argv[2], nullptr, 10), read_len = std::strtoull(argv[3], nullptr, 10);
  const unsigned k = seed_mode ? 31u : (unsigned)std::atoi(argv[4]);
  const unsigned h = (unsigned)std::atoi(argv[seed_mode ? 4 : 5]);
  const std::vector<std::string> seeds = { "1010101010101010101010101010101", "1101101101101101011011011011011" };
  std::string bases = gen_bases(n_reads * read_len, 42);
  if (const char* e = std::getenv("NTHASH_BENCH_DIRTY")) {
    const uint64_t ppm = std::strtoull(e, nullptr, 10);
    uint64_t st = 7;
    for (size_t i = 0; i < bases.size(); ++i) {
      st = st * 6364136223846793005ULL + 1442695040888963407ULL;
      if ((st >> 33) % 1000000 < ppm) bases[i] = "NnYacgt.RU"[(st >> 20) % 10];
    }
  }
  const char* pe = std::getenv("NTHASH_BENCH_PASSES");
  const int passes = pe ? std::max(1, std::atoi(pe)) : 1;
  const bool strands = std::getenv("NTHASH_BENCH_NO_STRANDS") == nullptr;
  uint64_t sum = 0, x = 0, visited = 0;
  double sec = 0, best = 1e30;
  for (int pass = 0; pass < passes; ++pass) {
  sum = x = visited = 0;
  const auto t0 = std::chrono::steady_clock::now();
  for (size_t r = 0; r < n_reads; ++r) {
    const char* sq = bases.data() + r * read_len;
    if (seed_mode) {
      nthash::SeedNtHash it(sq, read_len, seeds, (nthash::typedefs::NUM_HASHES_TYPE)h, (nthash::typedefs::K_TYPE)k);
      const unsigned H = h * (unsigned)seeds.size();
      while (it.roll()) {
        for (unsigned j = 0; j < H; ++j) {
          sum += it.hashes()[j];
          x ^= it.hashes()[j];
        }
        if (strands) sum += it.get_forward_hash()[1] ^ it.get_reverse_hash()[0];
        ++visited;
      }
    } else {
      nthash::NtHash it(sq, read_len, (nthash::typedefs::NUM_HASHES_TYPE)h, (nthash::typedefs::K_TYPE)k);
      while (it.roll()) {
        for (unsigned j = 0; j < h; ++j) {
          sum += it.hashes()[j];
          x ^= it.hashes()[j];
        }
        if (strands) sum += it.get_forward_hash() ^ it.get_reverse_hash();
        ++visited;
      }
    }
  }
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (pass == 0) sec = dt;
  best = std::min(best, dt);
  }
  // peak resident set of THIS program: VmHWM (getrusage's ru_maxrss also carries the parent's size across fork/exec)
  double rss_mb = 0;
  if (FILE* f = std::fopen("/proc/self/status", "r")) {
    char line[256];
    while (std::fgets(line, sizeof line, f))
      if (std::strncmp(line, "VmHWM:", 6) == 0) rss_mb = std::strtod(line + 6, nullptr) / 1024.0;
    std::fclose(f);
  }
  std::printf("{\"mode\": \"%s\", \"n_reads\": %zu, \"read_len\": %zu, \"k\": %u, \"h\": %u, \"seconds\": %.6f, \"seconds_best\": %.6f, \"windows\": %llu, "
              "\"windows_per_sec\": %.4g, \"windows_per_sec_best\": %.4g, \"sum\": \"%016llx\", \"xor\": \"%016llx\", \"max_rss_mb\": %.1f}\n",
              argv[1], n_reads, read_len, k, h, sec, best, (unsigned long long)visited, visited / sec, visited / best, (unsigned long long)sum,
              (unsigned long long)x, rss_mb);
  return 0;
}
