// Exercises the drop-in C++ classes of include/nthash/nthash.hpp (this repo's, GPU-backed) with the
// same scenarios the reference checks in its tests/tests.cpp (block numbers below refer to
// SURVEY.md §4).  Restated, not copied: same sequences and expected values, own harness.
// Usage: shim_tests [--host-only | --short-only]   (--host-only: the Blind* blocks; --short-only: everything but the
// multi-megabase sequence — with the default NTHASH_B200_HOST_CUTOFF none of that needs a GPU)
#include <nthash/nthash.hpp>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

static int failures = 0;
#define CHECK(cond)                                                          \
  do {                                                                       \
    if (!(cond)) {                                                           \
      std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);   \
      ++failures;                                                            \
    }                                                                        \
  } while (0)

static bool same(const uint64_t* a, const uint64_t* b, unsigned n) { return std::memcmp(a, b, n * sizeof(uint64_t)) == 0; }
static std::vector<uint64_t> snap(const uint64_t* p, unsigned n) { return std::vector<uint64_t>(p, p + n); }

static void host_only_blocks()
{
  { // block 14: BlindSeedNtHash follows SeedNtHash (values checked against the GPU in gpu_blocks); here self-consistency
    std::string seq = "ATGCTAGTAGCTGAC";
    std::vector<std::string> seeds = { "110011", "101101" };
    nthash::BlindSeedNtHash a(seq.data(), seeds, 3, 6);
    nthash::BlindSeedNtHash fresh(seq.data() + 1, seeds, 3, 6);
    a.roll(seq[6]);
    CHECK(same(a.hashes(), fresh.hashes(), 6));
    a.roll_back(seq[0]);
    nthash::BlindSeedNtHash first(seq.data(), seeds, 3, 6);
    CHECK(same(a.hashes(), first.hashes(), 6) && a.get_pos() == 0);
  }
  { // block 16: copy constructor keeps rolling identically
    std::string seq = "ATGCTAGTAGCTGAC";
    std::vector<std::string> seeds = { "110011", "101101" };
    nthash::BlindSeedNtHash h1(seq.data(), seeds, 1, 6);
    h1.roll('A');
    h1.roll('C');
    nthash::BlindSeedNtHash h2(h1);
    CHECK(same(h1.hashes(), h2.hashes(), 2));
    h1.roll('G');
    h2.roll('G');
    CHECK(same(h1.hashes(), h2.hashes(), 2));
  }
  { // BlindNtHash golden values (reference tests.cpp:54-57, second half of block 1)
    std::string seq = "ACATGCATGCA";
    const uint64_t want[2][3] = { { 0x38cc00f940aebdaeULL, 0xab7e1b110e086fc6ULL, 0x11a1818bcfdd553ULL },
                                  { 0x603a48c5a11c794aULL, 0xe66016e61816b9c4ULL, 0xc5b13cb146996ffeULL } };
    nthash::BlindNtHash blind(seq.data(), 3, 5);
    for (auto& w : want) {
      blind.roll(seq[blind.get_pos() + 5]);
      CHECK(same(w, blind.hashes(), 3));
    }
    auto cur = snap(blind.hashes(), 3);
    blind.peek('T');
    auto peeked = snap(blind.hashes(), 3);
    blind.roll('T');
    CHECK(same(peeked.data(), blind.hashes(), 3));
    blind.roll_back(seq[2]);
    CHECK(same(cur.data(), blind.hashes(), 3) && blind.get_k() == 5 && blind.get_hash_num() == 3);
  }
  CHECK(std::string(nthash::NTHASH_FN_NAME) == "ntHash_v2");
  auto ps = nthash::parse_seeds({ "1101", "0110" });
  CHECK(ps.size() == 2 && ps[0] == std::vector<unsigned>({ 2 }) && ps[1] == std::vector<unsigned>({ 0, 3 }));
}

static void gpu_blocks(bool with_long)
{
  { // block 1: k-mer hash values
    std::string seq = "ACATGCATGCA";
    const uint64_t want[2][3] = { { 0x38cc00f940aebdaeULL, 0xab7e1b110e086fc6ULL, 0x11a1818bcfdd553ULL },
                                  { 0x603a48c5a11c794aULL, 0xe66016e61816b9c4ULL, 0xc5b13cb146996ffeULL } };
    nthash::NtHash h(seq, 3, 5);
    h.roll();
    nthash::BlindNtHash blind(seq.data(), 3, 5);
    CHECK(same(h.hashes(), blind.hashes(), 3));
    for (auto& w : want) {
      CHECK(h.roll());
      CHECK(same(w, h.hashes(), 3));
      blind.roll(seq[blind.get_pos() + 5]);
      CHECK(same(w, blind.hashes(), 3) && blind.get_forward_hash() == h.get_forward_hash());
    }
  }
  { // block 2: rolling count, identical first/last 4-mer
    std::string seq = "AGTCAGTC";
    nthash::NtHash h(seq, 3, 4);
    std::vector<std::vector<uint64_t>> all;
    while (h.roll()) all.push_back(snap(h.hashes(), 3));
    CHECK(all.size() == seq.size() - 4 + 1 && all.front() == all.back());
  }
  { // block 3: rolled == freshly initialised
    std::string seq = "ACGTACACTGGACTGAGTCT";
    nthash::NtHash h(seq, 3, seq.size() - 2);
    size_t i = 0;
    for (; h.roll(); ++i) {
      std::string sub = seq.substr(i, 18);
      nthash::NtHash f(sub, 3, 18);
      CHECK(f.roll() && same(h.hashes(), f.hashes(), 3));
    }
    CHECK(i == 3);
  }
  { // block 4: canonical
    std::string f = "ACGTACACTGGACTGAGTCT", r = "AGACTCAGTCCAGTGTACGT";
    nthash::NtHash hf(f, 3, 20), hr(r, 3, 20);
    CHECK(hf.roll() && hr.roll() && same(hf.hashes(), hr.hashes(), 3));
  }
  { // block 5: back rolling replays the forward pass
    std::string seq = "ACTAGCTG";
    nthash::NtHash h(seq, 3, 5);
    std::vector<std::vector<uint64_t>> st;
    while (h.roll()) st.push_back(snap(h.hashes(), 3));
    CHECK(st.size() == 4);
    do {
      CHECK(!st.empty() && snap(h.hashes(), 3) == st.back());
      st.pop_back();
    } while (h.roll_back());
    CHECK(st.empty());
  }
  { // block 6: peeking
    std::string seq = "ACTGATCAG";
    nthash::NtHash h(seq, 3, 6);
    h.roll();
    for (int s = 0; s < 3; ++s) {
      h.peek();
      auto pk = snap(h.hashes(), 3);
      h.peek(seq[h.get_pos() + 6]);
      CHECK(snap(h.hashes(), 3) == pk);
      h.roll();
      CHECK(snap(h.hashes(), 3) == pk);
    }
    h.peek_back();
    auto pb = snap(h.hashes(), 3);
    h.roll_back();
    CHECK(snap(h.hashes(), 3) == pb);
  }
  { // block 7: skipping Ns (emitted positions)
    std::string seq = "ACGTACACTGGACTGAGTCT";
    seq[10] = seq[11] = 'N';
    nthash::NtHash h(seq, 3, 8);
    std::vector<size_t> pos;
    while (h.roll()) pos.push_back(h.get_pos());
    CHECK(pos == std::vector<size_t>({ 0, 1, 2, 12 }));
  }
  { // block 8: RNA
    std::string d = "ACGTACACTGGACTGAGTCTACGG", r = "ACGUACACUGGACUGAGUCUACGG";
    nthash::NtHash hd(d, 3, 20), hr(r, 3, 20);
    while (hd.roll()) CHECK(hr.roll() && same(hd.hashes(), hr.hashes(), 3));
    CHECK(!hr.roll());
  }
  { // block 9: spaced seed hash values
    std::string seq = "ACATGCATGCA";
    const uint64_t want[3][3] = { { 0x10be4904ad8de5dULL, 0x3e29e4f4c991628cULL, 0x3f35c984b13feb20ULL },
                                  { 0x8200a7aa3eaf17c8ULL, 0x344198402f4c2a9cULL, 0xb6423fe62e69c40cULL },
                                  { 0x3ce8adcbeaa56532ULL, 0x162e91a4dbedbf11ULL, 0x53173f786a031f45ULL } };
    nthash::SeedNtHash h(seq, std::vector<std::string>{ "11100111" }, 3, 8);
    for (auto& w : want) CHECK(h.roll() && same(w, h.hashes(), 3));
  }
  { // block 10: don't-care mutations, roll == base, peek_back == previous, counts
    std::string seq = "ACGTACACTGGACTGAGTCT";
    std::vector<std::string> seeds = { "111110000000011111", "111111100001111111" };
    std::vector<std::string> muts = { "ACGTACACTTGACTGAGTCT", "ACGTACACTGTACTGAGTCT", "ACGTACACTGCACTGAGTCT" };
    nthash::SeedNtHash h(seq, seeds, 2, 18);
    CHECK(h.get_hash_num() == 4 && h.get_hash_num_per_seed() == 2);
    std::vector<nthash::SeedNtHash> hm;
    for (auto& m : muts) hm.emplace_back(m, seeds, 2, 18);
    std::vector<std::vector<uint64_t>> hist;
    size_t steps = 0;
    for (; h.roll(); ++steps) {
      for (size_t j = 0; j < muts.size(); ++j) {
        CHECK(hm[j].roll() && same(h.hashes(), hm[j].hashes(), 4));
        std::string sub = muts[j].substr(steps, 18);
        nthash::SeedNtHash base(sub, seeds, 2, 18);
        CHECK(base.roll() && same(h.hashes(), base.hashes(), 4));
      }
      hist.push_back(snap(h.hashes(), 4));
      if (h.get_pos() > 0) {
        h.peek_back();
        CHECK(snap(h.hashes(), 4) == hist[hist.size() - 2]);
        h.peek_back(seq[h.get_pos() - 1]);
        CHECK(snap(h.hashes(), 4) == hist[hist.size() - 2]);
      }
    }
    CHECK(steps == 3);
    for (auto& x : hm) CHECK(!x.roll());
  }
  { // block 11: spaced seed back roll
    std::string seq = "ACTAGCTG";
    nthash::SeedNtHash h(seq, std::vector<std::string>{ "110011" }, 3, 6);
    std::vector<std::vector<uint64_t>> st;
    while (h.roll()) st.push_back(snap(h.hashes(), 3));
    CHECK(st.size() == 3);
    do {
      CHECK(!st.empty() && snap(h.hashes(), 3) == st.back());
      st.pop_back();
    } while (h.roll_back());
  }
  { // block 12: strand symmetry with eight palindromic seeds, k = 50
    std::string f = "CACTCGGCCACACACACACACACACACCCTCACACACACAAAACGCACAC", r = "GTGTGCGTTTTGTGTGTGTGAGGGTGTGTGTGTGTGTGTGTGGCCGAGTG";
    std::vector<std::string> seeds = { "11011000001100101101011000011010110100110000011011", "01010000101001110100111011011100101110010100001010",
                                       "11100000100111010111000100100011101011100100000111", "01111000011000111101000011000010111100011000011110",
                                       "00111000011000111101000011000010111100011000011100", "00000000000000000000000011000000000000000000000000",
                                       "11111111111111111111111100111111111111111111111111", "11111111111111111111111111111111111111111111111111" };
    nthash::SeedNtHash h1(f, seeds, 4, 50), h2(r, seeds, 4, 50);
    CHECK(h1.roll() && h2.roll() && same(h1.hashes(), h2.hashes(), 32));
  }
  { // block 13: copying SeedNtHash objects
    std::string seq = "AACGTGACTACTGACTAGCTAGCTAGCTGATCGT";
    std::vector<std::string> seeds = { "111111111101111111111", "110111010010010111011" };
    nthash::SeedNtHash h1(seq, seeds, 4, 21);
    h1.roll();
    nthash::SeedNtHash h2(h1);
    while (h1.roll()) CHECK(h2.roll() && same(h1.hashes(), h2.hashes(), 8));
    CHECK(!h2.roll());
  }
  { // block 14: BlindSeedNtHash == SeedNtHash
    std::string seq = "ATGCTAGTAGCTGAC";
    std::vector<std::string> seeds = { "110011", "101101" };
    nthash::SeedNtHash h1(seq, seeds, 3, 6);
    h1.roll();
    nthash::BlindSeedNtHash h2(seq.data(), seeds, 3, 6);
    CHECK(same(h1.hashes(), h2.hashes(), 6));
    while (h1.roll()) {
      h2.roll(seq[h2.get_pos() + 6]);
      CHECK(same(h1.hashes(), h2.hashes(), 6) && same(h1.get_forward_hash(), h2.get_forward_hash(), 2));
    }
  }
  { // block 17: k-mer vs full-care spaced seed; parsed-seed constructor
    std::string seq = "ATGCTAGTAGCTGAC";
    nthash::NtHash km(seq, 3, 5);
    nthash::SeedNtHash sd(seq, std::vector<std::string>{ "11111" }, 3, 5);
    nthash::SeedNtHash sp(seq, nthash::parse_seeds({ "11111" }), 3, 5);
    while (km.roll()) CHECK(sd.roll() && sp.roll() && same(km.hashes(), sd.hashes(), 3) && same(km.hashes(), sp.hashes(), 3));
    CHECK(!sd.roll());
  }
  if (with_long) { // a sequence longer than one GPU chunk, with a start position and an N far inside
    std::string seq(5000000, 'A');
    uint64_t x = 88172645463325252ULL;
    for (auto& c : seq) {
      x ^= x << 13; x ^= x >> 7; x ^= x << 17;
      c = "ACGT"[x & 3];
    }
    seq[4500000] = 'N';
    nthash::NtHash h(seq, 1, 31, 100);
    size_t n = 0, last = 0;
    uint64_t sum = 0;
    while (h.roll()) {
      ++n;
      last = h.get_pos();
      sum += h.hashes()[0];
    }
    CHECK(n == seq.size() - 31 + 1 - 100 - 31 && last == seq.size() - 31);
    const std::string piece = seq.substr(4194300, 40);   // the classes borrow the sequence, as the reference does
    nthash::NtHash spot(piece, 1, 31);                   // straddles the chunk boundary at window 2^22
    nthash::NtHash big(seq, 1, 31, 4194300);
    for (int i = 0; i < 10; ++i) CHECK(spot.roll() && big.roll() && spot.hashes()[0] == big.hashes()[0]);
    std::printf("long sequence: %zu windows, checksum %016llx\n", n, (unsigned long long)sum);
  }
}

int main(int argc, char** argv)
{
  const bool host_only = argc > 1 && std::string(argv[1]) == "--host-only";
  const bool short_only = argc > 1 && std::string(argv[1]) == "--short-only";
  host_only_blocks();
  if (!host_only) gpu_blocks(!short_only);
  std::printf("%s: %d failure(s)\n", host_only ? "host-only blocks" : "all blocks", failures);
  return failures ? 1 : 0;
}
