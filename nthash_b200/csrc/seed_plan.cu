// seed_plan.cu — see seed_plan.hpp.  Host code only.
#include "seed_plan.hpp"
#include "nthash_dev.cuh"

#include <cstring>

namespace nthb {

namespace {

constexpr uint32_t TABLE_BUDGET = 64 * 1024; // bytes of lookup tables a CTA is willing to hold
constexpr uint64_t SEED_OF_CODE[4] = { SEED_A, SEED_C, SEED_T, SEED_G }; // code = (byte >> 1) & 3

// The reference's care/ignore decomposition of one seed (seed.cpp:19-66): returns its block list
// (pairs [start,end)) in the order SeedNtHash::init walks it.
std::vector<uint32_t> reference_blocks(const std::string& seed)
{
  const uint32_t k = (uint32_t)seed.size();
  std::vector<uint32_t> care_b, ign_b;
  uint32_t n_care_m = 0, n_ign_m = 0;
  const char pad = seed[k - 1] == '1' ? '0' : '1';
  uint32_t start = 0;
  bool care = seed[0] == '1';
  for (uint32_t p = 0; p <= k; ++p) {
    const char ch = p < k ? seed[p] : pad;
    if (care && ch == '0') {
      if (p - start == 1) ++n_care_m;
      else { care_b.push_back(start); care_b.push_back(p); }
      start = p;
      care = false;
    } else if (!care && ch == '1') {
      if (p - start == 1) ++n_ign_m;
      else { ign_b.push_back(start); ign_b.push_back(p); }
      start = p;
      care = true;
    }
  }
  const uint32_t num_cares = (uint32_t)care_b.size() + n_care_m;      // 2 per block + 1 per monomer
  const uint32_t num_ignores = (uint32_t)ign_b.size() + n_ign_m + 2;
  if (num_ignores < num_cares) {
    ign_b.push_back(0);
    ign_b.push_back(k);
    return ign_b;
  }
  return care_b;
}

} // namespace

std::string build_seed_plan(const char* const* seeds, uint32_t n_seeds, uint32_t k, uint32_t h, SeedPlanHost& plan)
{
  if (n_seeds == 0 || !seeds) return "at least one spaced seed is required";
  std::vector<std::string> sv;
  for (uint32_t i = 0; i < n_seeds; ++i) {
    if (!seeds[i]) return "NULL seed string";
    sv.emplace_back(seeds[i]);
    if (sv.back().size() != k) // seed.cpp:90-95
      return "Spaced seed string length (" + std::to_string(sv.back().size()) + ") not equal to k=" + std::to_string(k) +
             " in " + sv.back();
    if (sv.back().find_first_not_of("01") != std::string::npos) return "spaced seed " + sv.back() + " has characters other than 0/1";
  }
  plan = SeedPlanHost();
  plan.k = k;
  plan.h = h;
  plan.n_seeds = n_seeds;

  // positions each seed will look up, and in which mode
  std::vector<std::vector<uint32_t>> lookups(n_seeds);
  std::vector<uint32_t> ignore_mode(n_seeds, 0);
  uint32_t total_pos = 0;
  for (uint32_t s = 0; s < n_seeds; ++s) {
    const std::string& sd = sv[s];
    if (!std::equal(sd.begin(), sd.end(), sd.rbegin())) plan.all_symmetric = false;
    std::vector<uint32_t> care, ign;
    for (uint32_t q = 0; q < k; ++q) (sd[q] == '1' ? care : ign).push_back(q); // parse_seeds: anything but '1' is a don't-care (seed.cpp:431-447)
    // ignore-mode costs one shared full-window roll (about four lookups' worth) on top of its positions
    if (ign.size() + 4 < care.size()) {
      ignore_mode[s] = 1;
      plan.any_ignore = true;
      lookups[s] = ign;
    } else {
      lookups[s] = care;
    }
    total_pos += (uint32_t)lookups[s].size();
  }
  // positions per group: the largest of 4,3,2,1 whose tables fit the budget
  uint32_t gsz = 4;
  for (; gsz > 1; --gsz) {
    uint64_t bytes = 0;
    for (uint32_t s = 0; s < n_seeds; ++s) {
      const uint32_t np = (uint32_t)lookups[s].size();
      bytes += (uint64_t)(np / gsz) * (16ull << (2 * gsz)) + (np % gsz ? (16ull << (2 * (np % gsz))) : 0);
    }
    if (bytes <= TABLE_BUDGET) break;
  }
  if (gsz == 1 && (uint64_t)total_pos * 64 > TABLE_BUDGET)
    return "seed set needs " + std::to_string(total_pos) + " per-window lookups: more than the engine's table budget";

  std::vector<SeedDesc> descs(n_seeds);
  std::vector<SeedGroup> groups;
  std::vector<uint4> tables;
  std::vector<uint32_t> refblk;
  for (uint32_t s = 0; s < n_seeds; ++s) {
    SeedDesc& d = descs[s];
    memset(&d, 0, sizeof d);
    d.g0 = (uint32_t)groups.size();
    d.ignore_mode = ignore_mode[s];
    const std::vector<uint32_t>& lp = lookups[s];
    for (size_t a = 0; a < lp.size(); a += gsz) {
      SeedGroup g;
      memset(&g, 0, sizeof g);
      g.npos = (uint32_t)std::min<size_t>(gsz, lp.size() - a);
      for (uint32_t j = 0; j < g.npos; ++j) g.pos[j] = (uint16_t)lp[a + j];
      g.tab_off = (uint32_t)(tables.size() * sizeof(uint4));
      const uint32_t n_entries = 1u << (2 * g.npos);
      for (uint32_t idx = 0; idx < n_entries; ++idx) {
        uint64_t f = 0, r = 0;
        for (uint32_t j = 0; j < g.npos; ++j) {
          const uint32_t code = (idx >> (2 * (g.npos - 1 - j))) & 3, q = g.pos[j];
          f ^= srol_n(SEED_OF_CODE[code], k - 1 - q); // seed.cpp:155
          r ^= srol_n(SEED_OF_CODE[code ^ 2], q);     // seed.cpp:156 (complement strand)
        }
        tables.push_back(make_uint4((uint32_t)f, (uint32_t)(f >> 32), (uint32_t)r, (uint32_t)(r >> 32)));
      }
      groups.push_back(g);
    }
    d.g1 = (uint32_t)groups.size();
    d.rb0 = (uint32_t)refblk.size() / 2;
    const std::vector<uint32_t> rb = reference_blocks(sv[s]);
    refblk.insert(refblk.end(), rb.begin(), rb.end());
    d.rb1 = (uint32_t)refblk.size() / 2;
  }
  plan.seed_strings = sv;
  plan.lookups = lookups;
  plan.ignore_mode = ignore_mode;
  plan.n_groups = (uint32_t)groups.size();
  plan.care_words = (k + 31) / 32;
  std::vector<uint32_t> care_bits((size_t)plan.care_words * n_seeds, 0);
  for (uint32_t s = 0; s < n_seeds; ++s)
    for (uint32_t q = 0; q < k; ++q)
      if (sv[s][q] == '1') care_bits[(size_t)s * plan.care_words + q / 32] |= 1u << (q % 32);

  auto align16 = [](size_t x) { return (x + 15) & ~size_t(15); };
  plan.groups_off = (uint32_t)align16(descs.size() * sizeof(SeedDesc));
  plan.tables_off = (uint32_t)align16(plan.groups_off + groups.size() * sizeof(SeedGroup));
  plan.smem_bytes = (uint32_t)align16(plan.tables_off + tables.size() * sizeof(uint4));
  plan.care_off = plan.smem_bytes;
  plan.refblk_off = (uint32_t)align16(plan.care_off + care_bits.size() * 4);
  plan.blob.assign(align16(plan.refblk_off + refblk.size() * 4 + 16), 0);
  memcpy(plan.blob.data(), descs.data(), descs.size() * sizeof(SeedDesc));
  memcpy(plan.blob.data() + plan.groups_off, groups.data(), groups.size() * sizeof(SeedGroup));
  memcpy(plan.blob.data() + plan.tables_off, tables.data(), tables.size() * sizeof(uint4));
  memcpy(plan.blob.data() + plan.care_off, care_bits.data(), care_bits.size() * 4);
  memcpy(plan.blob.data() + plan.refblk_off, refblk.data(), refblk.size() * 4);
  return std::string();
}

} // namespace nthb
