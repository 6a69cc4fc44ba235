// seed_plan.hpp — host-side compilation of spaced-seed patterns into what the seed kernels consume.
//
// Replaces the seed handling of the SeedNtHash constructors (src/seed.cpp:449-491): check_seeds
// (:85-104), get_blocks (:19-66) and parsed_seeds_to_blocks (:68-83).  The reference turns a seed
// into care/ignore blocks and monomers for its O(#blocks) rolling update; hash VALUES depend only
// on the care mask, so the engine is free to evaluate the mask its own way:
//   hash_s(p) = [FULL(p) if ignore-mode] ^ XOR over groups g of TAB_g[codes of the bases at g's positions]
// with up to four positions per group and one precombined 16-byte table entry per code tuple.
// The reference's own block lists are kept too: SeedNtHash::init rejects a window only for a NUL byte
// at a *block* position (seed.cpp:151), so the exact emission rule needs them.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace nthb {

struct SeedGroup // one table lookup: up to four window positions, first position most significant in the index
{
  uint16_t pos[4];
  uint32_t tab_off; // byte offset of the group's table inside the tables section
  uint32_t npos;
};

struct SeedDesc
{
  uint32_t g0, g1;      // groups [g0, g1) belong to this seed
  uint32_t ignore_mode; // 1: start from the full-window hash and XOR the ignored positions back out
  uint32_t rb0, rb1;    // reference blocks [rb0, rb1) (pairs start,end) in the ref_blocks section
  uint32_t pad[3];
};

// Layout of the device blob: [SeedDesc x n_seeds][SeedGroup x n_groups][tables][care bitmasks][ref blocks]
struct SeedPlanHost
{
  uint32_t k = 0, h = 0, n_seeds = 0, n_groups = 0;
  uint32_t groups_off = 0, tables_off = 0, care_off = 0, refblk_off = 0; // byte offsets in the blob
  uint32_t smem_bytes = 0; // descs + groups + tables: what every CTA copies into shared memory
  uint32_t care_words = 0; // u32 words per seed in the care bitmask section
  bool any_ignore = false;
  bool all_symmetric = true; // false where the reference prints its "not symmetric" warning (seed.cpp:96-102)
  std::vector<uint8_t> blob;
  // kept for the specialised (run-time compiled) kernel, seed_jit.cu
  std::vector<std::string> seed_strings;
  std::vector<std::vector<uint32_t>> lookups; // per seed: the window positions it looks up (care or ignored ones)
  std::vector<uint32_t> ignore_mode;
};

// Returns an empty string on success, else the reason (what the reference answers with raise_error()).
std::string build_seed_plan(const char* const* seeds, uint32_t n_seeds, uint32_t k, uint32_t h, SeedPlanHost& plan);

} // namespace nthb
