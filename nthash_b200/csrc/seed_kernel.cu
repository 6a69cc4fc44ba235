// seed_kernel.cu — batched SeedNtHash (spaced-seed) kernels for sm_100a.
//
// Replaces, for whole batches, `nthash::SeedNtHash it(seq, len, seeds, h, k); while (it.roll()) ...`:
//   SeedNtHash::init / ::roll   src/seed.cpp:493-544
//   ntmsm64 (base and roll)     src/seed.cpp:130-175, :177-207, :230-249
//   multi-hash extension        src/seed.cpp:167-172 (same arithmetic as extend_hashes)
//
// Structure.  Same work decomposition and TMA-staged base tiles as the k-mer kernel:
// one thread per item (run of consecutive windows of one read), 32 items per warp.  A window's hash
// for seed s depends only on the seed's care mask, not on how the reference walks it, so it is
// evaluated here as
//     hash_s(p) = [FULL(p) if ignore-mode] ^ XOR_g TAB_g[codes of the bases at group g's positions]
// with the groups/tables prepared on the host (seed_plan.cu) and held in shared memory.
//
// The reference's two byte-level quirks are kept exact by two rare slow paths:
//  * non-ACGTU bytes are HASHED by SeedNtHash (forward seed 0, reverse seed SEED_TAB[c & 7],
//    src/internal.hpp:132-165): windows containing such a byte are recomputed byte-exactly;
//  * which windows are visited at all follows SeedNtHash::roll's jump rule (seed.cpp:524-530) and
//    init's NUL rule (seed.cpp:151, :497-511): reads containing such bytes are replayed by
//    seed_emit_kernel, one thread per read, which clears the rows/bits the reference never visits.
#include "engine.hpp"
#include "kmer_common.cuh"
#include "seed_plan.hpp"

namespace nthb {

namespace {

constexpr int S_LUT_BYTES = 256, S_PAIR_BYTES = 256, S_IN_BYTES = 64;

NTH_D uint32_t code_of(uint32_t c) { return (c >> 1) & 3u; }

// Byte-exact contribution of one window to seed `sd` (slow path): seed.cpp:149-166.
NTH_D void exact_window(const uint8_t* w, const uint32_t* care, uint32_t k, uint64_t& f, uint64_t& r)
{
  f = 0;
  r = 0;
  for (uint32_t q = 0; q < k; ++q) {
    if (care[q >> 5] >> (q & 31) & 1u) {
      const unsigned c = w[q];
      f ^= srol_n(seed_of_byte(c), k - 1 - q);
      r ^= srol_n(seed_of_byte(c & 7u), q);
    }
  }
}

template<bool STRANDS>
__global__ void __launch_bounds__(KMER_NT) seed_kernel(const __grid_constant__ SeedParams P)
{
  extern __shared__ __align__(16) uint8_t smem[];
  // [plan: descs | groups | tables][validity LUT][pair table][in-only table][mbarrier][tile]
  const uint32_t plan_bytes = P.plan_smem_bytes;
  uint8_t* lut = smem + plan_bytes;
  uint4* pair = reinterpret_cast<uint4*>(lut + S_LUT_BYTES);
  uint4* in_tab = reinterpret_cast<uint4*>(lut + S_LUT_BYTES + S_PAIR_BYTES);
  uint64_t* bar = reinterpret_cast<uint64_t*>(lut + S_LUT_BYTES + S_PAIR_BYTES + S_IN_BYTES);
  uint8_t* tile = reinterpret_cast<uint8_t*>(bar) + 16;
  constexpr int TILE_PAD = 16;
  __shared__ uint64_t s_range[2];

  const uint32_t tid = threadIdx.x;
  const uint64_t i0 = (uint64_t)blockIdx.x * KMER_NT;
  const uint64_t i1 = min(i0 + (uint64_t)KMER_NT, P.g.n_items);
  const uint32_t k = P.k, h = P.h, m = P.n_seeds, H = m * h;

  uint64_t my_byte = 0, my_out = 0;
  uint32_t n = 0;
  if (i0 + tid < i1) {
    const KmerGeom& g = P.g;
    const uint64_t i = i0 + tid;
    if (g.item_byte) {
      my_byte = g.item_byte[i];
      my_out = g.item_out[i];
      n = (uint32_t)(g.item_out[i + 1] - my_out);
    } else {
      const uint64_t r = g.segs > 1 ? i / g.segs : i;
      const uint32_t sg = (uint32_t)(i - r * g.segs);
      my_byte = r * g.read_len + (uint64_t)sg * g.seg;
      my_out = r * g.nk + (uint64_t)sg * g.seg;
      n = min(g.seg, g.nk - sg * g.seg);
    }
    if (tid == 0) s_range[0] = my_byte;
    if (i == i1 - 1) s_range[1] = my_byte + (n ? n + k - 1 : 0);
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (tid < TILE_PAD) tile[tid] = 'A';
  __syncthreads();
  const uint64_t g0 = (s_range[0] ? s_range[0] - 1 : 0) & ~15ull, g1 = s_range[1];
  if (g1 - g0 > P.tile_cap) __trap();
  const uint64_t bulk_end = min((g1 + 15) & ~15ull, P.n_bases & ~15ull);
  const uint32_t bulk_bytes = bulk_end > g0 ? (uint32_t)(bulk_end - g0) : 0u;
  if (tid == 0) { // base tile and the plan (descriptors + tables) arrive through the TMA engine
    mbar_expect_tx(bar, bulk_bytes + plan_bytes);
    if (bulk_bytes) bulk_g2s(tile + TILE_PAD, P.bases + g0, bulk_bytes, bar);
    bulk_g2s(smem, P.plan_blob, plan_bytes, bar);
  }
  // 1: the byte's windows need the byte-exact path (and the read the emission replay, which applies SeedNtHash::roll's own
  // test SEED_TAB[c] == SEED_N, seed.cpp:527).  That is every byte without a seed, and the raw bytes 1, 3, 4, 5, 7: they do
  // have seeds (SEED_TAB's complement slots) but their 2-bit code (c >> 1) & 3 is not the base that seed belongs to.
  lut[tid] = (seed_of_byte(tid) != 0 && tid > 7) ? 0 : 1;
  if (tid < 16) {
    auto seed_code = [&](int c) { return c ^ (c >> 1); }; // code -> index into P.s/P.sk (A,C,G,T order)
    const int ci = tid >> 2, co = tid & 3;
    const uint64_t f = P.s[seed_code(ci)] ^ P.sk[seed_code(co)];
    const uint64_t r = P.sk[seed_code(ci ^ 2)] ^ P.s[seed_code(co ^ 2)];
    pair[tid] = make_uint4((uint32_t)f, (uint32_t)(f >> 32), (uint32_t)r, (uint32_t)(r >> 32));
    if (tid < 4) {
      const uint64_t fi = P.s[seed_code(tid)], ri = P.sk[seed_code(tid ^ 2)];
      in_tab[tid] = make_uint4((uint32_t)fi, (uint32_t)(fi >> 32), (uint32_t)ri, (uint32_t)(ri >> 32));
    }
  }
  for (uint64_t g = max(bulk_end, g0) + tid; g < g1; g += KMER_NT) tile[TILE_PAD + (g - g0)] = P.bases[g];
  mbar_wait(bar, 0);
  __syncthreads();
  if (n == 0) return;

  const SeedDesc* descs = reinterpret_cast<const SeedDesc*>(smem);
  const SeedGroup* groups = reinterpret_cast<const SeedGroup*>(smem + P.groups_off);
  const uint8_t* tables = smem + P.tables_off;
  const uint8_t* ps = tile + TILE_PAD + (my_byte - g0);

  // does the item hold a byte SEED_TAB maps to zero?  (decides both slow paths)
  uint32_t bad = 0;
  for (uint32_t j = 0; j < n + k - 1; ++j) bad |= lut[ps[j]];

  // full-window hash for ignore-mode seeds: the NtHash recurrence on 2-bit codes (kmer_fast_kernel.cu)
  State full = { 0u, 0u, 0u, 0u };
  if (P.any_ignore) {
    for (uint32_t j = 0; j < k; ++j) {
      const uint4 e = in_tab[code_of(ps[(int)j - 1])];
      fwd_step(full, e.x, e.y, 0u, 0u);
      rev_step(full, e.z, e.w, 0u, 0u);
    }
  }

  for (uint32_t p = 0; p < n; ++p) {
    const uint8_t* w = ps + p;
    if (P.any_ignore) {
      const uint4 e = pair[code_of(w[k - 1]) * 4 + code_of(w[-1])];
      fwd_step(full, e.x, e.y, 0u, 0u);
      rev_step(full, e.z, e.w, 0u, 0u);
    }
    const uint64_t row = my_out + p;
    uint64_t* o = P.out + row * H;
    for (uint32_t s = 0; s < m; ++s) {
      const SeedDesc d = descs[s];
      uint32_t flo = 0, fhi = 0, rlo = 0, rhi = 0;
      if (d.ignore_mode) {
        flo = full.flo; fhi = full.fhi; rlo = full.rlo; rhi = full.rhi;
      }
      for (uint32_t g = d.g0; g < d.g1; ++g) {
        const SeedGroup sg = groups[g];
        uint32_t idx = 0;
        for (uint32_t j = 0; j < sg.npos; ++j) idx = idx * 4 + code_of(w[sg.pos[j]]);
        const uint4 e = *reinterpret_cast<const uint4*>(tables + sg.tab_off + idx * 16);
        flo ^= e.x; fhi ^= e.y; rlo ^= e.z; rhi ^= e.w;
      }
      const uint64_t f = ((uint64_t)fhi << 32) | flo, r = ((uint64_t)rhi << 32) | rlo;
      const uint64_t h0 = f + r;
      o[s * h] = h0;
      for (uint32_t q = 1; q < h; ++q) o[s * h + q] = ext_hash(h0, ext_mult(q, k));
      if (STRANDS) {
        P.out_fwd[row * m + s] = f;
        P.out_rev[row * m + s] = r;
      }
    }
  }

  if (bad) {
    // (1) windows holding a zero-seed byte: byte-exact values (the 2-bit codes above were garbage there)
    const uint32_t* care = reinterpret_cast<const uint32_t*>(P.plan_blob + P.care_off);
    uint32_t run = 0;
    for (uint32_t j = 0; j < n + k - 1; ++j) {
      run = lut[ps[j]] ? 0 : run + 1;
      if (j >= k - 1 && run < k) {
        const uint32_t p = j - (k - 1);
        const uint64_t row = my_out + p;
        for (uint32_t s = 0; s < m; ++s) {
          uint64_t f, r;
          exact_window(ps + p, care + (size_t)s * P.care_words, k, f, r);
          const uint64_t h0 = f + r;
          P.out[row * H + s * h] = h0;
          for (uint32_t q = 1; q < h; ++q) P.out[row * H + s * h + q] = ext_hash(h0, ext_mult(q, k));
          if (STRANDS) {
            P.out_fwd[row * m + s] = f;
            P.out_rev[row * m + s] = r;
          }
        }
      }
    }
    // (2) the read needs the sequential emission replay
    const uint64_t i = i0 + tid;
    const uint64_t rd = P.item_read ? P.item_read[i] : (P.g.item_byte ? i : (P.g.segs > 1 ? i / P.g.segs : i));
    P.read_dirty[rd] = 1;
  }
}

// One thread per read that holds a zero-seed byte: replay SeedNtHash's visiting order
// (init: seed.cpp:493-516, roll: :518-544) and clear every row it does not visit.
template<bool STRANDS>
__global__ void __launch_bounds__(128) seed_emit_kernel(const __grid_constant__ SeedParams P, uint64_t n_reads)
{
  const uint64_t rd = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (rd >= n_reads || !P.read_dirty[rd]) return;
  uint64_t byte0, row0, len;
  if (P.read_off) {
    byte0 = P.read_off[rd];
    len = P.read_off[rd + 1] - byte0;
    row0 = P.koff[rd];
  } else {
    byte0 = rd * P.g.read_len;
    len = P.g.read_len;
    row0 = rd * P.g.nk;
  }
  const uint32_t k = P.k, m = P.n_seeds, H = m * P.h;
  if (len < k) return;
  const uint8_t* s = P.bases + byte0;
  const SeedDesc* descs = reinterpret_cast<const SeedDesc*>(P.plan_blob);
  const uint32_t* refblk = reinterpret_cast<const uint32_t*>(P.plan_blob + P.refblk_off);
  const uint64_t nk = len - k + 1;

  auto clear_rows = [&](uint64_t a, uint64_t b) { // windows [a, b) are never visited
    for (uint64_t p = a; p < b; ++p) {
      const uint64_t row = row0 + p;
      for (uint32_t q = 0; q < H; ++q) P.out[row * H + q] = 0;
      if (STRANDS)
        for (uint32_t q = 0; q < m; ++q) {
          P.out_fwd[row * m + q] = 0;
          P.out_rev[row * m + q] = 0;
        }
      if (P.valid_bits) atomicAnd(&P.valid_bits[(P.valid_row0 + row) >> 5], ~(1u << ((P.valid_row0 + row) & 31)));
    }
  };
  // ntmsm64 (base) fails only on a NUL byte at a block position, scanning seeds, blocks, positions in order
  auto first_nul = [&](uint64_t pos, uint32_t& loc) {
    for (uint32_t sd = 0; sd < m; ++sd)
      for (uint32_t b = descs[sd].rb0; b < descs[sd].rb1; ++b)
        for (uint32_t q = refblk[2 * b]; q < refblk[2 * b + 1]; ++q)
          if (s[pos + q] == 0) {
            loc = q;
            return true;
          }
    return false;
  };

  uint64_t pos = 0, next_unvisited = 0;
  bool done = false;
  while (!done) {
    // init()
    uint32_t loc = 0;
    while (pos < nk && first_nul(pos, loc)) pos += loc + 1;
    if (pos > len - k) break;
    clear_rows(next_unvisited, pos);
    next_unvisited = pos + 1; // pos is visited
    // roll() until the next jump
    for (;;) {
      if (pos >= len - k) {
        done = true;
        break;
      }
      if (seed_of_byte(s[pos + k]) == 0) {
        pos += k;
        break; // -> init()
      }
      ++pos;
      next_unvisited = pos + 1;
    }
  }
  clear_rows(next_unvisited, nk);
}

// Fused consumer, second launch (uniform batches): the items the specialised kernel left out because they hold a byte
// that needs the exact path.  One thread per flagged read replays SeedNtHash's visiting order (init: seed.cpp:493-516,
// roll: :518-544) and adds the byte-exact hashes (seed.cpp:149-171) of every visited window that lies in a flagged item;
// the read's other items were complete and are already in the result.
__global__ void __launch_bounds__(128) seed_reduce_dirty_kernel(const __grid_constant__ SeedParams P, uint64_t n_reads)
{
  const uint64_t rd = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (rd >= n_reads || !P.read_dirty[rd]) return;
  const uint32_t k = P.k, m = P.n_seeds, h = P.h;
  const uint64_t len = P.g.read_len, nk = P.g.nk, seg = P.g.seg, segs = P.g.segs ? P.g.segs : 1;
  if (len < k) return;
  const uint8_t* s = P.bases + rd * len;
  const SeedDesc* descs = reinterpret_cast<const SeedDesc*>(P.plan_blob);
  const uint32_t* refblk = reinterpret_cast<const uint32_t*>(P.plan_blob + P.refblk_off);
  const uint32_t* care = reinterpret_cast<const uint32_t*>(P.plan_blob + P.care_off);
  uint64_t cnt = 0, sum = 0, x = 0;
  auto visit = [&](uint64_t p) {
    if (!P.item_dirty[rd * segs + p / seg]) return;
    ++cnt;
    for (uint32_t sd = 0; sd < m; ++sd) {
      uint64_t f, r;
      exact_window(s + p, care + (size_t)sd * P.care_words, k, f, r);
      const uint64_t h0 = f + r;
      sum += h0;
      x ^= h0;
      for (uint32_t q = 1; q < h; ++q) {
        const uint64_t e = ext_hash(h0, ext_mult(q, k));
        sum += e;
        x ^= e;
      }
    }
  };
  auto first_nul = [&](uint64_t pos, uint32_t& loc) { // ntmsm64 (base) fails only on a NUL byte at a block position
    for (uint32_t sd = 0; sd < m; ++sd)
      for (uint32_t b = descs[sd].rb0; b < descs[sd].rb1; ++b)
        for (uint32_t q = refblk[2 * b]; q < refblk[2 * b + 1]; ++q)
          if (s[pos + q] == 0) {
            loc = q;
            return true;
          }
    return false;
  };
  uint64_t pos = 0;
  bool done = false;
  while (!done) {
    uint32_t loc = 0;
    while (pos < nk && first_nul(pos, loc)) pos += loc + 1; // init()
    if (pos > len - k) break;
    visit(pos);
    for (;;) { // roll() until the next jump
      if (pos >= len - k) {
        done = true;
        break;
      }
      if (seed_of_byte(s[pos + k]) == 0) {
        pos += k;
        break; // -> init()
      }
      ++pos;
      visit(pos);
    }
  }
  if (cnt) {
    atomicAdd(reinterpret_cast<unsigned long long*>(P.reduce_out), (unsigned long long)cnt);
    atomicAdd(reinterpret_cast<unsigned long long*>(P.reduce_out) + 1, (unsigned long long)sum);
    atomicXor(reinterpret_cast<unsigned long long*>(P.reduce_out) + 2, (unsigned long long)x);
  }
}

} // namespace

cudaError_t launch_seed_reduce_dirty(const SeedParams& P, uint64_t n_reads, cudaStream_t st)
{
  seed_reduce_dirty_kernel<<<(unsigned)((n_reads + 127) / 128), 128, 0, st>>>(P, n_reads);
  return cudaGetLastError();
}

uint32_t seed_smem_bytes(uint32_t plan_smem, uint32_t tile_cap)
{
  return plan_smem + S_LUT_BYTES + S_PAIR_BYTES + S_IN_BYTES + 16 + 16 + tile_cap + 16;
}

cudaError_t launch_seed(SeedParams P, uint64_t n_reads, cudaStream_t st)
{
  const uint64_t base[4] = { SEED_A, SEED_C, SEED_G, SEED_T };
  for (int x = 0; x < 4; ++x) {
    P.s[x] = base[x];
    P.sk[x] = srol_n(base[x], P.k);
  }
  const uint32_t smem = seed_smem_bytes(P.plan_smem_bytes, P.tile_cap);
  const uint64_t ctas = (P.g.n_items + KMER_NT - 1) / KMER_NT;
  if (ctas == 0) return cudaSuccess;
  if (ctas > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  const bool strands = P.out_fwd != nullptr;
  auto fn = strands ? seed_kernel<true> : seed_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  fn<<<(unsigned)ctas, KMER_NT, smem, st>>>(P);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return launch_seed_emit(P, n_reads, st);
}

cudaError_t launch_seed_emit(const SeedParams& P, uint64_t n_reads, cudaStream_t st)
{
  const unsigned eb = (unsigned)((n_reads + 127) / 128);
  if (P.out_fwd) seed_emit_kernel<true><<<eb, 128, 0, st>>>(P, n_reads);
  else seed_emit_kernel<false><<<eb, 128, 0, st>>>(P, n_reads);
  return cudaGetLastError();
}

} // namespace nthb
