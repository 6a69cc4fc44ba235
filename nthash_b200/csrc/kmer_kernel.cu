// kmer_kernel.cu — batched NtHash (contiguous k-mer) kernel for sm_100a.
//
// Replaces, for whole batches, the per-k-mer loop `while (nthash.roll())` of the reference:
//   NtHash::init / NtHash::roll          src/kmer.cpp:228-264
//   next_forward_hash / next_reverse_hash src/kmer.cpp:84-94, :164-174
//   base_forward_hash / base_reverse_hash src/kmer.cpp:43-73, :123-152 (closed form, evaluated
//                                         here by an in-only warm-up of the same recurrence)
//   extend_hashes                         src/internal.hpp:104-118
//
// Design: the batch is cut into ITEMS = runs of <= seg consecutive windows of one
// read.  One thread rolls one item start to finish (the recurrence is serial per item, items are
// independent), 32 items per warp, NT items per CTA.  Consecutive items cover a contiguous byte
// range of the concatenated reads, so the CTA stages that range into shared memory with ONE 1-D
// TMA bulk copy (cp.async.bulk -> UBLKCP) and every lane then walks its own row with LDS.U8.
// Per-base seeds come from two 256-entry shared-memory tables indexed by the raw byte (no
// per-base decode, exact for any byte value); hashes leave as full 32-byte sectors
// (st.global.v4.u64 -> STG.E.ENL2.256), four consecutive windows of one item per store.
//
// This is the FALLBACK kernel: what kmer_fast_kernel.cu does not take (more than 8 hashes per k-mer, unaligned output
// pointers, NTHASH_B200_DISABLE_TMA_STORE=1 for A/B runs).  Its per-lane stores reach ~0.1-0.6 of the HBM peak.
#include "engine.hpp"
#include "kmer_common.cuh"

namespace nthb {

namespace {

constexpr int TAB_BYTES = 2 * 256 * 16;  // tab_in + tab_out
constexpr int TILE_OFF = TAB_BYTES + 16; // mbarrier lives in the 16 bytes after the tables
constexpr int TILE_PAD = 16;             // front pad: byte "-1" of the very first item lands here

NTH_D void item_geom(const KmerGeom& g, uint64_t i, uint64_t& byte, uint64_t& out, uint32_t& n)
{
  if (g.item_byte) { // ragged: arrays are read_off/koff themselves when every read is one item
    byte = g.item_byte[i];
    out = g.item_out[i];
    n = (uint32_t)(g.item_out[i + 1] - out);
  } else {
    uint64_t r = i;
    uint32_t s = 0;
    if (g.segs > 1) {
      r = i / g.segs;
      s = (uint32_t)(i - r * g.segs);
    }
    byte = r * g.read_len + (uint64_t)s * g.seg;
    out = r * g.nk + (uint64_t)s * g.seg;
    n = min(g.seg, g.nk - s * g.seg);
  }
}

// Slow, exact clean-up for an item that saw a non-ACGTU byte: windows touching one are not
// emitted by the reference (kmer.cpp:232-235, :255-258), so they read back 0 / bit cleared.
template<int H, bool STRANDS>
__device__ __noinline__ void
scrub_item(const KmerParams& P, const uint4* tab_in, const uint8_t* ps, uint64_t o, uint32_t n)
{
  const uint32_t k = P.k, h = H ? H : P.h;
  uint32_t run = 0;
  for (uint32_t j = 0; j < n + k - 1; ++j) {
    run = tab_in[ps[j]].x != 0 ? run + 1 : 0;
    if (j >= k - 1 && run < k) {
      const uint64_t w = o + (j - (k - 1));
      for (uint32_t q = 0; q < h; ++q) P.out[w * h + q] = 0;
      if (STRANDS) {
        P.out_fwd[w] = 0;
        P.out_rev[w] = 0;
      }
      if (P.valid_bits) atomicAnd(&P.valid_bits[(P.valid_row0 + w) >> 5], ~(1u << ((P.valid_row0 + w) & 31)));
    }
  }
}

template<int H, bool STRANDS>
__global__ void __launch_bounds__(KMER_NT) kmer_kernel(const __grid_constant__ KmerParams P)
{
  extern __shared__ __align__(16) uint8_t smem[];
  uint4* tab_in = reinterpret_cast<uint4*>(smem);
  uint4* tab_out = tab_in + 256;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + TAB_BYTES);
  uint8_t* tile = smem + TILE_OFF;
  __shared__ uint64_t s_range[2];

  const uint32_t tid = threadIdx.x;
  const uint64_t i0 = (uint64_t)blockIdx.x * KMER_NT;
  const uint64_t i1 = min(i0 + (uint64_t)KMER_NT, P.g.n_items);
  const uint32_t k = P.k;

  uint64_t my_byte = 0, my_out = 0;
  uint32_t n = 0;
  if (i0 + tid < i1) {
    item_geom(P.g, i0 + tid, my_byte, my_out, n);
    if (tid == 0) s_range[0] = my_byte;
    if (i0 + tid == i1 - 1) s_range[1] = my_byte + (n ? n + k - 1 : 0);
  }
  for (uint32_t i = tid; i < 512; i += KMER_NT) tab_in[i] = make_uint4(0, 0, 0, 0);
  if (tid < TILE_PAD) tile[tid] = 'N';
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  // stage [g0, g1): 16-byte aligned body through the TMA engine, ragged tail by hand
  const uint64_t g0 = (s_range[0] ? s_range[0] - 1 : 0) & ~15ull;
  const uint64_t g1 = s_range[1];
  if (g1 - g0 > P.tile_cap) __trap(); // the host-side plan guarantees the tile fits
  const uint64_t bulk_end = min((g1 + 15) & ~15ull, P.n_bases & ~15ull);
  const uint32_t bulk_bytes = bulk_end > g0 ? (uint32_t)(bulk_end - g0) : 0u;
  if (tid == 0 && bulk_bytes) {
    mbar_expect_tx(bar, bulk_bytes);
    bulk_g2s(tile + TILE_PAD, P.bases + g0, bulk_bytes, bar);
  }
  if (tid < 10) { // the only bytes NtHash hashes: ACGTU in either case
    const int x = tid % 5 < 4 ? tid % 5 : 3;         // A C G T U(=T)
    const unsigned c = (unsigned)"ACGTU"[tid % 5] | (tid >= 5 ? 0x20u : 0u);
    const uint64_t s = P.s[x], sk = P.sk[x], sc = P.s[3 - x], skc = P.sk[3 - x]; // complement = 3 - x
    tab_in[c] = make_uint4((uint32_t)s, (uint32_t)(s >> 32), (uint32_t)skc, (uint32_t)(skc >> 32));
    tab_out[c] = make_uint4((uint32_t)sk, (uint32_t)(sk >> 32), (uint32_t)sc, (uint32_t)(sc >> 32));
  }
  for (uint64_t g = max(bulk_end, g0) + tid; g < g1; g += KMER_NT) tile[TILE_PAD + (g - g0)] = P.bases[g];
  if (bulk_bytes) mbar_wait(bar, 0);
  __syncthreads();
  if (n == 0) return;

  const uint8_t* ps = tile + TILE_PAD + (my_byte - g0); // ps[j] = base j of the item; ps[-1] exists
  const uint32_t h = H ? H : P.h;

  // ---- warm-up: k in-only steps over ps[-1 .. k-2]; ps[-1] is cancelled by the first roll ----
  State s = { 0u, 0u, 0u, 0u };
  uint32_t acc = ~0u; // bit 2 is set in the low word of all four seeds: survives iff all bases hashable
  {
    const uint4 e = tab_in[ps[-1]];
    fwd_step(s, e.x, e.y, 0u, 0u);
    rev_step(s, e.z, e.w, 0u, 0u);
  }
#pragma unroll 4
  for (uint32_t j = 0; j + 1 < k; ++j) {
    const uint4 e = tab_in[ps[j]];
    acc &= e.x;
    fwd_step(s, e.x, e.y, 0u, 0u);
    rev_step(s, e.z, e.w, 0u, 0u);
  }

  const uint8_t* pin = ps + (k - 1);
  const uint8_t* pout = ps - 1;
  auto roll = [&](uint32_t p) -> uint64_t {
    const uint4 ei = tab_in[pin[p]];
    const uint4 eo = tab_out[pout[p]];
    acc &= ei.x;
    fwd_step(s, ei.x, ei.y, eo.x, eo.y);
    rev_step(s, ei.z, ei.w, eo.z, eo.w);
    return canonical(s);
  };
  auto emit_one = [&](uint32_t p, uint64_t h0) {
    const uint64_t w = my_out + p;
    if (H == 1) {
      P.out[w] = h0;
    } else if (H == 2) {
      st_global_v2_u64(P.out + w * 2, h0, ext_hash(h0, P.mult[1]));
    } else if (H == 4) {
      st_global_v4_u64(P.out + w * 4, h0, ext_hash(h0, P.mult[1]), ext_hash(h0, P.mult[2]), ext_hash(h0, P.mult[3]));
    } else {
      uint64_t* o = P.out + w * h;
      o[0] = h0;
      for (uint32_t q = 1; q < h; ++q) o[q] = ext_hash(h0, ext_mult(q, k));
    }
    if (STRANDS) {
      P.out_fwd[w] = ((uint64_t)s.fhi << 32) | s.flo;
      P.out_rev[w] = ((uint64_t)s.rhi << 32) | s.rlo;
    }
  };

  if (P.reduce_out) { // fused consumer (only instantiated work when H == 0 && !STRANDS is launched with reduce_out)
    uint32_t run = 0, cnt = 0;
    uint64_t sum = 0, xr = 0;
    for (uint32_t j = 0; j + 1 < k; ++j) run = tab_in[ps[j]].x != 0 ? run + 1 : 0;
    for (uint32_t p = 0; p < n; ++p) {
      run = tab_in[pin[p]].x != 0 ? run + 1 : 0;
      const uint64_t h0 = roll(p);
      if (run >= k) {
        ++cnt;
        sum += h0;
        xr ^= h0;
        for (uint32_t q = 1; q < h; ++q) {
          const uint64_t e = ext_hash(h0, ext_mult(q, k));
          sum += e;
          xr ^= e;
        }
      }
    }
    atomicAdd(reinterpret_cast<unsigned long long*>(P.reduce_out), (unsigned long long)cnt);
    atomicAdd(reinterpret_cast<unsigned long long*>(P.reduce_out) + 1, (unsigned long long)sum);
    atomicXor(reinterpret_cast<unsigned long long*>(P.reduce_out) + 2, (unsigned long long)xr);
    return;
  }

  uint32_t p = 0;
  if (H == 1 && !STRANDS) {
    // four consecutive windows = one 32-byte sector; peel to sector alignment first
    const uint32_t peel = min(n, (uint32_t)((4 - (my_out & 3)) & 3));
    for (; p < peel; ++p) emit_one(p, roll(p));
    uint64_t* op = P.out + my_out + p;
    for (; p + 4 <= n; p += 4, op += 4) {
      const uint64_t a = roll(p), b = roll(p + 1), c = roll(p + 2), d = roll(p + 3);
      st_global_v4_u64(op, a, b, c, d);
    }
  }
  for (; p < n; ++p) emit_one(p, roll(p));

  if (!(acc & 4u)) scrub_item<H, STRANDS>(P, tab_in, ps, my_out, n);
}

template<int H, bool STRANDS>
cudaError_t launch_t(const KmerParams& P, uint32_t smem_bytes, cudaStream_t st)
{
  auto fn = kmer_kernel<H, STRANDS>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return e;
  const uint64_t ctas = (P.g.n_items + KMER_NT - 1) / KMER_NT;
  if (ctas == 0) return cudaSuccess;
  if (ctas > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  fn<<<(unsigned)ctas, KMER_NT, smem_bytes, st>>>(P);
  return cudaGetLastError();
}

} // namespace

uint32_t kmer_smem_bytes(uint32_t tile_cap) { return TILE_OFF + TILE_PAD + tile_cap + 16; }

cudaError_t launch_kmer(KmerParams P, cudaStream_t st)
{
  const uint64_t base[4] = { SEED_A, SEED_C, SEED_G, SEED_T };
  for (int x = 0; x < 4; ++x) {
    P.s[x] = base[x];
    P.sk[x] = srol_n(base[x], P.k);
  }
  for (unsigned q = 0; q < 4; ++q) P.mult[q] = ext_mult(q, P.k);
  if (P.packed) // 2-bit packed input: only the nibble-strip kernel reads it (the caller checked kmer_packed_direct_ok)
    return P.use_tma && kmer_fast_ok(P) ? launch_kmer_fast(P, st) : cudaErrorNotSupported;
  if (P.use_tma && kmer_fast_ok(P)) {
    const cudaError_t e = launch_kmer_fast(P, st);
    if (e != cudaErrorInvalidConfiguration) return e; // does not fit shared memory (huge k): the general kernel below
    cudaGetLastError();
  }
  if (P.bloom_mode) return cudaErrorNotSupported; // the Bloom consumer only exists in the fast kernel
  if (!P.general_fits) return cudaErrorInvalidConfiguration;
  const uint32_t smem = kmer_smem_bytes(P.tile_cap);
  const bool strands = P.out_fwd != nullptr;
  if (P.reduce_out) return launch_t<0, false>(P, smem, st);
  if (strands) {
    switch (P.h) {
      case 1: return launch_t<1, true>(P, smem, st);
      default: return launch_t<0, true>(P, smem, st);
    }
  }
  switch (P.h) {
    case 1: return launch_t<1, false>(P, smem, st);
    case 2: return launch_t<2, false>(P, smem, st);
    case 4: return launch_t<4, false>(P, smem, st);
    default: return launch_t<0, false>(P, smem, st);
  }
}

} // namespace nthb
