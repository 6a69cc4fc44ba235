/*
 * nthash_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's rolling-hash path.  See
 * nthash_oracle.h for scope and the parity-pinning statement.  Citations are
 * file:line in /root/reference (bcgsc/ntHash 2.4.0).
 *
 * Where this file restates instead of transliterating:
 *   - srol_table (internal.hpp:343-348) is evaluated by rotating the two
 *     sub-words arithmetically instead of through the 31/33-entry tables
 *     (internal.hpp:167-341); tests pin the equality against oracle/_ref.
 *   - base_forward_hash / base_reverse_hash (kmer.cpp:43-73,123-152) are
 *     evaluated one base at a time instead of through the 2/3/4-mer tables
 *     (internal.hpp:420-541); same closed form, pinned against oracle/_ref.
 *   - NtHash::init's loop bound `pos <= len-k+1` (kmer.cpp:232) reads one
 *     byte past the sequence; the restatement stops at `pos <= len-k`, which
 *     yields the same emissions without the over-read.
 *   - Raw bytes 1,3,4,5,7 (complement slots of SEED_TAB, internal.hpp:133)
 *     make the reference's NtHash disagree with itself (init uses
 *     CONVERT_TAB=255, roll uses SEED_TAB); here NtHash treats them as
 *     invalid bases.  SeedNtHash is self-consistent on them and is restated
 *     literally.
 */
#include "nthash_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* internal.hpp:124-128 */
#define NTO_SEED_A 0x3c8bfbb395c60474ULL
#define NTO_SEED_C 0x3193c18562a02b4cULL
#define NTO_SEED_G 0x20323ed082572324ULL
#define NTO_SEED_T 0x295549f54be24456ULL
/* internal.hpp:91,94 */
#define NTO_MULTISHIFT 27
#define NTO_MULTISEED 0x90b45d39fb6da1faULL

/* ------------------------------------------------------------------ L1 -- */

/* internal.hpp:41-47: rotate bits [63:33] (31 wide) and [32:0] (33 wide) left by one. */
uint64_t nto_srol(uint64_t x)
{
  uint64_t carry33 = (x >> 63) << 33; /* top of the 31-bit word wraps to bit 33 */
  uint64_t carry0 = (x >> 32) & 1ULL; /* top of the 33-bit word wraps to bit 0  */
  return ((x << 1) & ~(1ULL << 33)) | carry33 | carry0;
}

/* internal.hpp:83-88: exact inverse of nto_srol. */
uint64_t nto_sror(uint64_t x)
{
  uint64_t wrap63 = ((x >> 33) & 1ULL) << 63;
  uint64_t wrap32 = (x & 1ULL) << 32;
  return ((x >> 1) & ~(1ULL << 32) & ~(1ULL << 63)) | wrap63 | wrap32;
}

/* d-fold srol for any d: the halves have periods 31 and 33
 * (what MS_TAB_31L / MS_TAB_33R encode, internal.hpp:343-348). */
uint64_t nto_srol_n(uint64_t x, unsigned d)
{
  const uint64_t m33 = (1ULL << 33) - 1;
  uint64_t lo = x & m33;  /* 33-bit word */
  uint64_t hi = x >> 33;  /* 31-bit word */
  unsigned a = d % 33, b = d % 31;
  if (a) lo = ((lo << a) | (lo >> (33 - a))) & m33;
  if (b) hi = ((hi << b) | (hi >> (31 - b))) & ((1ULL << 31) - 1);
  return (hi << 33) | lo;
}

/* SEED_TAB, internal.hpp:132-165. */
uint64_t nto_seed(unsigned char c)
{
  switch (c) {
    case 'A': case 'a': case 4: case 5: return NTO_SEED_A;
    case 'C': case 'c': case 7: return NTO_SEED_C;
    case 'G': case 'g': case 3: return NTO_SEED_G;
    case 'T': case 't': case 'U': case 'u': case 1: return NTO_SEED_T;
    default: return 0;
  }
}

/* internal.hpp:343-348 */
uint64_t nto_srol_table(unsigned char c, unsigned d)
{
  return nto_srol_n(nto_seed(c), d);
}

/* internal.hpp:104-118 (note the precedence: i ^ (k * MULTISEED)). */
void nto_extend_hashes(uint64_t fwd, uint64_t rev, unsigned k, unsigned h, uint64_t* out)
{
  out[0] = fwd + rev; /* canonical(), internal.hpp:24-29 */
  for (unsigned i = 1; i < h; i++) {
    uint64_t t = out[0] * ((uint64_t)i ^ ((uint64_t)k * NTO_MULTISEED));
    t ^= t >> NTO_MULTISHIFT;
    out[i] = t;
  }
}

/* ------------------------------------------------------- NtHash (L2) ---- */

static int kmer_base_ok(unsigned char c)
{
  switch (c) {
    case 'A': case 'a': case 'C': case 'c': case 'G': case 'g':
    case 'T': case 't': case 'U': case 'u': return 1;
    default: return 0;
  }
}

typedef struct
{
  const unsigned char* s;
  size_t len;
  unsigned k, h;
  size_t pos;
  int initialized;
  uint64_t fwd, rev;
  uint64_t* hashes;
} kmer_it;

/* kmer.cpp:25-35: rightmost invalid byte of the window. */
static int kmer_window_invalid(const unsigned char* w, unsigned k, size_t* pos_n)
{
  for (size_t i = k; i-- > 0;) {
    if (!kmer_base_ok(w[i])) {
      *pos_n = i;
      return 1;
    }
  }
  return 0;
}

/* kmer.cpp:43-73 closed form: F = xor_i srol^(k-1-i)(S[c_i]). */
static uint64_t kmer_base_fwd(const unsigned char* w, unsigned k)
{
  uint64_t hv = 0;
  for (unsigned i = 0; i < k; i++) hv = nto_srol(hv) ^ nto_seed(w[i]);
  return hv;
}

/* kmer.cpp:123-152 closed form: R = xor_i srol^i(S[c_i & CP_OFF]). */
static uint64_t kmer_base_rev(const unsigned char* w, unsigned k)
{
  uint64_t hv = 0;
  for (unsigned i = k; i-- > 0;) hv = nto_srol(hv) ^ nto_seed(w[i] & 7);
  return hv;
}

/* kmer.cpp:228-244 */
static int kmer_init(kmer_it* it)
{
  size_t pos_n = 0;
  while (it->pos <= it->len - it->k && kmer_window_invalid(it->s + it->pos, it->k, &pos_n)) {
    it->pos += pos_n + 1;
  }
  if (it->pos > it->len - it->k) return 0;
  it->fwd = kmer_base_fwd(it->s + it->pos, it->k);
  it->rev = kmer_base_rev(it->s + it->pos, it->k);
  nto_extend_hashes(it->fwd, it->rev, it->k, it->h, it->hashes);
  it->initialized = 1;
  return 1;
}

/* kmer.cpp:246-264 with next_forward_hash :84-94 and next_reverse_hash :164-174. */
static int kmer_roll(kmer_it* it)
{
  if (!it->initialized) return kmer_init(it);
  if (it->pos >= it->len - it->k) return 0;
  unsigned char c_in = it->s[it->pos + it->k], c_out = it->s[it->pos];
  if (!kmer_base_ok(c_in)) {
    it->pos += it->k;
    return kmer_init(it);
  }
  it->fwd = nto_srol(it->fwd) ^ nto_seed(c_in) ^ nto_srol_table(c_out, it->k);
  it->rev = nto_sror(it->rev ^ nto_srol_table(c_in & 7, it->k) ^ nto_seed(c_out & 7));
  nto_extend_hashes(it->fwd, it->rev, it->k, it->h, it->hashes);
  ++it->pos;
  return 1;
}

size_t nto_kmer_read(const char* seq, size_t len, unsigned k, unsigned h, size_t pos0,
                     uint64_t* pos_out, uint64_t* hash_out, uint64_t* fwd_out,
                     uint64_t* rev_out, size_t cap)
{
  /* ctor checks, kmer.cpp:212-225; k<3 and h==0 are outside the reference's
   * working domain and rejected the same way. */
  if (k == 0 || h == 0 || len < k || pos0 > len - k) return (size_t)-1;
  kmer_it it = { (const unsigned char*)seq, len, k, h, pos0, 0, 0, 0, NULL };
  it.hashes = (uint64_t*)malloc(sizeof(uint64_t) * h);
  size_t n = 0;
  while (kmer_roll(&it)) {
    if (n < cap) {
      if (pos_out) pos_out[n] = it.pos;
      if (hash_out) memcpy(hash_out + n * h, it.hashes, sizeof(uint64_t) * h);
      if (fwd_out) fwd_out[n] = it.fwd;
      if (rev_out) rev_out[n] = it.rev;
    }
    n++;
  }
  free(it.hashes);
  return n;
}

/* kmer.cpp:338-364 */
void nto_blind_read(const char* kmer, unsigned k, unsigned h, const char* chars_in, size_t n_in,
                    uint64_t* hash0_out, uint64_t* hash_out, uint64_t* fwd_out, uint64_t* rev_out)
{
  unsigned char* win = (unsigned char*)malloc(k + n_in + 1);
  memcpy(win, kmer, k);
  uint64_t* hv = (uint64_t*)malloc(sizeof(uint64_t) * h);
  uint64_t fwd = kmer_base_fwd(win, k), rev = kmer_base_rev(win, k);
  nto_extend_hashes(fwd, rev, k, h, hv);
  if (hash0_out) memcpy(hash0_out, hv, sizeof(uint64_t) * h);
  for (size_t i = 0; i < n_in; i++) {
    unsigned char c_in = (unsigned char)chars_in[i], c_out = win[i]; /* deque front */
    fwd = nto_srol(fwd) ^ nto_seed(c_in) ^ nto_srol_table(c_out, k);
    rev = nto_sror(rev ^ nto_srol_table(c_in & 7, k) ^ nto_seed(c_out & 7));
    nto_extend_hashes(fwd, rev, k, h, hv);
    win[k + i] = c_in; /* push_back; the front advances with i */
    if (hash_out) memcpy(hash_out + i * h, hv, sizeof(uint64_t) * h);
    if (fwd_out) fwd_out[i] = fwd;
    if (rev_out) rev_out[i] = rev;
  }
  free(hv);
  free(win);
}

/* --------------------------------------------------- SeedNtHash (L2) ---- */

typedef struct
{
  unsigned n_blocks, n_monos;
  unsigned* blocks; /* pairs [start,end) */
  unsigned* monos;
} seed_shape;

/* seed.cpp:19-66 for one seed string of length k. */
static void seed_get_blocks(const char* seed, unsigned k, seed_shape* sh)
{
  unsigned* cb = (unsigned*)malloc(sizeof(unsigned) * 2 * (k + 2));
  unsigned* ib = (unsigned*)malloc(sizeof(unsigned) * 2 * (k + 2));
  unsigned* cm = (unsigned*)malloc(sizeof(unsigned) * (k + 2));
  unsigned* im = (unsigned*)malloc(sizeof(unsigned) * (k + 2));
  unsigned ncb = 0, nib = 0, ncm = 0, nim = 0;
  char pad = seed[k - 1] == '1' ? '0' : '1';
  unsigned start = 0;
  int care = seed[0] == '1';
  for (unsigned p = 0; p <= k; p++) {
    char ch = p < k ? seed[p] : pad;
    if (care && ch == '0') {
      if (p - start == 1) cm[ncm++] = start;
      else { cb[2 * ncb] = start; cb[2 * ncb + 1] = p; ncb++; }
      start = p;
      care = 0;
    } else if (!care && ch == '1') {
      if (p - start == 1) im[nim++] = start;
      else { ib[2 * nib] = start; ib[2 * nib + 1] = p; nib++; }
      start = p;
      care = 1;
    }
  }
  unsigned num_cares = ncb * 2 + ncm;
  unsigned num_ignores = nib * 2 + nim + 2;
  if (num_ignores < num_cares) { /* seed.cpp:52-58: whole window minus the ignored runs */
    ib[2 * nib] = 0; ib[2 * nib + 1] = k; nib++;
    sh->blocks = ib; sh->n_blocks = nib; sh->monos = im; sh->n_monos = nim;
    free(cb); free(cm);
  } else {
    sh->blocks = cb; sh->n_blocks = ncb; sh->monos = cm; sh->n_monos = ncm;
    free(ib); free(im);
  }
}

int nto_get_blocks(const char* seed, unsigned* blocks_out, unsigned* n_blocks, unsigned* monos_out,
                   unsigned* n_monos, unsigned cap)
{
  seed_shape sh;
  unsigned k = (unsigned)strlen(seed);
  if (k == 0) return -1;
  seed_get_blocks(seed, k, &sh);
  int rc = (sh.n_blocks > cap || sh.n_monos > cap) ? -1 : 0;
  if (rc == 0) {
    memcpy(blocks_out, sh.blocks, sizeof(unsigned) * 2 * sh.n_blocks);
    memcpy(monos_out, sh.monos, sizeof(unsigned) * sh.n_monos);
    *n_blocks = sh.n_blocks;
    *n_monos = sh.n_monos;
  }
  free(sh.blocks);
  free(sh.monos);
  return rc;
}

typedef struct
{
  const unsigned char* s;
  size_t len;
  unsigned k, m, h; /* m seeds, h hashes per seed */
  size_t pos;
  int initialized;
  const seed_shape* shapes;
  uint64_t *fh_nomonos, *rh_nomonos, *fh, *rh, *hashes;
} seed_it;

static void seed_finish(seed_it* it, unsigned i_seed, uint64_t f, uint64_t r)
{
  it->fh[i_seed] = f;
  it->rh[i_seed] = r;
  nto_extend_hashes(f, r, it->k, it->h, it->hashes + (size_t)i_seed * it->h); /* seed.cpp:167-172 */
}

/* seed.cpp:130-175 (base ntmsm64): false + loc_n on a NUL byte at a block position. */
static int seed_base(seed_it* it, const unsigned char* w, unsigned* loc_n)
{
  for (unsigned i = 0; i < it->m; i++) {
    const seed_shape* sh = &it->shapes[i];
    uint64_t f = 0, r = 0;
    for (unsigned b = 0; b < sh->n_blocks; b++) {
      for (unsigned p = sh->blocks[2 * b]; p < sh->blocks[2 * b + 1]; p++) {
        if (w[p] == 0) { *loc_n = p; return 0; }
        f ^= nto_srol_table(w[p], it->k - 1 - p);
        r ^= nto_srol_table(w[p] & 7, p);
      }
    }
    it->fh_nomonos[i] = f;
    it->rh_nomonos[i] = r;
    for (unsigned q = 0; q < sh->n_monos; q++) {
      unsigned p = sh->monos[q];
      f ^= nto_srol_table(w[p], it->k - 1 - p);
      r ^= nto_srol_table(w[p] & 7, p);
    }
    seed_finish(it, i, f, r);
  }
  return 1;
}

/* seed.cpp:177-207 + :230-249 (forward-roll ntmsm64); w is the PREVIOUS window, w[k] the new base. */
static void seed_step(seed_it* it, const unsigned char* w)
{
  for (unsigned i = 0; i < it->m; i++) {
    const seed_shape* sh = &it->shapes[i];
    uint64_t f = nto_srol(it->fh_nomonos[i]);
    uint64_t r = it->rh_nomonos[i];
    for (unsigned b = 0; b < sh->n_blocks; b++) {
      unsigned i_out = sh->blocks[2 * b], i_in = sh->blocks[2 * b + 1];
      unsigned char c_out = w[i_out], c_in = w[i_in];
      f ^= nto_srol_table(c_out, it->k - i_out);
      f ^= nto_srol_table(c_in, it->k - i_in);
      r ^= nto_srol_table(c_out & 7, i_out);
      r ^= nto_srol_table(c_in & 7, i_in);
    }
    r = nto_sror(r);
    it->fh_nomonos[i] = f;
    it->rh_nomonos[i] = r;
    for (unsigned q = 0; q < sh->n_monos; q++) {
      unsigned p = sh->monos[q];
      f ^= nto_srol_table(w[p + 1], it->k - 1 - p);
      r ^= nto_srol_table(w[p + 1] & 7, p);
    }
    seed_finish(it, i, f, r);
  }
}

/* seed.cpp:493-516 */
static int seed_init(seed_it* it)
{
  unsigned pos_n = 0;
  while (it->pos < it->len - it->k + 1 && !seed_base(it, it->s + it->pos, &pos_n)) {
    it->pos += pos_n + 1;
  }
  if (it->pos > it->len - it->k) return 0;
  it->initialized = 1;
  return 1;
}

/* seed.cpp:518-544 */
static int seed_roll(seed_it* it)
{
  if (!it->initialized) return seed_init(it);
  if (it->pos >= it->len - it->k) return 0;
  if (nto_seed(it->s[it->pos + it->k]) == 0) {
    it->pos += it->k;
    return seed_init(it);
  }
  seed_step(it, it->s + it->pos);
  ++it->pos;
  return 1;
}

static seed_shape* shapes_new(const char* const* seeds, unsigned n_seeds, unsigned k)
{
  /* check_seeds, seed.cpp:85-104: length mismatch is the only hard error. */
  for (unsigned i = 0; i < n_seeds; i++)
    if (strlen(seeds[i]) != k) return NULL;
  seed_shape* sh = (seed_shape*)malloc(sizeof(seed_shape) * n_seeds);
  for (unsigned i = 0; i < n_seeds; i++) seed_get_blocks(seeds[i], k, &sh[i]);
  return sh;
}

static void shapes_free(seed_shape* sh, unsigned n_seeds)
{
  for (unsigned i = 0; i < n_seeds; i++) { free(sh[i].blocks); free(sh[i].monos); }
  free(sh);
}

static void seed_it_open(seed_it* it, const unsigned char* s, size_t len, unsigned k, unsigned m,
                         unsigned h, size_t pos0, const seed_shape* shapes)
{
  it->s = s; it->len = len; it->k = k; it->m = m; it->h = h; it->pos = pos0;
  it->initialized = 0; it->shapes = shapes;
  uint64_t* buf = (uint64_t*)calloc((size_t)m * (4 + h), sizeof(uint64_t));
  it->fh_nomonos = buf; it->rh_nomonos = buf + m; it->fh = buf + 2 * m; it->rh = buf + 3 * m;
  it->hashes = buf + 4 * m;
}

size_t nto_seed_read(const char* seq, size_t len, const char* const* seeds, unsigned n_seeds,
                     unsigned h, unsigned k, size_t pos0, uint64_t* pos_out, uint64_t* hash_out,
                     uint64_t* fwd_out, uint64_t* rev_out, size_t cap)
{
  if (k == 0 || h == 0 || n_seeds == 0 || len < k || pos0 > len - k) return (size_t)-1;
  seed_shape* sh = shapes_new(seeds, n_seeds, k);
  if (!sh) return (size_t)-1;
  seed_it it;
  seed_it_open(&it, (const unsigned char*)seq, len, k, n_seeds, h, pos0, sh);
  size_t n = 0, H = (size_t)n_seeds * h;
  while (seed_roll(&it)) {
    if (n < cap) {
      if (pos_out) pos_out[n] = it.pos;
      if (hash_out) memcpy(hash_out + n * H, it.hashes, sizeof(uint64_t) * H);
      if (fwd_out) memcpy(fwd_out + n * n_seeds, it.fh, sizeof(uint64_t) * n_seeds);
      if (rev_out) memcpy(rev_out + n * n_seeds, it.rh, sizeof(uint64_t) * n_seeds);
    }
    n++;
  }
  free(it.fh_nomonos);
  shapes_free(sh, n_seeds);
  return n;
}

/* ------------------------------------------------------------ batches --- */

typedef struct
{
  const char* bases;
  const uint64_t* read_off;
  const uint64_t* koff; /* NULL when no dense output is wanted */
  uint64_t r0, r1;
  unsigned k, h, m;
  const seed_shape* shapes; /* NULL => NtHash */
  uint64_t *out, *out_fwd, *out_rev;
  uint8_t* valid;
  uint64_t n_emit, sum, xr;
} batch_job;

static void* batch_worker(void* arg)
{
  batch_job* j = (batch_job*)arg;
  const unsigned k = j->k, h = j->h, m = j->m ? j->m : 1;
  const size_t H = (size_t)m * h;
  uint64_t* hv = (uint64_t*)malloc(sizeof(uint64_t) * H);
  uint64_t n_emit = 0, sum = 0, xr = 0;
  for (uint64_t r = j->r0; r < j->r1; r++) {
    const unsigned char* s = (const unsigned char*)j->bases + j->read_off[r];
    size_t len = (size_t)(j->read_off[r + 1] - j->read_off[r]);
    if (len < k) continue;
    uint64_t base = j->koff ? j->koff[r] : 0;
    if (j->shapes) {
      seed_it it;
      seed_it_open(&it, s, len, k, m, h, 0, j->shapes);
      while (seed_roll(&it)) {
        for (size_t q = 0; q < H; q++) { sum += it.hashes[q]; xr ^= it.hashes[q]; }
        n_emit++;
        uint64_t o = base + it.pos;
        if (j->out) memcpy(j->out + o * H, it.hashes, sizeof(uint64_t) * H);
        if (j->valid) j->valid[o] = 1;
        if (j->out_fwd) memcpy(j->out_fwd + o * m, it.fh, sizeof(uint64_t) * m);
        if (j->out_rev) memcpy(j->out_rev + o * m, it.rh, sizeof(uint64_t) * m);
      }
      free(it.fh_nomonos);
    } else {
      kmer_it it = { s, len, k, h, 0, 0, 0, 0, hv };
      while (kmer_roll(&it)) {
        for (size_t q = 0; q < H; q++) { sum += hv[q]; xr ^= hv[q]; }
        n_emit++;
        uint64_t o = base + it.pos;
        if (j->out) memcpy(j->out + o * H, hv, sizeof(uint64_t) * H);
        if (j->valid) j->valid[o] = 1;
        if (j->out_fwd) j->out_fwd[o] = it.fwd;
        if (j->out_rev) j->out_rev[o] = it.rev;
      }
    }
  }
  free(hv);
  j->n_emit = n_emit; j->sum = sum; j->xr = xr;
  return NULL;
}

static uint64_t run_batch(const char* bases, const uint64_t* read_off, uint64_t n_reads, unsigned k,
                          unsigned h, unsigned m, const seed_shape* shapes, uint64_t* out,
                          uint8_t* valid, uint64_t* out_fwd, uint64_t* out_rev, int n_threads,
                          uint64_t* sum_out, uint64_t* xor_out)
{
  uint64_t* koff = NULL;
  const unsigned mm = m ? m : 1;
  if (out || valid || out_fwd || out_rev) {
    koff = (uint64_t*)malloc(sizeof(uint64_t) * (n_reads + 1));
    uint64_t acc = 0;
    for (uint64_t r = 0; r < n_reads; r++) {
      koff[r] = acc;
      uint64_t len = read_off[r + 1] - read_off[r];
      if (len >= k) acc += len - k + 1;
    }
    koff[n_reads] = acc;
    /* positions the reference never visits read back as 0 / invalid */
    if (out) memset(out, 0, sizeof(uint64_t) * acc * mm * h);
    if (valid) memset(valid, 0, acc);
    if (out_fwd) memset(out_fwd, 0, sizeof(uint64_t) * acc * mm);
    if (out_rev) memset(out_rev, 0, sizeof(uint64_t) * acc * mm);
  }
  if (n_threads < 1) n_threads = 1;
  if ((uint64_t)n_threads > n_reads) n_threads = n_reads ? (int)n_reads : 1;
  batch_job* jobs = (batch_job*)calloc((size_t)n_threads, sizeof(batch_job));
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
  for (int t = 0; t < n_threads; t++) {
    batch_job* j = &jobs[t];
    j->bases = bases; j->read_off = read_off; j->koff = koff;
    j->r0 = n_reads * (uint64_t)t / (uint64_t)n_threads;
    j->r1 = n_reads * (uint64_t)(t + 1) / (uint64_t)n_threads;
    j->k = k; j->h = h; j->m = m; j->shapes = shapes;
    j->out = out; j->valid = valid; j->out_fwd = out_fwd; j->out_rev = out_rev;
    if (n_threads == 1) batch_worker(j);
    else pthread_create(&th[t], NULL, batch_worker, j);
  }
  uint64_t n_emit = 0, sum = 0, xr = 0;
  for (int t = 0; t < n_threads; t++) {
    if (n_threads > 1) pthread_join(th[t], NULL);
    n_emit += jobs[t].n_emit; sum += jobs[t].sum; xr ^= jobs[t].xr;
  }
  free(th); free(jobs); free(koff);
  if (sum_out) *sum_out = sum;
  if (xor_out) *xor_out = xr;
  return n_emit;
}

uint64_t nto_kmer_batch(const char* bases, const uint64_t* read_off, uint64_t n_reads, unsigned k,
                        unsigned h, uint64_t* out, uint8_t* valid, uint64_t* out_fwd,
                        uint64_t* out_rev, int n_threads, uint64_t* sum_out, uint64_t* xor_out)
{
  if (k == 0 || h == 0) return (uint64_t)-1;
  return run_batch(bases, read_off, n_reads, k, h, 0, NULL, out, valid, out_fwd, out_rev,
                   n_threads, sum_out, xor_out);
}

uint64_t nto_seed_batch(const char* bases, const uint64_t* read_off, uint64_t n_reads,
                        const char* const* seeds, unsigned n_seeds, unsigned h, unsigned k,
                        uint64_t* out, uint8_t* valid, uint64_t* out_fwd, uint64_t* out_rev,
                        int n_threads, uint64_t* sum_out, uint64_t* xor_out)
{
  if (k == 0 || h == 0 || n_seeds == 0) return (uint64_t)-1;
  seed_shape* sh = shapes_new(seeds, n_seeds, k);
  if (!sh) return (uint64_t)-1;
  uint64_t n = run_batch(bases, read_off, n_reads, k, h, n_seeds, sh, out, valid, out_fwd, out_rev,
                         n_threads, sum_out, xor_out);
  shapes_free(sh, n_seeds);
  return n;
}

/* SURVEY.md Appendix C generator. */
void nto_gen_bases(char* dst, uint64_t n, uint64_t seed)
{
  uint64_t state = seed;
  for (uint64_t i = 0; i < n;) {
    uint64_t z = (state += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    z ^= z >> 31;
    for (int b = 0; b < 32 && i < n; b++, i++) {
      dst[i] = "ACGT"[z & 3];
      z >>= 2;
    }
  }
}
