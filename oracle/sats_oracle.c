/*
 * sats_oracle.c -- CPU restatement of the SA tableau-search hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (cuda_satabsearch_b200/, the CLI,
 * the C-ABI library) may include, link or call this file.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() use it, as the checker.
 *
 * Parity status: PINNED.  In drand48 mode this restatement reproduces, bit for bit,
 *   - the unmodified reference binary (oracle/_ref/cudaSaTabsearch_ref -c) on every fixture,
 *   - the reference's captured 2013 job output old/nvcc_src_cuda5/cpu_cudaSaTabsearch.o1462445
 *     (pool threshold 32), see tests/test_oracle_golden.py and tests/golden/.
 *
 * What it restates (paths relative to the reference's nvcc_src_current/):
 *   zeta()            <- tscord        cudaSaTabsearch_kernel.cu:306-332
 *   full_score()      <- tmscord       cudaSaTabsearch_kernel.cu:396-440
 *   delta_score()     <- deltasd       cudaSaTabsearch_kernel.cu:502-535
 *   seed_matching()   <- thinit        cudaSaTabsearch_kernel.cu:588-648
 *   pick_candidate()  <- randtypeind   cudaSaTabsearch_kernel.cu:677-714
 *   anneal_restart()  <- restart body  cudaSaTabsearch_kernel.cu:1014-1191
 *   search drivers    <- db/restart loops and block arg-max, cudaSaTabsearch_kernel.cu:932-1233
 *   uniform sources   <- drand48 (kernel.cu:623,708,1040,1164; seed cudaSaTabsearch.cu:871),
 *                        cuRAND XORWOW (cudaSaTabsearch.cu:258-264), and the product's Philox4x32-10
 *   gumbel statistics <- gumbelstats.c:50-94, gumbelstats.h:27-28
 *
 * Three uniform sources, one annealing body:
 *   (i)   SATS_ORNG_DRAND48  the reference's `-c` path: one process-global glibc drand48 stream, consumed
 *         entry -> restart -> draw in program order, each value narrowed to float.
 *   (ii)  SATS_ORNG_XORWOW   the reference's GPU decomposition: a grid of B x T logical threads (128 x 128 in
 *         the reference), stream id = T*b + t initialised like curand_init(seed, id, 0); block b walks
 *         entries b, b+B, ...; thread t runs restarts t, t+T, ... (restart count rounded up to a multiple
 *         of T); per-thread running maximum; block arg-max with lowest-t tie-break; states persist.
 *   (iii) SATS_ORNG_PHILOX   the product's production mode: chain (query, entry id, restart) draws from
 *         Philox4x32-10 at *static* positions, so every chain is independent of every other and of the
 *         launch geometry.  Stream layout "v2" (counter word 0 selects the block):
 *           seeding pass  block 0x80000000: one random BIT per query SSE (bit i&31 of word i>>5); the draw is
 *                         0.25f when the bit is set and 0.75f when it is clear (thinit only compares it with 0.5)
 *           move m        block m>>1, words 2(m&1) and 2(m&1)+1:
 *                         u1 (SSE pick)        = unit(word a)
 *                         u2 (candidate pick)  = unit(word a * 0x9E3779B9)   -- drawn only with >= 2 candidates, like
 *                                                the reference; the pick uses the leading bits of word a, the Fibonacci
 *                                                hash spreads the remaining ones
 *                         u3 (Metropolis test) = unit(word b)
 *         i.e. one Philox block per two moves and one per seeding pass.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_MAXDIM 128      /* working arrays: saparams.h:15 has 111; entries up to 128 serve the opt-in large-order tests */
#define ORACLE_MAPSTRIDE 111    /* row stride of the SSE-map output = the reference's MAXDIM (kernel.cu:1231) */
#define ORACLE_MOVES 100       /* saparams.h:31 MAXITER */
#define ORACLE_T0 10.0f        /* saparams.h:34 */
#define ORACLE_COOL 0.95f      /* saparams.h:37 */
#define ORACLE_GATE 4.0f       /* saparams.h:28 MXSSED */
#define ORACLE_SEEDPROB 0.5    /* saparams.h:43 INIT_MATCHPROB (a double literal) */
#define ORACLE_EPS 1.1e-7      /* cudaSaTabsearch_kernel.cu:67 (a double literal) */
#define ORACLE_NEG_INIT (-99999) /* cudaSaTabsearch_kernel.cu:1009 */

enum { SATS_ORNG_DRAND48 = 0, SATS_ORNG_XORWOW = 1, SATS_ORNG_PHILOX = 2 };

/* ------------------------------------------------------------------ uniform sources */

typedef struct {
  int kind;
  uint32_t *xw;            /* XORWOW: d, v[0..4] of the logical thread that owns the chain */
  uint32_t key[2];         /* Philox key = seed */
  uint32_t chain[3];       /* Philox counter words 1..3 = restart, entry id, query index */
  uint32_t cached_block;   /* Philox: which 4-draw block `cache` holds (0xffffffff = none) */
  uint32_t cache[4];
} usrc_t;

/* XORWOW step, Marsaglia 2003 "Xorshift RNGs" p.5 as used by cuRAND (curand_kernel.h:863-874). */
static uint32_t xorwow_next(uint32_t *s)
{
  uint32_t t = s[1] ^ (s[1] >> 2);
  s[1] = s[2]; s[2] = s[3]; s[3] = s[4]; s[4] = s[5];
  s[5] = (s[5] ^ (s[5] << 4)) ^ (t ^ (t << 1));
  s[0] += 362437u;
  return s[5] + s[0];
}

/* Philox4x32-10, Salmon et al. SC'11; same constants as Random123 / cuRAND. */
void sats_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int round = 0; round < 10; round++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 32 random bits -> float in (0,1], cuRAND's _curand_uniform (curand_uniform.h:69-72).  The product
 * of a float with 2^-32 is exact, so fused and unfused evaluation agree bit for bit. */
static float u32_to_unit(uint32_t x)
{
  return (float)x * 2.3283064e-10f + 1.1641532e-10f;
}

#define DOMAIN_SEED 0x80000000u  /* Philox counter word 0: top bit separates the seeding pass from the moves */
#define PHILOX_U2_HASH 0x9E3779B9u /* 2^32 / golden ratio: Fibonacci hashing of the pick word into the candidate draw */

static float draw(usrc_t *r, uint32_t domain, uint32_t position)
{
  switch (r->kind) {
  case SATS_ORNG_DRAND48:
    return (float)drand48();           /* reference narrows the double into `float randnum` */
  case SATS_ORNG_XORWOW:
    return u32_to_unit(xorwow_next(r->xw));
  default: {
    /* `position` is the draw's logical position: query SSE i in the seeding pass, 3*move + slot in the moves */
    uint32_t block = domain == DOMAIN_SEED ? DOMAIN_SEED : (position / 3u) >> 1;
    if (block != r->cached_block) {
      uint32_t ctr[4] = { block, r->chain[0], r->chain[1], r->chain[2] };
      sats_oracle_philox4x32_10(ctr, r->key, r->cache);
      r->cached_block = block;
    }
    if (domain == DOMAIN_SEED)
      return ((r->cache[(position >> 5) & 3u] >> (position & 31u)) & 1u) ? 0.25f : 0.75f;
    uint32_t move = position / 3u, slot = position % 3u;
    uint32_t a = r->cache[2u * (move & 1u)], b = r->cache[2u * (move & 1u) + 1u];
    if (slot == 0u) return u32_to_unit(a);
    if (slot == 1u) return u32_to_unit(a * PHILOX_U2_HASH);
    return u32_to_unit(b);
  }
  }
}

/* (u - EPS) * n evaluated in double and truncated toward zero; a (theoretical) u == 0 maps to 0. */
static int scaled_index(float u, int n)
{
  double x = ((double)u - ORACLE_EPS) * (double)n;
  return x <= 0.0 ? 0 : (int)x;
}

/* ------------------------------------------------------------------ scoring */

typedef struct {
  int n1, n2;
  const uint8_t *qt; const float *qd; int qp;   /* query tableau / distances, row pitch qp elements */
  const uint8_t *et; const float *ed; int ep;   /* entry tableau / distances, row pitch ep elements */
  uint8_t qtype[ORACLE_MAXDIM], etype[ORACLE_MAXDIM];
  int lorder, lsoln;
} pair_t;

int sats_oracle_zeta(int x, int y)
{
  int same_hi = ((x ^ y) & 0xF0) == 0;
  int same_lo = ((x ^ y) & 0x0F) == 0;
  int hits = same_hi + same_lo;
  return hits ? hits : -2;
}

static int term(const pair_t *p, int i, int k, int j, int l)
{
  float gap = fabsf(p->qd[i * p->qp + k] - p->ed[j * p->ep + l]);
  if (!(gap <= ORACLE_GATE)) return 0;
  return sats_oracle_zeta(p->qt[i * p->qp + k], p->et[j * p->ep + l]);
}

static int full_score(const pair_t *p, const int *m)
{
  int total = 0;
  for (int i = 0; i < p->n1; i++) {
    if (m[i] < 0) continue;
    for (int k = i + 1; k < p->n1; k++)
      if (m[k] >= 0) total += term(p, i, k, m[i], m[k]);
  }
  return total;
}

static int delta_score(const pair_t *p, const int *m, int i, int from, int to)
{
  int d = 0;
  for (int k = 0; k < p->n1; k++) {
    int l = m[k];
    if (l < 0 || k == i) continue;
    if (from >= 0 && l != from) d -= term(p, i, k, from, l);
    if (to >= 0 && l != to) d += term(p, i, k, to, l);
  }
  return d;
}

/* ------------------------------------------------------------------ annealing */

static void seed_matching(const pair_t *p, usrc_t *r, int *m, int *owner)
{
  for (int i = 0; i < p->n1; i++) m[i] = -1;
  for (int j = 0; j < p->n2; j++) owner[j] = -1;
  int j = 0;
  for (int i = 0; i < p->n1; i++) {
    float u = draw(r, DOMAIN_SEED, (uint32_t)i);
    if (!((double)u < ORACLE_SEEDPROB)) continue;
    while (j < p->n2 && p->etype[j] != p->qtype[i]) j++;
    if (j >= p->n2) return;              /* the rest of the query draws nothing */
    m[i] = j; owner[j] = i; j++;
  }
}

static int pick_candidate(const pair_t *p, usrc_t *r, uint32_t pos, const int *owner,
                          int lo, int hi, int type)
{
  int cand[ORACLE_MAXDIM], n = 0;
  for (int j = lo; j < hi; j++)
    if (p->etype[j] == type && owner[j] < 0) cand[n++] = j;
  if (n == 0) return -1;
  if (n == 1) return cand[0];
  return cand[scaled_index(draw(r, 0u, pos), n)];
}

/* One restart: seeding pass, full score, ORACLE_MOVES Metropolis moves.  `best`/`bestmap` carry the
 * running maximum of whoever owns this restart (a CPU run, a logical GPU thread, or a single chain). */
static void anneal_restart(const pair_t *p, usrc_t *r, int *best, int *bestmap)
{
  int m[ORACLE_MAXDIM], owner[ORACLE_MAXDIM];
  seed_matching(p, r, m, owner);
  int score = full_score(p, m);
  if (score > *best) {
    *best = score;
    memcpy(bestmap, m, sizeof(int) * (size_t)p->n1);
  }
  float temp = ORACLE_T0;
  for (int mv = 0; mv < ORACLE_MOVES; mv++) {
    uint32_t base = 3u * (uint32_t)mv;
    int i = scaled_index(draw(r, 0u, base), p->n1);
    int lo, hi;
    if (p->lorder) {
      lo = -1;
      for (int k = i; k >= 0 && lo < 0; k--) lo = m[k];
      if (lo < 0) lo = p->n2;                         /* nothing mapped at or before i: empty window */
      if (i == p->n1 - 1) hi = p->n2;
      else {
        hi = -1;                                      /* nothing mapped after i: empty window */
        for (int k = i + 1; k < p->n1 && hi < 0; k++) hi = m[k];
      }
    } else { lo = 0; hi = p->n2; }
    int to = pick_candidate(p, r, base + 1u, owner, lo, hi, p->qtype[i]);
    int from = m[i];
    int d = delta_score(p, m, i, from, to);
    int cand_score = score + d;
    if (cand_score > *best) {
      *best = cand_score;
      if (p->lsoln) {
        memcpy(bestmap, m, sizeof(int) * (size_t)p->n1);
        bestmap[i] = to;
      }
    }
    float u = draw(r, 0u, base + 2u);
    if (expf((float)d / temp) > u) {
      score = cand_score;
      if (from >= 0) owner[from] = -1;
      if (to >= 0) owner[to] = i;
      m[i] = to;
    }
    temp *= ORACLE_COOL;
  }
}

/* ------------------------------------------------------------------ database view + drivers */

/* Entry e is a dense order[e] x order[e] matrix pair starting at element offset off[e] of tabs / dmats. */
typedef struct {
  int count;
  const int32_t *order;
  const int64_t *off;
  const uint8_t *tabs;
  const float *dmats;
} dbview_t;

static void bind_query(pair_t *p, int n1, const uint8_t *qtab, const float *qdmat, int lorder, int lsoln)
{
  memset(p, 0, sizeof(*p));
  p->n1 = n1; p->qt = qtab; p->qd = qdmat; p->qp = n1;
  p->lorder = lorder; p->lsoln = lsoln;
  for (int i = 0; i < n1; i++) p->qtype[i] = qtab[i * n1 + i];
}

static void bind_entry(pair_t *p, const dbview_t *db, int e)
{
  int n2 = db->order[e];
  p->n2 = n2; p->ep = n2;
  p->et = db->tabs + db->off[e];
  p->ed = db->dmats + db->off[e];
  for (int j = 0; j < n2; j++) p->etype[j] = p->et[j * n2 + j];
}

static void emit(const pair_t *p, int e, int best, const int *bestmap, int32_t *outscore, int32_t *outmap)
{
  outscore[e] = best;
  if (outmap)
    for (int i = 0; i < p->n1; i++) outmap[(size_t)e * ORACLE_MAPSTRIDE + i] = bestmap[i];
}

void sats_oracle_srand48(long seed) { srand48(seed); }

/* (i) the reference's `-c` path for one query against one pool (sa_tabsearch_host with grid = block = 1). */
int sats_oracle_search_drand48(int n1, const uint8_t *qtab, const float *qdmat,
                               int count, const int32_t *order, const int64_t *off,
                               const uint8_t *tabs, const float *dmats,
                               int lorder, int lsoln, int restarts,
                               int32_t *outscore, int32_t *outmap)
{
  if (n1 < 1 || n1 > ORACLE_MAPSTRIDE) return -1;
  dbview_t db = { count, order, off, tabs, dmats };
  pair_t p; bind_query(&p, n1, qtab, qdmat, lorder, lsoln);
  usrc_t r; memset(&r, 0, sizeof r); r.kind = SATS_ORNG_DRAND48;
  for (int e = 0; e < count; e++) {
    if (order[e] < 1 || order[e] > ORACLE_MAXDIM) return -2;
    bind_entry(&p, &db, e);
    int best = ORACLE_NEG_INIT, bestmap[ORACLE_MAXDIM];
    for (int i = 0; i < n1; i++) bestmap[i] = -1;
    for (int s = 0; s < restarts; s++) anneal_restart(&p, &r, &best, bestmap);
    emit(&p, e, best, bestmap, outscore, lsoln ? outmap : NULL);
  }
  return 0;
}

/* ---- XORWOW stream initialisation: curand_init(seed, subsequence, 0) restated from the published
 * construction (curand_kernel.h:795-826): salted seed scramble, then jump 2^67 * subsequence steps.
 * The jump is done with our own GF(2) matrix powers of the one-step map (no cuRAND tables). */

typedef struct { uint32_t row[160][5]; } gf2mat_t;   /* row i = image of unit vector e_i */

static void gf2_apply(const gf2mat_t *a, const uint32_t v[5], uint32_t out[5])
{
  uint32_t acc[5] = { 0, 0, 0, 0, 0 };
  for (int i = 0; i < 160; i++)
    if (v[i >> 5] & (1u << (i & 31)))
      for (int w = 0; w < 5; w++) acc[w] ^= a->row[i][w];
  memcpy(out, acc, sizeof acc);
}

static void gf2_square(const gf2mat_t *a, gf2mat_t *out)
{
  gf2mat_t tmp;
  for (int i = 0; i < 160; i++) gf2_apply(a, a->row[i], tmp.row[i]);
  *out = tmp;
}

static void xorwow_jump_tables(gf2mat_t *pow2 /* [k] = step^(2^67 * 2^k) */, int levels)
{
  gf2mat_t step;
  for (int i = 0; i < 160; i++) {
    uint32_t s[6] = { 0, 0, 0, 0, 0, 0 };
    s[1 + (i >> 5)] = 1u << (i & 31);
    (void)xorwow_next(s);
    memcpy(step.row[i], &s[1], 5 * sizeof(uint32_t));
  }
  for (int sq = 0; sq < 67; sq++) gf2_square(&step, &step);
  pow2[0] = step;
  for (int k = 1; k < levels; k++) gf2_square(&pow2[k - 1], &pow2[k]);
}

/* states: nstates x 6 words (d, v0..v4); stream i == curand_init(seed, i, 0). */
void sats_oracle_xorwow_init(uint32_t *states, int nstates, uint64_t seed)
{
  enum { LEVELS = 32 };
  static gf2mat_t pow2[LEVELS];
  static int ready = 0;
  if (!ready) { xorwow_jump_tables(pow2, LEVELS); ready = 1; }
  uint32_t s0 = (uint32_t)seed ^ 0xaad26b49u;
  uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
  uint32_t t0 = 1099087573u * s0, t1 = 2591861531u * s1;
  for (int i = 0; i < nstates; i++) {
    uint32_t *s = states + (size_t)i * 6;
    s[0] = 6615241u + t1 + t0;
    s[1] = 123456789u + t0;
    s[2] = 362436069u ^ t0;
    s[3] = 521288629u + t1;
    s[4] = 88675123u ^ t1;
    s[5] = 5783321u + t0;
    uint32_t sub = (uint32_t)i;
    for (int k = 0; sub; k++, sub >>= 1)
      if (sub & 1u) gf2_apply(&pow2[k], &s[1], &s[1]);
    /* the Weyl word d advances by 362437 * 2^67 * i == 0 (mod 2^32) */
  }
}

uint32_t sats_oracle_xorwow_next(uint32_t *state6) { return xorwow_next(state6); }

/* (ii) the reference's GPU decomposition for one query against one pool (sa_tabsearch_gpu<<<B,T>>>). */
int sats_oracle_search_xorwow_grid(int n1, const uint8_t *qtab, const float *qdmat,
                                   int count, const int32_t *order, const int64_t *off,
                                   const uint8_t *tabs, const float *dmats,
                                   int lorder, int lsoln, int restarts,
                                   uint32_t *states, int nblocks, int nthreads,
                                   int32_t *outscore, int32_t *outmap)
{
  if (n1 < 1 || n1 > ORACLE_MAPSTRIDE) return -1;
  dbview_t db = { count, order, off, tabs, dmats };
  pair_t p; bind_query(&p, n1, qtab, qdmat, lorder, lsoln);
  int (*maps)[ORACLE_MAXDIM] = malloc(sizeof(int[ORACLE_MAXDIM]) * (size_t)nthreads);
  int *best = malloc(sizeof(int) * (size_t)nthreads);
  if (!maps || !best) { free(maps); free(best); return -3; }
  for (int b = 0; b < nblocks; b++) {
    for (int e = b; e < count; e += nblocks) {
      bind_entry(&p, &db, e);
      for (int t = 0; t < nthreads; t++) {
        usrc_t r; memset(&r, 0, sizeof r);
        r.kind = SATS_ORNG_XORWOW;
        r.xw = states + ((size_t)b * nthreads + t) * 6;
        best[t] = ORACLE_NEG_INIT;
        for (int i = 0; i < n1; i++) maps[t][i] = -1;
        for (int s = 0; s < restarts; s += nthreads) anneal_restart(&p, &r, &best[t], maps[t]);
      }
      int win = 0;
      for (int t = 1; t < nthreads; t++) if (best[t] > best[win]) win = t;
      emit(&p, e, best[win], maps[win], outscore, lsoln ? outmap : NULL);
    }
  }
  free(maps); free(best);
  return 0;
}

/* (iii) production semantics: exactly `restarts` independent chains per entry, chain r keyed by
 * (seed, query_index, entry_id[e], r); entry result = max over chains, lowest r wins ties. */
int sats_oracle_search_philox(int n1, const uint8_t *qtab, const float *qdmat,
                              int count, const int32_t *order, const int64_t *off,
                              const uint8_t *tabs, const float *dmats, const int32_t *entry_id,
                              int lorder, int lsoln, int restarts,
                              uint64_t seed, uint32_t query_index,
                              int32_t *outscore, int32_t *outmap)
{
  if (n1 < 1 || n1 > ORACLE_MAPSTRIDE) return -1;
  dbview_t db = { count, order, off, tabs, dmats };
  pair_t p; bind_query(&p, n1, qtab, qdmat, lorder, lsoln);
  for (int e = 0; e < count; e++) {
    bind_entry(&p, &db, e);
    int best = ORACLE_NEG_INIT, bestmap[ORACLE_MAXDIM];
    for (int i = 0; i < n1; i++) bestmap[i] = -1;
    for (int s = 0; s < restarts; s++) {
      usrc_t r; memset(&r, 0, sizeof r);
      r.kind = SATS_ORNG_PHILOX;
      r.key[0] = (uint32_t)seed; r.key[1] = (uint32_t)(seed >> 32);
      r.chain[0] = (uint32_t)s;
      r.chain[1] = (uint32_t)(entry_id ? entry_id[e] : e);
      r.chain[2] = query_index;
      r.cached_block = 0xffffffffu;
      int cbest = ORACLE_NEG_INIT, cmap[ORACLE_MAXDIM];
      for (int i = 0; i < n1; i++) cmap[i] = -1;
      anneal_restart(&p, &r, &cbest, cmap);
      if (cbest > best) { best = cbest; memcpy(bestmap, cmap, sizeof(int) * (size_t)n1); }
    }
    emit(&p, e, best, bestmap, outscore, lsoln ? outmap : NULL);
  }
  return 0;
}

/* ------------------------------------------------------------------ building blocks, exported for unit tests */

int sats_oracle_full_score(int n1, const uint8_t *qtab, const float *qdmat,
                           int n2, const uint8_t *etab, const float *edmat, const int32_t *map)
{
  pair_t p; bind_query(&p, n1, qtab, qdmat, 1, 0);
  p.n2 = n2; p.ep = n2; p.et = etab; p.ed = edmat;
  int m[ORACLE_MAXDIM];
  for (int i = 0; i < n1; i++) m[i] = map[i];
  return full_score(&p, m);
}

int sats_oracle_delta_score(int n1, const uint8_t *qtab, const float *qdmat,
                            int n2, const uint8_t *etab, const float *edmat, const int32_t *map,
                            int i, int from, int to)
{
  pair_t p; bind_query(&p, n1, qtab, qdmat, 1, 0);
  p.n2 = n2; p.ep = n2; p.et = etab; p.ed = edmat;
  int m[ORACLE_MAXDIM];
  for (int k = 0; k < n1; k++) m[k] = map[k];
  return delta_score(&p, m, i, from, to);
}

/* Acceptance threshold the reference evaluates at move `mv` for score change `d`:
 * expf((float)d / T_mv) with T_0 = 10 and T <- 0.95f*T in fp32 (kernel.cu:1030,1166,1189). */
float sats_oracle_accept_threshold(int mv, int d)
{
  float temp = ORACLE_T0;
  for (int k = 0; k < mv; k++) temp *= ORACLE_COOL;
  return expf((float)d / temp);
}

/* What the product's integer thresholds must reproduce, as functions of the raw 32 random bits:
 * the SSE / candidate pick (kernel.cu:1042, :710) and the Metropolis test (kernel.cu:1166) on unit(x). */
int sats_oracle_pick_from_bits(uint32_t x, int n) { return scaled_index(u32_to_unit(x), n); }
int sats_oracle_seed_attempt_from_bits(uint32_t x) { return (double)u32_to_unit(x) < ORACLE_SEEDPROB; }
int sats_oracle_accept_from_bits(int mv, int d, uint32_t x)
{
  return sats_oracle_accept_threshold(mv, d) > u32_to_unit(x);
}

/* ------------------------------------------------------------------ Gumbel statistics (gumbelstats.c:50-94) */

double sats_oracle_norm2(int score, int n1, int n2) { return 2.0 * score / (double)(n1 + n2); }

double sats_oracle_zscore(double norm2score)
{
  const double a = 0.3780327676087335, b = 0.3582596175507505;   /* gumbelstats.h:27-28 */
  const double euler = 0.5772156649015328606;
  int x = (int)norm2score;              /* z_gumbel takes an int: the normalised score is truncated */
  return (x - (a + b * euler)) / ((M_PI / sqrt(6.0)) * b);
}

double sats_oracle_pvalue(double z)
{
  const double euler = 0.5772156649015328606;
  return 1 - exp(-exp(-((M_PI / sqrt(6.0)) * z + euler)));
}
