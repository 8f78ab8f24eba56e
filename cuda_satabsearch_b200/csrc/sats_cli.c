/*
 * sats_cli.c -- drop-in `cudaSaTabsearch` command line over the libsats C ABI (include/sats.h).
 *
 * Keeps the reference driver's contract (stivalaa/cuda_satabsearch nvcc_src_current/cudaSaTabsearch.cu):
 *   argv   -c | -q dbfile | -r restarts                    (:605-626)
 *   stdin  db path / "LTYPE LORDER LSOLN" / query structures   (:667-693), or with -q a list of ids (:631-664)
 *   stdout per (pool, query): three '#' lines then "name rawscore norm2score z-score p-value"
 *          rows (+ SSE map pairs when LSOLN=T), all queries on the small pool (order <= 96) first,
 *          then all queries on the large pool (:987-1115, :1196-1269, :1272-1310)
 *   stderr diagnostics and timings; errors are "message + exit status 1"
 * Row formatting follows the `-c` path (:442-454), which is what the north star pins parity to.
 *
 * Differences, all deliberate: -c is refused (this build has no CPU search path); -q computes norm2 with the
 * query's own order like the reference's CPU branch (:366, :380) rather than the GPU branch's orders[qi]
 * (:996-998); the db file may also be a packed SATSDB1 cache.  Extra options:
 *   -g N      N GPUs (devices are shared when N exceeds the GPUs present).  Philox: N cost-weighted shards of the
 *             size-sorted db.  xorwow: the db is replicated and GPU g runs the reference blocks b with b % N == g
 *             (cudaSaTabsearch_kernel.cu:932), so the output equals the single-GPU validation run
 *   -R mode   philox (default) | xorwow  -- xorwow = the reference GPU run's 128x128 cuRAND streams
 *   -A mode   table (default) | fast     -- Metropolis thresholds: host libm table | device fast-math
 *   -s seed   RNG seed (default 1234)
 *   -L        also search database structures of order 112..128 (the reference drops everything above 111)
 *   -k N      print only the N best-scoring structures of each (query, pool) block, best first (selected on the
 *             device with sats_search_topk, so only N rows per query leave each GPU); LSOLN = F
 *   -z Z      print only the structures whose z-score (4th column) is >= Z, in database order of decreasing size.  The
 *             cut is bound to the searchers (sats_search_bind_cut), so the kernels stream the hits into a device list
 *             while they run and only the hits are copied back (bytes reported on stderr); LSOLN = F, not together with -k
 */
#include <getopt.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "sats.h"

#define MAX_GPUS 16

static double now_ms(void)
{
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec * 1e3 + t.tv_nsec * 1e-6;
}

static void die(const char *what)
{
  fprintf(stderr, "%s: %s\n", what, sats_last_error());
  exit(1);
}

/* one searcher per GPU, created side by side: each builds and uploads its own part of the database */
typedef struct { const sats_db *db; int device, rank, count, rc; sats_searcher *sr; char err[512]; } create_job;
static void *create_worker(void *arg)
{
  create_job *j = (create_job *)arg;
  j->rc = sats_searcher_create(j->db, j->device, j->rank, j->count, &j->sr);
  if (j->rc != SATS_OK) { strncpy(j->err, sats_last_error(), sizeof j->err - 1); j->err[sizeof j->err - 1] = 0; }   /* the message is thread-local */
  return NULL;
}

/* one selected row while the shards' hit lists are merged */
typedef struct { int32_t index, score, order; } hit_t;
static int by_device_order(const void *a, const void *b)
{
  const hit_t *x = (const hit_t *)a, *y = (const hit_t *)b;
  if (x->order != y->order) return x->order > y->order ? -1 : 1;     /* larger structures first ... */
  return (x->index > y->index) - (x->index < y->index);              /* ... then file order */
}
static int by_score_then_device_order(const void *a, const void *b)
{
  const hit_t *x = (const hit_t *)a, *y = (const hit_t *)b;
  if (x->score != y->score) return x->score > y->score ? -1 : 1;
  return by_device_order(a, b);
}

static void usage(const char *prog)
{
  fprintf(stderr, "Usage: %s [-c] [-q dbfile] [-r restarts] [-g gpus] [-R philox|xorwow] [-A table|fast] [-s seed] [-k tophits] [-z zmin] [-L]\n", prog);
  fprintf(stderr, "  -c : (reference: run on host CPU) not available in this build\n");
  fprintf(stderr, "  -q dbfile : database is read from dbfile, list of query\n"
                  "              ids is read from stdin\n");
  fprintf(stderr, "   -r restarts : number of restarts. Default %d\n", SATS_DEFAULT_MAXSTART);
  exit(1);
}

static char *read_all(FILE *fp, size_t *len)
{
  size_t cap = 1 << 16, n = 0;
  char *buf = (char *)malloc(cap);
  if (!buf) return NULL;
  for (;;) {
    size_t got = fread(buf + n, 1, cap - n, fp);
    n += got;
    if (got == 0) break;
    if (n == cap) {
      cap *= 2;
      char *nb = (char *)realloc(buf, cap);
      if (!nb) { free(buf); return NULL; }
      buf = nb;
    }
  }
  *len = n;
  return buf;
}

static int load_db(const char *path, int max_order, sats_db **db)
{
  FILE *fp = fopen(path, "rb");
  char magic[8] = {0};
  if (!fp) { fprintf(stderr, "ERROR opening db file %s\n", path); exit(1); }
  size_t got = fread(magic, 1, 8, fp);
  fclose(fp);
  if (got == 8 && memcmp(magic, "SATSDB1", 8) == 0) return sats_db_read_packed(path, db);
  return sats_db_read_ascii_ext(path, max_order, db);
}

int main(int argc, char *argv[])
{
  char dbfile[4096] = "";
  int querydbmode = 0, maxstart = SATS_DEFAULT_MAXSTART, ngpus = 1, c;
  int rng_mode = SATS_RNG_PHILOX, accept_mode = SATS_ACCEPT_HOST_TABLE, topk = 0, max_order = SATS_MAXDIM, zcut = 0;
  double zmin = 0.0;
  unsigned long long seed = SATS_REF_SEED;
  int flags[3] = {1, 1, 0};
  sats_db *db = NULL, *queries = NULL;

  while ((c = getopt(argc, argv, "cq:r:g:R:A:s:k:z:L")) != -1) {
    switch (c) {
      case 'c':
        fprintf(stderr, "ERROR: -c (host CPU search) is not available: this build is GPU-only\n");
        exit(1);
      case 'q': querydbmode = 1; strncpy(dbfile, optarg, sizeof(dbfile) - 1); break;
      case 'r': maxstart = atoi(optarg); break;
      case 'g': ngpus = atoi(optarg); break;
      case 'R':
        if (!strcmp(optarg, "philox")) rng_mode = SATS_RNG_PHILOX;
        else if (!strcmp(optarg, "xorwow")) rng_mode = SATS_RNG_XORWOW_GRID;
        else usage(argv[0]);
        break;
      case 'A':
        if (!strcmp(optarg, "table")) accept_mode = SATS_ACCEPT_HOST_TABLE;
        else if (!strcmp(optarg, "fast")) accept_mode = SATS_ACCEPT_DEVICE_FAST;
        else usage(argv[0]);
        break;
      case 's': seed = strtoull(optarg, NULL, 0); break;
      case 'k': topk = atoi(optarg); break;
      case 'z': zcut = 1; zmin = atof(optarg); break;
      case 'L': max_order = SATS_MAXDIM_EXT; break;
      default: usage(argv[0]);
    }
  }
  if (maxstart < 1) { fprintf(stderr, "ERROR: restarts must be >= 1\n"); exit(1); }
  if (ngpus < 1 || ngpus > MAX_GPUS) { fprintf(stderr, "ERROR: -g must be 1..%d\n", MAX_GPUS); exit(1); }
  const int split_blocks = ngpus > 1 && rng_mode == SATS_RNG_XORWOW_GRID;      /* replicate the db, split the reference blocks */
  if (accept_mode == SATS_ACCEPT_DEVICE_FAST && rng_mode != SATS_RNG_XORWOW_GRID) {
    fprintf(stderr, "ERROR: -A fast (the reference GPU build's fast-math acceptance) goes with -R xorwow\n");
    exit(1);
  }
  if (topk < 0) { fprintf(stderr, "ERROR: -k needs a positive count\n"); exit(1); }
  if (zcut && topk > 0) { fprintf(stderr, "ERROR: -z cannot be combined with -k\n"); exit(1); }
  fprintf(stderr, "MAXDIM = %d\n", max_order);

  size_t inlen = 0;
  char *input = read_all(stdin, &inlen);
  if (!input) { fprintf(stderr, "ERROR reading stdin\n"); exit(1); }

  char *ids = NULL;
  int num_queries = 0;
  if (querydbmode) {
    flags[0] = 1; flags[1] = 1; flags[2] = 0;
    int cap = 1;
    for (size_t i = 0; i < inlen; i++) cap += input[i] == '\n';
    ids = (char *)calloc((size_t)cap + 1, 9);
    num_queries = sats_idlist_parse(input, inlen, ids, cap + 1);
    if (num_queries < 0) die("ERROR reading query ids");
  } else {
    if (sats_input_parse(input, inlen, dbfile, sizeof dbfile, flags, &queries) != SATS_OK) {
      fprintf(stderr, "%s\n", sats_last_error());
      fprintf(stderr, "ERROR loading query structures from stdin\n");
      exit(1);
    }
    num_queries = sats_db_count(queries);
    fprintf(stderr, "Read %d query structures\n", num_queries);
  }
  if (!flags[0]) {
    fprintf(stderr, "WARNING: LTYPE is always set to T\n");
    flags[0] = 1;
  }
  const int lorder = flags[1], lsoln = flags[2];
  if ((topk > 0 || zcut) && lsoln) { fprintf(stderr, "ERROR: -k / -z cannot be combined with LSOLN = T\n"); exit(1); }

  fprintf(stderr, "Loading database...\n");
  double t0 = now_ms();
  if (load_db(dbfile, max_order, &db) != SATS_OK) { fprintf(stderr, "%s\n", sats_last_error()); fprintf(stderr, "ERROR loading database\n"); exit(1); }
  const int dbsize = sats_db_count(db);
  int *small_idx = (int *)malloc(sizeof(int) * (size_t)(dbsize + 1));
  int *large_idx = (int *)malloc(sizeof(int) * (size_t)(dbsize + 1));
  int nsmall = 0, nlarge = 0;
  for (int i = 0; i < dbsize; i++) {
    if (sats_db_order(db, i) <= SATS_MAXDIM_GPU) small_idx[nsmall++] = i;
    else large_idx[nlarge++] = i;
  }
  fprintf(stderr, "Loaded %d db entries (%d order > %d) in %f ms\n", dbsize, nlarge, SATS_MAXDIM_GPU, now_ms() - t0);

  if (querydbmode) {
    fprintf(stderr, "Building query index list...\n");
    int *qidx = (int *)malloc(sizeof(int) * (size_t)(num_queries + 1));
    for (int i = 0; i < num_queries; i++) {
      qidx[i] = sats_db_find(db, ids + (size_t)i * 9);
      if (qidx[i] < 0) { fprintf(stderr, "%s\n", sats_last_error()); exit(1); }
    }
    if (sats_db_select(db, qidx, num_queries, &queries) != SATS_OK) die("ERROR building query set");
    free(qidx);
  }
  if (num_queries == 0) { fprintf(stderr, "ERROR: no query structures found on stdin\n"); exit(1); }

  int have = sats_device_count();
  if (have < 1) { fprintf(stderr, "ERROR: no CUDA device found (this build has no CPU search path)\n"); exit(1); }
  if (ngpus > have) fprintf(stderr, "WARNING: %d shards requested, %d GPU(s) present: shards share devices\n", ngpus, have);
  fprintf(stderr, "maxstart = %d\n", maxstart);
  fprintf(stderr, "Copying database to device...\n");
  t0 = now_ms();
  sats_searcher *sr[MAX_GPUS];
  {
    create_job jobs[MAX_GPUS];
    pthread_t tid[MAX_GPUS];
    /* contexts one after the other (the driver serialises their creation and concurrent attempts are slower), then
     * the searchers -- blob build + upload -- side by side */
    for (int g = 0; g < ngpus && g < have; g++)
      if (sats_device_init(g) != SATS_OK) die("ERROR initialising device");
    fprintf(stderr, "Initialised %d device context(s) in %f ms\n", ngpus < have ? ngpus : have, now_ms() - t0);
    for (int g = 0; g < ngpus; g++) {
      jobs[g].db = db; jobs[g].device = g % have; jobs[g].rank = split_blocks ? 0 : g; jobs[g].count = split_blocks ? 1 : ngpus;
      jobs[g].sr = NULL; jobs[g].rc = SATS_OK; jobs[g].err[0] = 0;
      if (g > 0 && pthread_create(&tid[g], NULL, create_worker, &jobs[g]) != 0) { fprintf(stderr, "ERROR starting a worker thread\n"); exit(1); }
    }
    create_worker(&jobs[0]);
    for (int g = 1; g < ngpus; g++) pthread_join(tid[g], NULL);
    for (int g = 0; g < ngpus; g++) {
      if (jobs[g].rc != SATS_OK) { fprintf(stderr, "ERROR creating searcher: %s\n", jobs[g].err); exit(1); }
      sr[g] = jobs[g].sr;
    }
  }
  fprintf(stderr, "Copied %d entries to %d GPU(s) in %f ms\n", dbsize, ngpus, now_ms() - t0);

  sats_params prm;
  sats_params_default(&prm);
  prm.lorder = lorder; prm.lsoln = lsoln; prm.restarts = maxstart;
  prm.rng_mode = rng_mode; prm.accept_mode = accept_mode; prm.seed = seed;
  prm.pool_threshold = SATS_MAXDIM_GPU;

  /* queries are processed in chunks that bound the result buffers */
  size_t per_query = (size_t)dbsize * (lsoln ? (SATS_MAP_STRIDE + 1) : 1) * sizeof(int32_t);
  int chunk = (int)((256u << 20) / (per_query ? per_query : 1));
  if (chunk < 1) chunk = 1;
  if (chunk > num_queries) chunk = num_queries;
  int32_t *scores = (int32_t *)malloc(sizeof(int32_t) * (size_t)chunk * (size_t)(dbsize + 1));
  int32_t *maps = lsoln ? (int32_t *)malloc(sizeof(int32_t) * (size_t)chunk * (size_t)(dbsize + 1) * SATS_MAP_STRIDE) : NULL;
  size_t outcap = 1 << 20;
  char *out = (char *)malloc(outcap);
  if (!scores || (lsoln && !maps) || !out) { fprintf(stderr, "malloc scores failed\n"); exit(1); }

  for (int pass = 0; pass < 2; pass++) {
    const int *idx = pass == 0 ? small_idx : large_idx;
    const int n = pass == 0 ? nsmall : nlarge;
    if (pass == 1 && n == 0) break;
    prm.pool = pass == 0 ? SATS_POOL_SMALL : SATS_POOL_LARGE;
    for (int q0 = 0; q0 < num_queries; q0 += chunk) {
      const int nq = (num_queries - q0) < chunk ? (num_queries - q0) : chunk;
      double t1 = now_ms();
      for (int g = 0; g < ngpus; g++)
        if (sats_search_upload(sr[g], queries, q0, nq) != SATS_OK) die("ERROR uploading queries");
      for (int g = 0; g < ngpus; g++) {
        if (zcut && sats_search_bind_cut(sr[g], zmin) != SATS_OK) die("ERROR binding the significance cut");
        if (split_blocks) { prm.grid_rank = g; prm.grid_count = ngpus; }
        if (sats_search_launch(sr[g], &prm, (uint32_t)q0, NULL) != SATS_OK) die("kernel launch failed");
      }
      /* hits-only modes: hcap rows per query come back from the device instead of one score per database entry */
      const int hits_only = topk > 0 || zcut;
      const int hcap = topk > 0 ? topk : (n > 0 ? n : 1);
      int32_t *top_idx = NULL, *top_sc = NULL, *top_n = NULL;
      if (hits_only) {
        top_idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)nq * (size_t)hcap);
        top_sc = (int32_t *)malloc(sizeof(int32_t) * (size_t)nq * (size_t)hcap);
        top_n = (int32_t *)malloc(sizeof(int32_t) * (size_t)nq);
        if (!top_idx || !top_sc || !top_n) { fprintf(stderr, "malloc failed\n"); exit(1); }
        /* every shard selects on its own device; the host merges the shards' short lists in the single-GPU row order
         * (-k: score descending, ties like the device order: larger structures first, then file order; -z: device order) */
        hit_t *cand = (hit_t *)malloc(sizeof(hit_t) * (size_t)ngpus * (size_t)hcap);
        int32_t *g_idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)nq * (size_t)hcap);
        int32_t *g_sc = (int32_t *)malloc(sizeof(int32_t) * (size_t)nq * (size_t)hcap);
        int32_t *all_idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)ngpus * (size_t)nq * (size_t)hcap);
        int32_t *all_sc = (int32_t *)malloc(sizeof(int32_t) * (size_t)ngpus * (size_t)nq * (size_t)hcap);
        if (!cand || !g_idx || !g_sc || !all_idx || !all_sc) { fprintf(stderr, "malloc failed\n"); exit(1); }
        for (int g = 0; g < ngpus; g++) {
          if (topk > 0) { if (sats_search_topk(sr[g], topk, g_idx, g_sc) != SATS_OK) die("ERROR selecting top hits"); }
          else {
            int64_t nbytes = 0;
            if (sats_search_streamed_hits(sr[g], hcap, top_n, g_idx, g_sc, &nbytes) != SATS_OK) die("ERROR selecting significant hits");
            fprintf(stderr, "GPU %d streamed its hits: %lld bytes device -> host for %d queries (%f per query; the dense scores would be %lld)\n",
                    g, (long long)nbytes, nq, (double)nbytes / nq, (long long)nq * (long long)sats_searcher_entry_count(sr[g]) * 4);
          }
          memcpy(all_idx + (size_t)g * nq * hcap, g_idx, sizeof(int32_t) * (size_t)nq * (size_t)hcap);
          memcpy(all_sc + (size_t)g * nq * hcap, g_sc, sizeof(int32_t) * (size_t)nq * (size_t)hcap);
        }
        for (int q = 0; q < nq; q++) {
          int nc = 0;
          for (int g = 0; g < ngpus; g++)
            for (int i = 0; i < hcap; i++) {
              const int32_t e = all_idx[((size_t)g * nq + q) * hcap + i];
              if (e < 0) break;
              cand[nc].index = e; cand[nc].score = all_sc[((size_t)g * nq + q) * hcap + i]; cand[nc].order = sats_db_order(db, e);
              nc++;
            }
          qsort(cand, (size_t)nc, sizeof(hit_t), topk > 0 ? by_score_then_device_order : by_device_order);
          for (int i = 0; i < hcap; i++) {
            top_idx[(size_t)q * hcap + i] = i < nc ? cand[i].index : -1;
            top_sc[(size_t)q * hcap + i] = i < nc ? cand[i].score : 0;
          }
        }
        free(cand); free(g_idx); free(g_sc); free(all_idx); free(all_sc);
      } else {
        for (int g = 0; g < ngpus; g++)          /* every GPU's copy in flight before the first wait */
          if (sats_search_collect_begin(sr[g]) != SATS_OK) die("ERROR collecting results");
        for (int g = 0; g < ngpus; g++)
          if (sats_search_collect(sr[g], scores, maps) != SATS_OK) die("ERROR collecting results");
      }
      double ms = now_ms() - t1;
      fprintf(stderr, "GPU execution time %f ms (%d queries x %d entries, %s pool)\n", ms, nq, n, pass ? "large" : "small");
      fprintf(stderr, "%f million iterations/sec\n", ((double)nq * n * ((double)maxstart * SATS_MAXITER) / (ms / 1000)) / 1.0e6);
      for (int q = 0; q < nq && hits_only; q++) {
        /* hits only: scatter the selected scores into this query's row and print them in rank order */
        int32_t *sc = scores + (size_t)q * dbsize;
        int nhit = 0;
        while (nhit < hcap && top_idx[(size_t)q * hcap + nhit] >= 0) { sc[top_idx[(size_t)q * hcap + nhit]] = top_sc[(size_t)q * hcap + nhit]; nhit++; }
        size_t need = sats_format_block(out, outcap, sats_db_name(queries, q0 + q), sats_db_order(queries, q0 + q), dbfile,
                                        lorder, 0, db, top_idx + (size_t)q * hcap, nhit, sc, NULL);
        if (need >= outcap) {
          outcap = need + 1;
          out = (char *)realloc(out, outcap);
          if (!out) { fprintf(stderr, "malloc failed\n"); exit(1); }
          sats_format_block(out, outcap, sats_db_name(queries, q0 + q), sats_db_order(queries, q0 + q), dbfile, lorder, 0, db,
                            top_idx + (size_t)q * hcap, nhit, sc, NULL);
        }
        fwrite(out, 1, need, stdout);
      }
      free(top_idx); free(top_sc); free(top_n);
      for (int q = 0; q < nq && !hits_only; q++) {
        const int32_t *sc = scores + (size_t)q * dbsize;
        const int32_t *mp = lsoln ? maps + (size_t)q * dbsize * SATS_MAP_STRIDE : NULL;
        size_t need = sats_format_block(out, outcap, sats_db_name(queries, q0 + q), sats_db_order(queries, q0 + q),
                                        dbfile, lorder, lsoln, db, idx, n, sc, mp);
        if (need >= outcap) {
          outcap = need + 1;
          out = (char *)realloc(out, outcap);
          if (!out) { fprintf(stderr, "malloc failed\n"); exit(1); }
          sats_format_block(out, outcap, sats_db_name(queries, q0 + q), sats_db_order(queries, q0 + q), dbfile, lorder,
                            lsoln, db, idx, n, sc, mp);
        }
        fwrite(out, 1, need, stdout);
      }
    }
  }
  for (int g = 0; g < ngpus; g++) sats_searcher_free(sr[g]);
  sats_db_free(db);
  sats_db_free(queries);
  free(scores); free(maps); free(out); free(small_idx); free(large_idx); free(ids); free(input);
  return 0;
}
