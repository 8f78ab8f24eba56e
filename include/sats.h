/*
 * sats.h -- C ABI of the B200-native SA tableau-search library (libsats.so).
 *
 * This is the drop-in boundary for the reference's hot path (stivalaa/cuda_satabsearch,
 * paths below relative to its nvcc_src_current/).  Every entry point names the reference
 * interface it replaces.  Plain C: opaque handles, plain pointers and sizes, no C++/torch types.
 *
 * Error convention: the reference prints to stderr and exit(1)s (parsetableaux.c:341-343,
 * cudaSaTabsearch.cu:702-706, checkCudaErrors).  A library cannot do that, so every function that
 * can fail returns SATS_OK (0) or a negative sats_status and leaves a human-readable message in
 * sats_last_error() (thread-local).  The CLI (csrc/sats_cli.c) turns a failure back into
 * "message on stderr + exit status 1".
 *
 * There is NO CPU search path in this library (north star): sats_search*() fails with
 * SATS_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef SATS_H
#define SATS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants (saparams.h:15-46) ------------------------------------------------------------ */
#define SATS_MAXDIM 111          /* MAXDIM: largest structure order the reference handles          */
#define SATS_MAXDIM_EXT 128      /* largest DATABASE structure order this library can search (opt-in)  */
#define SATS_MAXDIM_GPU 96       /* MAXDIM_GPU: the reference's small/large pool split             */
#define SATS_LABELSIZE 8         /* LABELSIZE: identifier length                                   */
#define SATS_MAXITER 100         /* MAXITER: moves per restart                                     */
#define SATS_DEFAULT_MAXSTART 128/* DEFAULT_MAXSTART                                               */
#define SATS_REF_GRID_BLOCKS 128 /* cudaSaTabsearch.cu:877: blocks of the reference launch         */
#define SATS_REF_GRID_THREADS 128/* cudaSaTabsearch.cu:878: threads per block                      */
#define SATS_REF_SEED 1234       /* cudaSaTabsearch.cu:263 (curand_init) and :871 (srand48)        */
#define SATS_MAP_STRIDE 111      /* row stride of the SSE-map output, = MAXDIM (kernel.cu:1231)    */

typedef enum sats_status {
  SATS_OK = 0,
  SATS_ERR_ARG = -1,      /* bad argument                                                          */
  SATS_ERR_IO = -2,       /* cannot open / read / write a file                                     */
  SATS_ERR_PARSE = -3,    /* malformed input (bad tableau code, bad helix type, truncated entry)   */
  SATS_ERR_NOMEM = -4,
  SATS_ERR_CUDA = -5,     /* CUDA runtime error, or no usable device                               */
  SATS_ERR_NOTFOUND = -6  /* -q query id not in the database (cudaSaTabsearch.cu:775-779)          */
} sats_status;

const char *sats_last_error(void);
const char *sats_version(void);

/* ---- structures: databases and query sets (replaces parsetableaux.h:29-39) ------------------- */
/* A sats_db is an ordered list of structures in ORIGINAL FILE ORDER: name (<= 8 chars), order n,
 * symmetric n x n tableau of 1-byte codes (hi nibble P0 R1 O2 L3 ?4, lo nibble E0 D1 S2 T3 ?4,
 * diagonal = SSE type e0 xa1 xi2 xg3; parsetableaux.c:52-138) and symmetric n x n fp32 distance
 * matrix in Angstrom.  Query sets use the same type.  Structures of order > SATS_MAXDIM are
 * skipped with a warning on stderr, as read_database/read_queries do (parsetableaux.c:457-465). */
typedef struct sats_db sats_db;

/* read_database(FILE*, ...) / read_queries(FILE*, ...): ASCII text from a file or from memory.   */
int sats_db_read_ascii(const char *path, sats_db **out);
int sats_db_parse_ascii(const char *text, size_t len, sats_db **out);
/* the same, but keeping database structures of order up to max_order (<= SATS_MAXDIM_EXT = 128) instead of
 * dropping everything above SATS_MAXDIM like the reference (parsetableaux.c:457-465).  Opt-in because it adds rows to
 * the output; queries stay limited to SATS_MAXDIM (the SSE-map row stride).                                         */
int sats_db_read_ascii_ext(const char *path, int max_order, sats_db **out);
int sats_db_parse_ascii_ext(const char *text, size_t len, int max_order, sats_db **out);
/* The reference's stdin grammar in non -q mode (cudaSaTabsearch.cu:667-693): line 1 db path,
 * line 2 "LTYPE LORDER LSOLN" as T/F, then query structures.  flags_tf[3] receives 0/1.         */
int sats_input_parse(const char *text, size_t len, char *dbfile, size_t dbfile_cap, int flags_tf[3],
                     sats_db **queries);
/* -q mode id list (cudaSaTabsearch.cu:631-664): one id per line, cut to 7 characters as the
 * reference does; ids_out receives count x 9 bytes.  Returns the count or a negative status.     */
int sats_idlist_parse(const char *text, size_t len, char *ids_out, int max_ids);

/* Build from caller arrays: entry e is a dense order[e] x order[e] row-major pair starting at
 * element offset off[e] of tabs / dmats; names is count x 9 bytes (NUL padded).                  */
int sats_db_from_arrays(int count, const int32_t *order, const char *names, const int64_t *off,
                        const uint8_t *tabs, const float *dmats, sats_db **out);
void sats_db_free(sats_db *db);

int sats_db_count(const sats_db *db);
int sats_db_order(const sats_db *db, int index);
const char *sats_db_name(const sats_db *db, int index);
int sats_db_max_order(const sats_db *db);
/* copies the dense order x order matrices of one structure into caller buffers                   */
int sats_db_get(const sats_db *db, int index, uint8_t *tab, float *dmat);
/* linear strcasecmp search as cudaSaTabsearch.cu:741-780; index or SATS_ERR_NOTFOUND            */
int sats_db_find(const sats_db *db, const char *name);
/* new db holding structures index[0..count) of src (used for -q queries and for sharding tests)  */
int sats_db_select(const sats_db *src, const int32_t *index, int count, sats_db **out);
/* Synthetic database generator of SURVEY section 8(d)/D: bootstrap-resample whole entries of src
 * to `count` entries (xorshift64* seeded with `seed`), rename them s%06d, optionally stable-sort
 * by order (scripts/convdb2.py:182-184).                                                         */
int sats_db_bootstrap(const sats_db *src, int count, uint64_t seed, int sort_by_order, sats_db **out);

/* ASCII writer with scripts/convdb2.py:182-231 semantics ("%6s %4d" header, "%6.3f " cells,
 * blank line between entries) and the packed binary cache ("SATSDB1").           */
int sats_db_write_ascii(const sats_db *db, const char *path);
int sats_db_write_packed(const sats_db *db, const char *path);
int sats_db_read_packed(const char *path, sats_db **out);

/* ---- query construction (replaces the pure core of scripts/pytableaucreate.py) -----------------------------
 * From the C-alpha traces of a protein's secondary-structure elements to a searchable structure, the way the reference's
 * scripts do it (Kamat & Lesk 2007; Konagurthu, Stuckey & Lesk 2008):
 *   sats_tabcode_from_angle  angle_to_tabcode (scripts/pttableau.py:434-469): omega in (-pi, pi] -> "PE", "OT", ...; every
 *                            interval is half-open on the left; out of range / NaN (ValueError there) -> SATS_ERR_ARG
 *   sats_relative_angle      PTNode.relative_angle (scripts/ptnode.py:752-880) over LineLineIntersect (scripts/geometry.py:
 *                            18-79): interaxial angle of two axes given as (centroid, direction cosines); returns 1 where
 *                            the reference returns None (parallel or degenerate axes), 0 with *omega set otherwise
 *   sats_build_structure     compute_tableau (scripts/pttableau.py:470-520, use_hk = False) + compute_sse_midpoint_dist_matrix
 *                            (scripts/ptdistmatrix.py:1014-1066) for n SSEs of type sse_type[i] (0 strand, 1 alpha, 2 pi,
 *                            3 3-10 helix), centroid / dircos = n x 3 doubles: codes from the pairwise angles ("??" where
 *                            there is none, "PE" with a warning for a NaN angle), midpoint distances rounded through
 *                            "%6.3f" as the writer does (scripts/convdb2.py:225; > 99.9 clamped as
 *                            scripts/pytableaucreate.py:114-116).  The result is a one-structure sats_db, usable as a query.
 *   sats_fit_axis            PTNodeHelix.fit_axis / PTNodeStrand.fit_axis (scripts/ptnode.py:1113-1292, :1846-1990): the axis
 *                            of one SSE from its C-alpha trace (n_res x 3 doubles, N- to C-terminus) as a total least squares
 *                            line through the triple-plane midpoints (helix, sse_type 1..3) or pair midpoints (strand, 0),
 *                            oriented N -> C; returns 1 where the reference returns None (helix < 3 residues, strand < 2)
 *   sats_build_structure_from_ca   the two together: n SSEs, n_res[i] residues each, C-alpha coordinates concatenated; an
 *                            SSE without an axis keeps "??" codes and 0.000 distances, as the reference's None propagates.
 * What remains outside is the PDB / DSSP front end that decides which residues form which SSE (Bio.PDB, DSSP / STRIDE).     */
int sats_tabcode_from_angle(double omega, char code[3]);
int sats_relative_angle(const double c_self[3], const double d_self[3], const double c_other[3],
                        const double d_other[3], double *omega);
int sats_build_structure(const char *name, int n, const uint8_t *sse_type, const double *centroid,
                         const double *dircos, sats_db **out);
int sats_fit_axis(int sse_type, int n_res, const double *ca_xyz, double dircos[3], double centroid[3]);
int sats_build_structure_from_ca(const char *name, int n, const uint8_t *sse_type, const int32_t *n_res,
                                 const double *ca_xyz, sats_db **out);

/* ---- Gumbel statistics (replaces gumbelstats.h:27-39) ----------------------------------------- */
extern const double sats_gumbel_a;   /* gumbelstats.h:27 */
extern const double sats_gumbel_b;   /* gumbelstats.h:28 */
double sats_norm2(int score, int size1, int size2);        /* gumbelstats.c:91 */
double sats_z_gumbel(int x, double a, double b);           /* gumbelstats.c:50 (int x: truncation!) */
double sats_pv_gumbel(double z);                           /* gumbelstats.c:69 */
/* The smallest raw score whose printed z-score (z_gumbel of the int-truncated norm2, as in cudaSaTabsearch.cu:447-450) reaches
 * z_min for a query of size1 and a structure of size2 SSEs; INT32_MAX if no score can.  sats_search_hits() cuts with it.  */
int32_t sats_score_threshold(double z_min, int size1, int size2);

/* Result block printer = cudaSaTabsearch.cu:415-420 + 442-454 for one (query, pool): three '#'
 * header lines, then "%-8s %d %g %g %g" rows (+ "%3d %3d" map pairs if lsoln).  `index` lists the
 * db entries to print, in print order; scores/maps are indexed by db entry (maps row stride
 * SATS_MAP_STRIDE, may be NULL unless lsoln).  Appends to buf (capacity cap); returns the number
 * of bytes the full text needs (call again with a larger buffer if > cap).                        */
size_t sats_format_block(char *buf, size_t cap, const char *query_id, int query_order,
                         const char *dbfile, int lorder, int lsoln, const sats_db *db,
                         const int32_t *index, int count, const int32_t *scores, const int32_t *maps);

/* ---- reading result files --------------------------------------------------------------
 * Parser for the five-column stdout grammar above ('#' header lines, rows, optional "%3d %3d" map pairs): what the
 * reference's tooling consumes (scripts/tsevalutils.py:223-296, scripts/parsessemap.py:43-147).  A block is one
 * (query, pool) section.                                                                                          */
typedef struct sats_results sats_results;
int sats_results_parse(const char *text, size_t len, sats_results **out);
void sats_results_free(sats_results *r);
int sats_results_blocks(const sats_results *r);
const char *sats_results_query(const sats_results *r, int block);
const char *sats_results_dbfile(const sats_results *r, int block);
int sats_results_flags(const sats_results *r, int block, int flags_tf[3]);
int sats_results_rows(const sats_results *r, int block);
int sats_results_row(const sats_results *r, int block, int row, char name[9], int32_t *score, double *norm2,
                     double *zscore, double *pvalue);
/* SSE map of a row (LSOLN=T): returns the number of pairs and copies up to cap 1-based (query, db) pairs          */
int sats_results_map(const sats_results *r, int block, int row, int32_t *pairs, int cap);
/* ROC AUC of scores against binary labels, ties counted one half (Mann-Whitney U / (P*N)); NaN if a class is empty */
double sats_roc_auc(const double *score, const uint8_t *positive, int n);

/* ---- search (replaces sa_tabsearch_gpu / _noshared / _host, cudaSaTabsearch_kernel.h:15-64,
 *      plus copyQueryToConstantMemory cudaSaTabsearch.cu:486-558 and init_rng :258-264) --------- */

typedef enum sats_rng_mode {
  /* Production: Philox4x32-10 keyed by seed, counter = (draw block, restart, original entry index,
   * query index); exactly `restarts` chains per entry, independent of launch geometry.  One block
   * per two moves (SSE pick + candidate pick from one word, Metropolis draw from the next) and one
   * block of seeding bits per chain: DESIGN.md "stream layout v2".                                */
  SATS_RNG_PHILOX = 0,
  /* Validation: the reference GPU run's 128 x 128 XORWOW streams (curand_init(seed, tid, 0)), block b
   * walking entries b, b+128, ... of the pool in file order, restarts rounded up to a multiple of
   * 128, states persisting from call to call exactly as devStates does.             */
  SATS_RNG_XORWOW_GRID = 1
} sats_rng_mode;

typedef enum sats_accept_mode {
  /* Metropolis thresholds expf((float)delta / T) tabulated on the host with libm, i.e. the exact
   * fp32 values the reference's `-c` path compares against (kernel.cu:1166).                     */
  SATS_ACCEPT_HOST_TABLE = 0,
  /* __expf(__fdividef(delta, T)) evaluated on the device: what the reference's GPU build does under
   * its --use_fast_math (Makefile:51).  For same-box parity with the reference GPU binary, hence
   * only with SATS_RNG_XORWOW_GRID (SATS_ERR_ARG otherwise).                                      */
  SATS_ACCEPT_DEVICE_FAST = 1
} sats_accept_mode;

typedef enum sats_pool {
  SATS_POOL_ALL = 0,     /* every entry (order <= SATS_MAXDIM)                                      */
  SATS_POOL_SMALL = 1,   /* order <= pool_threshold: the reference's shared-memory kernel launch   */
  SATS_POOL_LARGE = 2    /* order  > pool_threshold: the reference's no-shared kernel launch       */
} sats_pool;

typedef struct sats_params {
  int lorder;            /* LORDER: keep sequence order of matched SSEs                            */
  int lsoln;             /* LSOLN: also return the best SSE map                                    */
  int restarts;          /* maxstart (-r)                                                          */
  int rng_mode;          /* sats_rng_mode                                                          */
  int accept_mode;       /* sats_accept_mode                                                       */
  int pool;              /* sats_pool                                                              */
  int pool_threshold;    /* 0 -> SATS_MAXDIM_GPU                                                   */
  int grid_rank;         /* XORWOW_GRID only: this call runs reference blocks b with                */
  int grid_count;        /*   b % grid_count == grid_rank (0/1 -> all 128 blocks)                  */
  int reserved;
  uint64_t seed;         /* 0 -> SATS_REF_SEED                                                     */
} sats_params;

void sats_params_default(sats_params *p);

/* A searcher owns one GPU's copy of (a shard of) the database, laid out for the kernel: entries
 * sorted by decreasing order, one 16-byte-aligned blob each (header, SSE-type bit masks, n x n cells
 * of {fp32 distance, tableau code}), plus streams, the XORWOW state grid and result buffers.     */
typedef struct sats_searcher sats_searcher;

/* shard_count <= 1: whole db.  Otherwise this searcher holds part `shard_rank` of sats_partition(db,
 * shard_count): production (Philox) mode only -- chains are keyed by original entry index, so any split
 * of the entries gives the unsharded results.  XORWOW_GRID searches on such a searcher fail with
 * SATS_ERR_ARG: the validation streams belong to reference BLOCKS walking the whole pool, so a multi-GPU
 * validation run keeps the database replicated (shard_count 1 on every GPU) and splits the blocks with
 * sats_params.grid_rank / grid_count (block b runs where b % grid_count == grid_rank; merge the scores,
 * and take block b's final states from its owner).                                                    */
int sats_searcher_create(const sats_db *db, int device, int shard_rank, int shard_count,
                         sats_searcher **out);
void sats_searcher_free(sats_searcher *s);
int sats_searcher_entry_count(const sats_searcher *s);      /* entries resident on this GPU        */
int sats_searcher_device(const sats_searcher *s);

/* Cost-weighted partition: owner[e] in [0, shard_count) for every db entry.  Greedy
 * longest-processing-time-first: the entries in decreasing order of size are dealt one by one to the
 * shard with the least accumulated cost, cost(order) = 13 + order (the measured per-entry kernel time
 * on a B200 is proportional to it, csrc/sats_internal.h).  Parts are interleaved, not contiguous.    */
int sats_partition(const sats_db *db, int shard_count, int32_t *owner);

/* One-shot search: upload queries[qfirst .. qfirst+qcount), run, copy back.
 * scores: qcount x sats_db_count(db) int32, indexed by ORIGINAL db index; entries outside the pool
 * or outside this searcher's shard are left untouched.  maps: qcount x count x SATS_MAP_STRIDE
 * int32 (-1 = unmapped), required iff lsoln.  query_index_base keys the Philox streams.         */
int sats_search(sats_searcher *s, const sats_db *queries, int qfirst, int qcount,
                const sats_params *params, uint32_t query_index_base, int32_t *scores, int32_t *maps);

/* The same in three steps, for callers that keep data resident (and for measurement):
 *   sats_search_upload   host -> device copy of the queries (async on the searcher's stream)
 *   sats_search_launch   kernels only; results stay on the device.  If elapsed_ms != NULL the launch
 *                        is bracketed by CUDA events on the launching stream and synchronised.
 *   sats_search_collect  device -> host copy of scores (and maps), scatter to original order.   */
int sats_search_upload(sats_searcher *s, const sats_db *queries, int qfirst, int qcount);
int sats_search_launch(sats_searcher *s, const sats_params *params, uint32_t query_index_base,
                       float *elapsed_ms);
int sats_search_collect(sats_searcher *s, int32_t *scores, int32_t *maps);
/* Multi-GPU hosts: sats_search_collect_begin() only ENQUEUES the device -> host copies (pinned staging), so that every
 * GPU's copy is in flight before the first sats_search_collect() waits; collect without begin does both.             */
int sats_search_collect_begin(sats_searcher *s);
/* SURVEY 8(e) "one tiny gather after": the results where they are, for callers that gather the shards with their own
 * collective (ncclGather / AllGather) instead of N host copies.  *d_scores = device pointer to int32 [qcount][entries]
 * in DEVICE order (row = query slot, column = position in this searcher's decreasing-size order; entries outside the
 * launch's pool hold 0x80808080), valid until the next upload / launch on this searcher and produced on its stream
 * (call sats_searcher_sync() first).  slot_query (may be NULL) receives, per row, the query's position in the batch;
 * sats_searcher_entry_index() gives, per column, the ORIGINAL db index.                                             */
int sats_search_device_results(sats_searcher *s, const int32_t **d_scores, int *qcount, int *entries, int32_t *slot_query);
int sats_searcher_entry_index(const sats_searcher *s, int32_t *index);
/* after sats_search_launch(), the k best-scoring entries of every query slot, selected ON THE DEVICE
 * (exact counting select over the integer scores) so that only k (index, score) pairs per query are copied back
 * instead of one score per database entry.  Rows of k: score descending; ties by decreasing structure order, then
 * file order (the device order).  index_out receives ORIGINAL db indices (-1 padding), score_out the raw scores
 * (INT32_MIN padding).  A sharded search returns each shard's local top-k; merge the shards on the host.          */
int sats_search_topk(sats_searcher *s, int k, int32_t *index_out, int32_t *score_out);
/* after sats_search_launch(), the significance cut ON THE DEVICE: every entry whose Gumbel z-score (the
 * fourth output column, cudaSaTabsearch.cu:447-450) is >= z_min -- equivalently whose p-value is <= sats_pv_gumbel(z_min).
 * The host turns the cut into one integer score threshold per (query, structure order) with the same functions the result
 * printer uses (z is a non-decreasing step function of the raw score for fixed sizes), so the selection is exact; the
 * device compares and compacts, and only the hits are copied back.  count_out[q] = number of hits of query slot q (it may
 * exceed cap: the first cap hits are returned); index_out / score_out: rows of cap, hits in device order (decreasing
 * structure order, then file order), ORIGINAL db indices, -1 / INT32_MIN padding.  Sharded search: one call per shard.  */
int sats_search_hits(sats_searcher *s, double z_min, int cap, int32_t *count_out, int32_t *index_out, int32_t *score_out);
/* SURVEY 8(f2), streaming form: bind the cut BEFORE launching and the kernel itself appends every hit -- (query, entry,
 * score) with z-score >= z_min, same exact integer thresholds as sats_search_hits() -- to one device list from its arg-max
 * epilogue (one atomicAdd per hit), query after query, pool after pool, while the search runs.  The dense Q x D score
 * matrix never has to leave the device: sats_search_streamed_hits() copies back the hit count and the hits only
 * (*d2h_bytes, may be NULL, receives the bytes it copied) and returns exactly what sats_search_hits(s, z_min, cap, ...)
 * would (same rows, same device order); if the list overflowed (> 4 M hits) it falls back to that dense post-pass.
 * z_min = NaN unbinds.  The cut stays bound across launches; every launch starts a fresh list.                    */
int sats_search_bind_cut(sats_searcher *s, double z_min);
int sats_search_streamed_hits(sats_searcher *s, int cap, int32_t *count_out, int32_t *index_out, int32_t *score_out,
                              int64_t *d2h_bytes);
int sats_searcher_sync(sats_searcher *s);
/* kernels launched by this searcher since creation (for the bench's gpu_launches claim)          */
long long sats_searcher_launch_count(const sats_searcher *s);
/* copies the XORWOW state grid (16384 x 6 words: d, v0..v4) to the host -- validation aid        */
int sats_searcher_get_xorwow(sats_searcher *s, uint32_t *states6);
/* re-initialise the XORWOW grid as curand_init(seed, tid, 0) (the reference's init_rng)          */
int sats_searcher_reset_xorwow(sats_searcher *s, uint64_t seed);

int sats_device_count(void);
/* Create the device's CUDA context now.  Multi-GPU hosts: call it for every device from ONE thread before creating
 * searchers from several (concurrent context creation is serialised by the driver and measurably slower).       */
int sats_device_init(int device);

/* ---- validation aids: the integer cut-offs the kernel compares raw 32-bit draws with ----------------------------------
 * Every decision the reference takes on a uniform u = curand_uniform-style unit(x) of 32 random bits x is monotone in x,
 * so the library tabulates it once with the reference's own fp32 / fp64 expressions and the kernel never converts a draw
 * (csrc/sats_device.cu).  Exported so that tests can check the tables against an independent restatement.
 *   sats_pick_boundaries  cut[k], k < n: smallest x with (int)((unit(x) - 1.1e-7) * n) >= k   (kernel.cu:67, :1042)
 *   sats_accept_cutoffs   cut[100][230]: a move of score change -d at step m passes iff x < cut[m][d], i.e. iff
 *                         expf((float)(-d) / T_m) > unit(x) (kernel.cu:1166); temps[100] (may be NULL) = T_m in fp32
 *   sats_seed_cutoff      smallest x with unit(x) >= 0.5 (INIT_MATCHPROB, saparams.h:43)                           */
int sats_pick_boundaries(int n, uint32_t *cut);
int sats_accept_cutoffs(uint32_t *cut, float *temps);
uint32_t sats_seed_cutoff(void);

#ifdef __cplusplus
}
#endif
#endif /* SATS_H */
