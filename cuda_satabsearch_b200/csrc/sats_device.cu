// sats_device.cu -- device half of libsats: the searcher (device-resident packed database, query upload,
// kernel dispatch, result gather).  Replaces, for the hot path, the host driver code around the reference's
// kernel launches: db upload cudaSaTabsearch.cu:924-967, copyQueryToConstantMemory :486-558, init_rng :258-264
// and :896-922, the launches :1042/:1223 and the result copies :1075-1087.
#include <algorithm>
#include <array>
#include <map>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <set>
#include <thread>
#include <vector>

#include <time.h>

#include <cuda_runtime.h>
#include <curand_kernel.h>

#include "sats_internal.h"
#include "sats_kparams.h"

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      return sats_fail(SATS_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_), __FILE__, __LINE__, \
                       cudaGetErrorString(e_));                                                          \
  } while (0)

// SATS_TRACE=1: phase timings of searcher creation on stderr (diagnostics only)
static double trace_now()
{
  timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec * 1e3 + t.tv_nsec * 1e-6;
}
static const bool g_trace = getenv("SATS_TRACE") != nullptr;
#define TRACE(what) do { if (g_trace) { double n_ = trace_now(); fprintf(stderr, "[sats dev %d] %-28s %9.2f ms\n", device, what, n_ - t_trace); t_trace = n_; } } while (0)

static const int kScoreSentinel = (int)0x80808080;   // byte-memset pattern marking "not computed by this launch"
static const int kMaxSmem = 227 * 1024;

// the reference's init_rng (cudaSaTabsearch.cu:258-264): curand_init(seed, tid, 0) for the 128 x 128 grid
__global__ void sats_xorwow_init_kernel(uint32_t *states, int n, unsigned long long seed)
{
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n) return;
  curandStateXORWOW_t st;
  curand_init(seed, (unsigned long long)tid, 0ull, &st);
  uint32_t *o = states + (size_t)tid * 6;
  o[0] = st.d;
  for (int k = 0; k < 5; k++) o[1 + k] = st.v[k];
}

// ---- integer forms of the reference's decisions on a uniform --------------------------------------------------------
// The reference turns 32 random bits x into u = unit(x) in (0, 1] (curand_uniform) and then decides with fp32 / fp64
// arithmetic.  unit() is monotone in x, so every such decision is "x < cut" for a cut-off that bisection finds with the
// reference's own expressions; the kernel then compares integers and never converts a draw.
static float unit_from_bits(uint32_t x) { return (float)x * 2.3283064e-10f + 1.1641532e-10f; }     // curand_uniform.h:69-72
// (int)((u - 1.1e-7) * n) in double (cudaSaTabsearch_kernel.cu:67, :1042)
static int pick_from_bits(uint32_t x, int n) { return (int)(((double)unit_from_bits(x) - 1.1e-7) * (double)n); }
// smallest x in [0, 2^32) with pred(x) true, for a predicate that is monotone (false ... false true ... true); 0 if it is
// true everywhere; *none = true (and 0xffffffff returned) if it is false everywhere
template <class Pred> static uint32_t first_true(Pred pred, bool *none = nullptr)
{
  if (none) *none = false;
  if (!pred(0xffffffffu)) { if (none) *none = true; return 0xffffffffu; }
  uint64_t lo = 0, hi = 0xffffffffull;          // pred(hi) holds
  while (lo < hi) {
    const uint64_t mid = lo + (hi - lo) / 2;
    if (pred((uint32_t)mid)) hi = mid;
    else lo = mid + 1;
  }
  return (uint32_t)lo;
}
// pick_cut[k] = smallest x whose SSE pick is >= k, k = 0..n-1 (see pick_index() in sats_kernel.cuh)
extern "C" int sats_pick_boundaries(int n, uint32_t *cut)
{
  if (n < 1 || n > SATS_MAXDIM_EXT || !cut) return sats_fail(SATS_ERR_ARG, "sats_pick_boundaries: bad argument");
  for (int k = 0; k < n; k++) {
    bool none;
    cut[k] = first_true([&](uint32_t x) { return pick_from_bits(x, n) >= k; }, &none);
    // the kernel starts from umulhi(x, n) and steps down by at most one: both must bracket the exact boundary
    const uint64_t naive = (((uint64_t)k << 32) + (uint64_t)n - 1) / (uint64_t)n;       // smallest x with umulhi(x, n) >= k
    const uint64_t next = (((uint64_t)(k + 1) << 32) + (uint64_t)n - 1) / (uint64_t)n;
    if (none || cut[k] < naive || (uint64_t)cut[k] >= next)
      return sats_fail(SATS_ERR_ARG, "sats_pick_boundaries: boundary %d of %d outside its multiply-high step", k, n);
  }
  return SATS_OK;
}
// cut[m][nd], m = 0..99, nd = 0..229: a move of score change -nd at step m is accepted iff x < cut, i.e. iff
// expf((float)(-nd) / T_m) > unit(x) with T_0 = 10, T <- 0.95f T in fp32 (cudaSaTabsearch_kernel.cu:1030, :1166, :1189)
extern "C" int sats_accept_cutoffs(uint32_t *cut, float *temps)
{
  if (!cut) return sats_fail(SATS_ERR_ARG, "sats_accept_cutoffs: null argument");
  float t = 10.0f;
  for (int m = 0; m < SATS_K_MOVES; m++) {
    if (temps) temps[m] = t;
    for (int nd = 0; nd <= SATS_K_DCLAMP; nd++) {
      const float thr = expf((float)(-nd) / t);
      bool none;
      const uint32_t c = first_true([&](uint32_t x) { return !(thr > unit_from_bits(x)); }, &none);
      if (none) return sats_fail(SATS_ERR_ARG, "sats_accept_cutoffs: threshold %g above every uniform", (double)thr);
      cut[(size_t)m * (SATS_K_DCLAMP + 1) + nd] = c;
    }
    t *= 0.95f;
  }
  return SATS_OK;
}
// the seeding pass attempts a match iff unit(x) < 0.5 (INIT_MATCHPROB, saparams.h:43; kernel.cu:624)
extern "C" uint32_t sats_seed_cutoff(void) { return first_true([](uint32_t x) { return !((double)unit_from_bits(x) < 0.5); }); }

typedef sats_kernel_fn kernel_fn;
// one kernel launch of a search, as planned on the host
struct LaunchDesc {
  kernel_fn fn;
  unsigned grid_x, grid_y, threads;
  size_t smem;
  SatsKParams k;
};

struct sats_searcher {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // bucket launches of one search go round-robin over side streams so that the tail of one size bucket
  // overlaps the head of the next (forked from / joined to `stream` with events)
  static const int kSide = 4;
  cudaStream_t side[kSide] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t side_done[kSide] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t fork = nullptr;
  int db_count = 0;                       // entries of the whole database (output indexing)
  int shard_count = 1;                    // > 1: this searcher holds one part of sats_partition()
  std::vector<int32_t> sorted_orig;       // local sorted position -> original db index (decreasing order)
  std::vector<int32_t> sorted_order;
  std::vector<int32_t> file_rank;         // local sorted position -> rank in original file order among local entries
  uint8_t *d_blobs = nullptr;
  uint64_t *d_blob_off = nullptr;
  uint32_t *d_blob_bytes = nullptr;
  uint32_t *d_accept = nullptr;          // Metropolis cut-offs [moves][230], then the fp32 temperature schedule [moves]
  uint32_t accept_cut0 = 0, seed_cut = 0;
  uint32_t *d_xw = nullptr;
  bool xw_ready = false;
  uint64_t xw_seed = 0;
  int32_t *d_pool_list = nullptr;
  int32_t *d_xw_blocks = nullptr;
  int num_sms = 148;
  std::map<std::array<int, 8>, std::array<int, 4>> launch_cfg;
  // the launches of the last production-mode search, captured as a CUDA graph (counter reset, fork to the side streams,
  // one kernel per size bucket, join): an unchanged plan is replayed with a single cudaGraphLaunch
  std::vector<LaunchDesc> graph_plan;
  size_t graph_counters = 0;
  cudaGraphExec_t graph_exec = nullptr;   // launch shape per (kernel variant, shared-memory sizes)
  // everything the plan of a production-mode launch depends on: when the next launch presents the same key, planning is
  // skipped altogether and the instantiated graph is replayed (the planning loop is ~30 us of host time per search, which
  // shows once a GPU holds a 1/8 shard and a search lasts 1.4 ms)
  struct LaunchKey {
    sats_params p; uint32_t qbase; int q; int d; uint64_t qsig; const void *ptr[10];
  } graph_key;
  bool graph_key_valid = false;
  uint64_t qsig = 0;                      // hash of the uploaded queries' orders by slot (the plan depends on sizes only)
  int *d_counters = nullptr; size_t counter_cap = 0;   // one work counter per (bucket launch, query) of a search
  // queries
  std::array<std::array<uint32_t, SATS_MAXDIM + 1>, SATS_MAXDIM + 1> pick_cut;      // [n1]: SSE-pick boundaries, built on first use
  std::array<bool, SATS_MAXDIM + 1> pick_cut_ready{};
  uint8_t *d_qblobs = nullptr; size_t qblob_cap = 0;
  uint64_t *d_qoff = nullptr; uint32_t *d_qbytes = nullptr; int qmeta_cap = 0;
  uint8_t *h_qstage = nullptr; size_t h_qstage_cap = 0;
  // Queries live on the device in SLOTS: the batch sorted (stably) by size class, so that a launch -- which sizes its
  // shared memory for the largest query it covers -- spans queries of similar size.  slot_q[slot] = position in the batch.
  std::vector<int> q_n1;               // by slot
  std::vector<uint32_t> q_bytes;       // by slot
  std::vector<int> slot_q;
  // results
  int32_t *d_scores = nullptr; int8_t *d_maps = nullptr;
  int32_t *h_scores = nullptr; int8_t *h_maps = nullptr;
  size_t score_cap = 0, map_cap = 0;
  int last_q = 0, last_lsoln = 0;
  bool collect_pending = false;           // sats_search_collect_begin() has enqueued the copies of the last launch
  int32_t *d_topk = nullptr, *h_topk = nullptr; size_t topk_cap = 0;
  int32_t *d_sorted_order = nullptr;      // order of every resident entry, device order (for sats_search_hits)
  int32_t *d_hits = nullptr, *h_hits = nullptr; size_t hits_cap = 0;
  // streaming hits: a bound significance cut makes every launch append its hits to d_stream_list (see SatsKParams)
  bool cut_bound = false; double cut_z = 0.0;
  int32_t *d_stream_thr = nullptr; size_t stream_thr_cap = 0;      // [query slot][SATS_MAXDIM_EXT + 1]
  unsigned *d_stream_cursor = nullptr;
  int2 *d_stream_list = nullptr; size_t stream_list_cap = 0;
  int2 *h_stream = nullptr; size_t h_stream_cap = 0;
  bool stream_valid = false;              // the last launch ran with the cut bound
  long long launches = 0;
  std::set<kernel_fn> smem_opted;         // kernel variants already opted into 227 KB of dynamic shared memory
};


static kernel_fn pick_kernel(int w1, int w2, bool lorder, bool xorwow, bool lsoln)
{
  switch (w1) {
    case 1: return sats_pick_kernel_w1(w2, lorder, xorwow, lsoln);
    case 2: return sats_pick_kernel_w2(w2, lorder, xorwow, lsoln);
    default: return sats_pick_kernel_w4(w2, lorder, xorwow, lsoln);
  }
}
static int words_for(int n) { return n <= 32 ? 1 : (n <= 64 ? 2 : 4); }
// mask words of a QUERY: one bit more than its order, for the sentinel "SSE n1" the kernel keeps above the last one
static int qwords_for(int n) { return n < 32 ? 1 : (n < 64 ? 2 : 4); }
// query size classes (upper bounds of the order): one launch never mixes classes
static int query_class(int n)
{
  static const int bounds[] = {8, 12, 16, 20, 24, 31, 48, 63};      // 31 / 63: the last orders whose masks fit 1 / 2 words
  int c = 0;
  while (c < 8 && n > bounds[c]) c++;
  return c;
}
static size_t round16(size_t x) { return (x + 15) & ~(size_t)15; }
static size_t entry_blob_bytes(int n) { return round16(SATS_K_ENTRY_HDR + 8 * (size_t)n * (n + 1)); }     // header, NaN row, n x n cells
static size_t query_blob_bytes(int n) { return round16(SATS_K_QUERY_HDR + 8 * (size_t)n * n); }

static void fill_cells(const sats_db *db, int e, uint8_t *cells)
{
  int n = db->order[e];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      float d = db->dist(e, i, j);
      // (first letter << 4) | second letter, letters 0..4: the XOR of two codes indexes the kernel's zeta table
      // (gated() in sats_kernel.cuh); the diagonal holds the SSE type and is never scored
      const uint32_t raw = db->code(e, i, j);
      uint32_t code = (std::min(raw >> 4, 4u) << 4) | std::min(raw & 15u, 4u);
      memcpy(cells + 8 * ((size_t)i * n + j), &d, 4);
      memcpy(cells + 8 * ((size_t)i * n + j) + 4, &code, 4);
    }
}

extern "C" int sats_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// Creates the CUDA context of a device (cudaSetDevice + a no-op runtime call).  Context creation is serialised inside the
// driver and SLOWER when several host threads do it at once (measured: 0.75 s + 0.23 s one after the other, 1.9 s from two
// threads, profiles/r02_experiments.txt), so a multi-GPU host calls this for every device first, from one thread, and
// only then builds its searchers side by side.
extern "C" int sats_device_init(int device)
{
  int ndev = sats_device_count();
  if (device < 0 || device >= ndev) return sats_fail(SATS_ERR_ARG, "device %d out of range (have %d)", device, ndev);
  CK(cudaSetDevice(device));
  CK(cudaFree(nullptr));
  return SATS_OK;
}

extern "C" int sats_searcher_create(const sats_db *db, int device, int shard_rank, int shard_count, sats_searcher **out)
try {
  if (!db || !out) return sats_fail(SATS_ERR_ARG, "sats_searcher_create: null argument");
  if (shard_count < 1) shard_count = 1;
  if (shard_rank < 0 || shard_rank >= shard_count) return sats_fail(SATS_ERR_ARG, "bad shard %d of %d", shard_rank, shard_count);
  double t_trace = trace_now();
  int ndev = sats_device_count();
  if (ndev < 1) return sats_fail(SATS_ERR_CUDA, "no CUDA device available (this library has no CPU search path)");
  if (device < 0 || device >= ndev) return sats_fail(SATS_ERR_ARG, "device %d out of range (have %d)", device, ndev);
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return sats_fail(SATS_ERR_CUDA, "device %d is sm_%d%d; this build targets sm_100a", device, prop.major, prop.minor);

  TRACE("set device + properties");
  sats_searcher *s = new sats_searcher();
  s->device = device;
  s->num_sms = prop.multiProcessorCount;
  s->db_count = db->count();
  s->shard_count = shard_count;
  std::vector<int32_t> owner((size_t)db->count(), 0);
  if (shard_count > 1) sats_partition(db, shard_count, owner.data());
  std::vector<int32_t> local;
  for (int e = 0; e < db->count(); e++) if (owner[e] == shard_rank) local.push_back(e);
  std::vector<int32_t> pos(local.size());
  std::iota(pos.begin(), pos.end(), 0);
  std::stable_sort(pos.begin(), pos.end(), [&](int a, int b) { return db->order[local[a]] > db->order[local[b]]; });
  size_t total = 0;
  std::vector<uint64_t> off(local.size());
  std::vector<uint32_t> bytes(local.size());
  for (size_t k = 0; k < local.size(); k++) {
    int e = local[pos[k]];
    s->sorted_orig.push_back(e);
    s->sorted_order.push_back(db->order[e]);
    s->file_rank.push_back(pos[k]);
    off[k] = total;
    bytes[k] = (uint32_t)entry_blob_bytes(db->order[e]);
    total += bytes[k];
  }
  TRACE("partition + sort");
  std::vector<uint8_t> blobs(total ? total : 16, 0);
  // the blobs are independent: fill them on a few host threads (100 k structures = 240 MB of cells)
  auto fill_range = [&](size_t k0, size_t k1) {
  for (size_t k = k0; k < k1; k++) {
    int e = s->sorted_orig[k], n = db->order[e];
    uint8_t *b = blobs.data() + off[k];
    int32_t hdr[4] = {n, e, 0, 0};
    memcpy(b, hdr, 16);
    uint32_t tm[4][4] = {{0}};
    for (int j = 0; j < n; j++) tm[db->code(e, j, j) & 3][j >> 5] |= 1u << (j & 31);
    memcpy(b + 16, tm, 64);
    // "row -1": n cells of NaN distance right in front of the matrix, so that the missing side of a move (partner -1)
    // addresses a row whose gate never opens without any special casing in the kernel
    for (int j = 0; j < n; j++) { const uint32_t nan_cell[2] = {0x7fc00000u, 0u}; memcpy(b + SATS_K_ENTRY_HDR + 8 * (size_t)j, nan_cell, 8); }
    fill_cells(db, e, b + SATS_K_ENTRY_HDR + 8 * (size_t)n);
  }
  };
  {
    const size_t nthreads = total < (8u << 20) ? 1 : std::max<size_t>(1, std::min<size_t>(8, std::thread::hardware_concurrency()));
    struct Joiner {                                  // joins whatever was started, also when starting a thread throws
      std::vector<std::thread> pool;
      ~Joiner() { for (auto &th : pool) if (th.joinable()) th.join(); }
    } workers;
    size_t k0 = 0;
    for (size_t t = 0; t < nthreads; t++) {          // equal BYTES per thread (the list is sorted by decreasing size)
      size_t k1 = k0;
      const uint64_t upto = total / nthreads * (t + 1);
      while (k1 < local.size() && (t + 1 == nthreads || off[k1] < upto)) k1++;
      if (t + 1 == nthreads) fill_range(k0, local.size());
      else workers.pool.emplace_back(fill_range, k0, k1);
      k0 = k1;
    }
  }
  TRACE("blob fill");
  auto fail = [&](int rc) { sats_searcher_free(s); return rc; };
#define CKF(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(sats_fail(SATS_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_))); } while (0)
  CKF(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  CKF(cudaEventCreate(&s->ev0));
  CKF(cudaEventCreate(&s->ev1));
  CKF(cudaEventCreateWithFlags(&s->fork, cudaEventDisableTiming));
  for (int i = 0; i < sats_searcher::kSide; i++) {
    CKF(cudaStreamCreateWithFlags(&s->side[i], cudaStreamNonBlocking));
    CKF(cudaEventCreateWithFlags(&s->side_done[i], cudaEventDisableTiming));
  }
  TRACE("streams + events");
  CKF(cudaMalloc(&s->d_blobs, blobs.size()));
  CKF(cudaMalloc(&s->d_blob_off, std::max<size_t>(1, local.size()) * 8));
  CKF(cudaMalloc(&s->d_blob_bytes, std::max<size_t>(1, local.size()) * 4));
  CKF(cudaMalloc(&s->d_pool_list, std::max<size_t>(1, local.size()) * 4));
  CKF(cudaMalloc(&s->d_xw_blocks, SATS_REF_GRID_BLOCKS * 4));
  TRACE("cudaMalloc x5");
  CKF(cudaMemcpy(s->d_blobs, blobs.data(), blobs.size(), cudaMemcpyHostToDevice));
  TRACE("blob upload");
  CKF(cudaMalloc(&s->d_sorted_order, std::max<size_t>(1, local.size()) * 4));
  if (!local.empty()) {
    CKF(cudaMemcpy(s->d_blob_off, off.data(), local.size() * 8, cudaMemcpyHostToDevice));
    CKF(cudaMemcpy(s->d_blob_bytes, bytes.data(), local.size() * 4, cudaMemcpyHostToDevice));
    CKF(cudaMemcpy(s->d_sorted_order, s->sorted_order.data(), local.size() * 4, cudaMemcpyHostToDevice));
  }
  // Metropolis cut-offs and temperatures exactly as the reference's host path evaluates them
  std::vector<float> temps(SATS_K_MOVES);
  std::vector<uint32_t> tab((size_t)SATS_K_MOVES * (SATS_K_DCLAMP + 1));
  if (sats_accept_cutoffs(tab.data(), temps.data()) != SATS_OK) return fail(SATS_ERR_ARG);
  s->accept_cut0 = tab[0];                                                // d == 0: expf(0) = 1 at every step
  for (int m = 0; m < SATS_K_MOVES; m++)
    if (tab[(size_t)m * (SATS_K_DCLAMP + 1)] != s->accept_cut0) return fail(sats_fail(SATS_ERR_ARG, "accept cut-off for d = 0 varies"));
  s->seed_cut = sats_seed_cutoff();
  CKF(cudaMalloc(&s->d_accept, (tab.size() + temps.size()) * 4));      // cut-offs, then the temperature schedule
  CKF(cudaMemcpy(s->d_accept, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
  CKF(cudaMemcpy(s->d_accept + tab.size(), temps.data(), temps.size() * 4, cudaMemcpyHostToDevice));
  CKF(cudaMalloc(&s->d_xw, (size_t)SATS_REF_GRID_BLOCKS * SATS_REF_GRID_THREADS * 6 * 4));
  TRACE("tables");
#undef CKF
  *out = s;
  return SATS_OK;
}
SATS_CATCH_ALL

extern "C" void sats_searcher_free(sats_searcher *s)
{
  if (!s) return;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  cudaFree(s->d_blobs); cudaFree(s->d_blob_off); cudaFree(s->d_blob_bytes); cudaFree(s->d_accept); cudaFree(s->d_xw);
  cudaFree(s->d_pool_list); cudaFree(s->d_xw_blocks); cudaFree(s->d_counters); cudaFree(s->d_qblobs); cudaFree(s->d_qoff); cudaFree(s->d_qbytes);
  cudaFree(s->d_scores); cudaFree(s->d_maps); cudaFree(s->d_topk); cudaFreeHost(s->h_topk);
  cudaFree(s->d_sorted_order); cudaFree(s->d_hits); cudaFreeHost(s->h_hits);
  cudaFree(s->d_stream_thr); cudaFree(s->d_stream_cursor); cudaFree(s->d_stream_list); cudaFreeHost(s->h_stream);
  if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
  cudaFreeHost(s->h_qstage); cudaFreeHost(s->h_scores); cudaFreeHost(s->h_maps);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  if (s->fork) cudaEventDestroy(s->fork);
  for (int i = 0; i < sats_searcher::kSide; i++) {
    if (s->side_done[i]) cudaEventDestroy(s->side_done[i]);
    if (s->side[i]) cudaStreamDestroy(s->side[i]);
  }
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

extern "C" int sats_searcher_entry_count(const sats_searcher *s) { return s ? (int)s->sorted_orig.size() : 0; }
extern "C" int sats_searcher_device(const sats_searcher *s) { return s ? s->device : -1; }
extern "C" long long sats_searcher_launch_count(const sats_searcher *s) { return s ? s->launches : 0; }

extern "C" int sats_searcher_sync(sats_searcher *s)
{
  if (!s) return sats_fail(SATS_ERR_ARG, "null searcher");
  CK(cudaSetDevice(s->device));
  CK(cudaStreamSynchronize(s->stream));
  return SATS_OK;
}

extern "C" int sats_searcher_reset_xorwow(sats_searcher *s, uint64_t seed)
{
  if (!s) return sats_fail(SATS_ERR_ARG, "null searcher");
  CK(cudaSetDevice(s->device));
  const int n = SATS_REF_GRID_BLOCKS * SATS_REF_GRID_THREADS;
  sats_xorwow_init_kernel<<<SATS_REF_GRID_BLOCKS, SATS_REF_GRID_THREADS, 0, s->stream>>>(s->d_xw, n, seed);
  CK(cudaGetLastError());
  s->launches++;
  s->xw_ready = true;
  s->xw_seed = seed;
  return SATS_OK;
}

extern "C" int sats_searcher_get_xorwow(sats_searcher *s, uint32_t *states6)
{
  if (!s || !states6) return sats_fail(SATS_ERR_ARG, "null argument");
  CK(cudaSetDevice(s->device));
  CK(cudaStreamSynchronize(s->stream));
  CK(cudaMemcpy(states6, s->d_xw, (size_t)SATS_REF_GRID_BLOCKS * SATS_REF_GRID_THREADS * 24, cudaMemcpyDeviceToHost));
  return SATS_OK;
}

extern "C" int sats_search_upload(sats_searcher *s, const sats_db *queries, int qfirst, int qcount)
try {
  if (!s || !queries || qcount < 1 || qfirst < 0 || qfirst + qcount > queries->count())
    return sats_fail(SATS_ERR_ARG, "sats_search_upload: bad query range");
  CK(cudaSetDevice(s->device));
  CK(cudaStreamSynchronize(s->stream));   // staging buffers may still be in flight
  size_t total = 0;
  std::vector<uint64_t> off((size_t)qcount);
  for (int q = 0; q < qcount; q++) {
    int n = queries->order[qfirst + q];
    if (n > SATS_MAXDIM) return sats_fail(SATS_ERR_ARG, "query %s has order %d; queries are limited to %d SSEs", queries->name(qfirst + q), n, SATS_MAXDIM);
  }
  s->slot_q.resize((size_t)qcount);
  std::iota(s->slot_q.begin(), s->slot_q.end(), 0);
  std::stable_sort(s->slot_q.begin(), s->slot_q.end(), [&](int a, int b) {
    return query_class(queries->order[qfirst + a]) < query_class(queries->order[qfirst + b]);
  });
  s->q_n1.assign((size_t)qcount, 0);
  s->q_bytes.assign((size_t)qcount, 0);
  for (int slot = 0; slot < qcount; slot++) {
    int n = queries->order[qfirst + s->slot_q[slot]];
    s->q_n1[slot] = n;
    off[slot] = total;
    s->q_bytes[slot] = (uint32_t)query_blob_bytes(n);
    total += s->q_bytes[slot];
  }
  s->qsig = 1469598103934665603ull;                       // FNV-1a over the slot order and the orders
  for (int slot = 0; slot < qcount; slot++) {
    s->qsig = (s->qsig ^ (uint64_t)(uint32_t)s->q_n1[slot]) * 1099511628211ull;
    s->qsig = (s->qsig ^ (uint64_t)(uint32_t)s->slot_q[slot]) * 1099511628211ull;
  }
  size_t meta = (size_t)qcount * 12;
  if (total + meta > s->h_qstage_cap) {
    cudaFreeHost(s->h_qstage);
    s->h_qstage = nullptr;
    s->h_qstage_cap = 0;
    CK(cudaMallocHost(&s->h_qstage, total + meta));
    s->h_qstage_cap = total + meta;
  }
  if (total > s->qblob_cap) {
    cudaFree(s->d_qblobs);
    s->d_qblobs = nullptr;
    s->qblob_cap = 0;
    CK(cudaMalloc(&s->d_qblobs, total));
    s->qblob_cap = total;
  }
  if (qcount > s->qmeta_cap) {
    cudaFree(s->d_qoff); cudaFree(s->d_qbytes);
    s->d_qoff = nullptr; s->d_qbytes = nullptr; s->qmeta_cap = 0;
    CK(cudaMalloc(&s->d_qoff, (size_t)qcount * 8));
    CK(cudaMalloc(&s->d_qbytes, (size_t)qcount * 4));
    s->qmeta_cap = qcount;
  }
  memset(s->h_qstage, 0, total);
  for (int slot = 0; slot < qcount; slot++) {
    int e = qfirst + s->slot_q[slot], n = queries->order[e];
    uint8_t *b = s->h_qstage + off[slot];
    int32_t hdr[4] = {n, s->slot_q[slot], 0, 0};     // hdr[1]: position in the batch; + q_index_base = Philox query index
    memcpy(b, hdr, 16);
    for (int i = 0; i < n; i++) b[16 + i] = queries->code(e, i, i) & 3;      // SSE type 0..3 (the db constructors reject anything else)
    if (!s->pick_cut_ready[n]) {
      int prc = sats_pick_boundaries(n, s->pick_cut[n].data());
      if (prc) return prc;
      s->pick_cut_ready[n] = true;
    }
    memcpy(b + SATS_K_QUERY_PICK, s->pick_cut[n].data(), (size_t)n * 4);
    fill_cells(queries, e, b + SATS_K_QUERY_HDR);
  }
  memcpy(s->h_qstage + total, off.data(), (size_t)qcount * 8);
  memcpy(s->h_qstage + total + (size_t)qcount * 8, s->q_bytes.data(), (size_t)qcount * 4);
  CK(cudaMemcpyAsync(s->d_qblobs, s->h_qstage, total, cudaMemcpyHostToDevice, s->stream));
  CK(cudaMemcpyAsync(s->d_qoff, s->h_qstage + total, (size_t)qcount * 8, cudaMemcpyHostToDevice, s->stream));
  CK(cudaMemcpyAsync(s->d_qbytes, s->h_qstage + total + (size_t)qcount * 8, (size_t)qcount * 4, cudaMemcpyHostToDevice, s->stream));
  s->last_q = qcount;
  return SATS_OK;
}
SATS_CATCH_ALL

static int ensure_results(sats_searcher *s, int qcount, int lsoln)
{
  size_t n = (size_t)qcount * std::max<size_t>(1, s->sorted_orig.size());
  if (n > s->score_cap) {
    cudaFree(s->d_scores); cudaFreeHost(s->h_scores);
    s->d_scores = nullptr; s->h_scores = nullptr; s->score_cap = 0;
    CK(cudaMalloc(&s->d_scores, n * 4));
    CK(cudaMallocHost(&s->h_scores, n * 4));
    s->score_cap = n;
  }
  if (lsoln && n > s->map_cap) {
    cudaFree(s->d_maps); cudaFreeHost(s->h_maps);
    s->d_maps = nullptr; s->h_maps = nullptr; s->map_cap = 0;
    CK(cudaMalloc(&s->d_maps, n * SATS_K_MAPROW));
    CK(cudaMallocHost(&s->h_maps, n * SATS_K_MAPROW));
    s->map_cap = n;
  }
  return SATS_OK;
}

// Opt the kernel variants this searcher actually launches into the full dynamic shared memory, once each (setting the
// attribute on all 72 variants up front made the driver load every one of them: ~10 ms on a process's first search).
static int allow_full_smem(sats_searcher *s, kernel_fn fn)
{
  if (s->smem_opted.count(fn)) return SATS_OK;
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
  s->smem_opted.insert(fn);
  return SATS_OK;
}

// upper bounds (entry order) of the launch buckets: each bucket is one launch with its own shared-memory sizing
static const int kBucketBounds[] = {8, 12, 16, 20, 24, 32, 48, 64, 96, SATS_MAXDIM, SATS_MAXDIM_EXT};

extern "C" int sats_search_launch(sats_searcher *s, const sats_params *pp, uint32_t query_index_base, float *elapsed_ms)
try {
  if (!s || !pp) return sats_fail(SATS_ERR_ARG, "sats_search_launch: null argument");
  if (s->last_q < 1) return sats_fail(SATS_ERR_ARG, "sats_search_launch: no queries uploaded");
  if (pp->restarts < 1) return sats_fail(SATS_ERR_ARG, "restarts must be >= 1");
  if (pp->accept_mode == SATS_ACCEPT_DEVICE_FAST && pp->rng_mode != SATS_RNG_XORWOW_GRID)
    return sats_fail(SATS_ERR_ARG, "ACCEPT_DEVICE_FAST exists for same-box parity with the reference GPU binary: use it with XORWOW_GRID");
  CK(cudaSetDevice(s->device));
  int rc = SATS_OK;
  const int D = (int)s->sorted_orig.size();
  const int Q = s->last_q;
  const bool xorwow = pp->rng_mode == SATS_RNG_XORWOW_GRID;
  const int thr = pp->pool_threshold > 0 ? pp->pool_threshold : SATS_MAXDIM_GPU;
  const uint64_t seed = pp->seed ? pp->seed : SATS_REF_SEED;
  rc = ensure_results(s, Q, pp->lsoln);
  if (rc) return rc;
  s->last_lsoln = pp->lsoln;
  s->collect_pending = false;
  s->stream_valid = false;
  const int32_t *hit_thr = nullptr;
  if (s->cut_bound) {
    if (Q > 65536) return sats_fail(SATS_ERR_ARG, "streaming hits: at most 65536 queries per batch (a hit record keeps the query slot in 16 bits)");
    // one integer score threshold per (query slot, structure order), as in sats_search_hits(); list room for every
    // (query, entry) pair up to 4 M hits (32 MB) -- beyond that the reader falls back to the dense post-pass
    const int T = SATS_MAXDIM_EXT + 1;
    std::vector<int32_t> thr((size_t)Q * T);
    for (int q = 0; q < Q; q++)
      for (int n2 = 0; n2 < T; n2++) thr[(size_t)q * T + n2] = n2 == 0 ? INT_MAX : sats_score_threshold(s->cut_z, s->q_n1[q], n2);
    if (thr.size() > s->stream_thr_cap) {
      cudaFree(s->d_stream_thr); s->d_stream_thr = nullptr; s->stream_thr_cap = 0;
      CK(cudaMalloc(&s->d_stream_thr, thr.size() * 4));
      s->stream_thr_cap = thr.size();
    }
    const size_t want = std::min<size_t>((size_t)Q * std::max(1, D), (size_t)4 << 20);
    if (want > s->stream_list_cap) {
      cudaFree(s->d_stream_list); s->d_stream_list = nullptr; s->stream_list_cap = 0;
      CK(cudaMalloc(&s->d_stream_list, want * sizeof(int2)));
      s->stream_list_cap = want;
    }
    if (!s->d_stream_cursor) CK(cudaMalloc(&s->d_stream_cursor, sizeof(unsigned)));
    CK(cudaStreamSynchronize(s->stream));           // `thr` is pageable host memory
    CK(cudaMemcpyAsync(s->d_stream_thr, thr.data(), thr.size() * 4, cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemsetAsync(s->d_stream_cursor, 0, sizeof(unsigned), s->stream));
    CK(cudaStreamSynchronize(s->stream));
    hit_thr = s->d_stream_thr;
    s->stream_valid = true;
  }
  CK(cudaMemsetAsync(s->d_scores, 0x80, (size_t)Q * std::max(1, D) * 4, s->stream));
  if (elapsed_ms) CK(cudaEventRecord(s->ev0, s->stream));

  sats_searcher::LaunchKey key;
  memset(&key, 0, sizeof key);
  memcpy(&key.p, pp, sizeof key.p);       // (struct copy would leave padding bytes undefined)
  key.qbase = query_index_base; key.q = Q; key.d = D; key.qsig = s->qsig;
  const void *key_ptrs[10] = {s->d_qblobs, s->d_qoff, s->d_qbytes, s->d_scores, pp->lsoln ? (const void *)s->d_maps : nullptr, hit_thr,
                              s->d_stream_cursor, s->d_stream_list, s->d_counters, (const void *)(uintptr_t)s->stream_list_cap};
  memcpy(key.ptr, key_ptrs, sizeof key.ptr);
  static const bool no_graph_fast = getenv("SATS_NO_GRAPH") != nullptr || getenv("SATS_TW") != nullptr || getenv("SATS_TEAMS") != nullptr;
  if (!xorwow && !no_graph_fast && s->graph_exec && s->graph_key_valid && memcmp(&key, &s->graph_key, sizeof key) == 0) {
    CK(cudaGraphLaunch(s->graph_exec, s->stream));
    s->launches += (long long)s->graph_plan.size();
    if (elapsed_ms) {
      CK(cudaEventRecord(s->ev1, s->stream));
      CK(cudaEventSynchronize(s->ev1));
      CK(cudaEventElapsedTime(elapsed_ms, s->ev0, s->ev1));
    }
    return SATS_OK;
  }
  s->graph_key_valid = false;

  // pool -> contiguous range of the (decreasing-order) sorted list
  int first_small = 0;
  while (first_small < D && s->sorted_order[first_small] > thr) first_small++;
  int r0 = 0, r1 = D;
  if (pp->pool == SATS_POOL_SMALL) r0 = first_small;
  else if (pp->pool == SATS_POOL_LARGE) r1 = first_small;

  SatsKParams k;
  memset(&k, 0, sizeof k);
  k.blobs = s->d_blobs; k.blob_off = s->d_blob_off; k.blob_bytes = s->d_blob_bytes;
  k.qblobs = s->d_qblobs; k.qblob_off = s->d_qoff; k.qblob_bytes = s->d_qbytes;
  k.restarts = pp->restarts; k.lsoln = pp->lsoln; k.accept_mode = pp->accept_mode;
  for (uint32_t r = 0; r < 10; r++) {
    k.rk[2 * r] = (uint32_t)seed + r * 0x9E3779B9u;
    k.rk[2 * r + 1] = (uint32_t)(seed >> 32) + r * 0xBB67AE85u;
  }
  k.accept_cut = s->d_accept;
  k.accept_cut0 = s->accept_cut0;
  k.seed_cut = s->seed_cut;
  k.q_index_base = query_index_base;
  k.temps = reinterpret_cast<const float *>(s->d_accept + (size_t)SATS_K_MOVES * (SATS_K_DCLAMP + 1));
  k.out_scores = s->d_scores; k.out_maps = pp->lsoln ? s->d_maps : nullptr; k.out_stride = std::max(1, D);
  k.xw_states = s->d_xw; k.pool_list = s->d_pool_list; k.xw_blocks = s->d_xw_blocks;
  k.hit_thr = hit_thr; k.hit_cursor = s->d_stream_cursor; k.hit_list = s->d_stream_list; k.hit_cap = (unsigned)s->stream_list_cap;

  if (r1 > r0) {
    if (xorwow) {
      // The validation streams belong to reference blocks, and block b walks pool positions b, b + 128, ... of the WHOLE
      // pool (cudaSaTabsearch_kernel.cu:932): a searcher that holds only a part of the database cannot reproduce that.
      // Multi-GPU validation runs keep the database replicated and split the blocks (grid_rank / grid_count).
      if (s->shard_count > 1)
        return sats_fail(SATS_ERR_ARG, "XORWOW_GRID needs the whole database on the device: create the searcher with shard_count 1 "
                                       "and split the reference blocks with sats_params.grid_rank / grid_count");
      if (!s->xw_ready || s->xw_seed != seed) { rc = sats_searcher_reset_xorwow(s, seed); if (rc) return rc; }
      // pool positions in original file order (cudaSaTabsearch_kernel.cu:932 walks d_orders[] as loaded)
      std::vector<int32_t> list;
      for (int kpos = r0; kpos < r1; kpos++) list.push_back(kpos);
      std::sort(list.begin(), list.end(), [&](int a, int b) { return s->file_rank[a] < s->file_rank[b]; });
      int grid_count = pp->grid_count > 1 ? pp->grid_count : 1;
      int grid_rank = grid_count > 1 ? pp->grid_rank : 0;
      if (grid_rank < 0 || grid_rank >= grid_count) return sats_fail(SATS_ERR_ARG, "bad grid_rank %d of %d", grid_rank, grid_count);
      std::vector<int32_t> blocks;
      for (int b = 0; b < SATS_REF_GRID_BLOCKS; b++) if (b % grid_count == grid_rank) blocks.push_back(b);
      CK(cudaStreamSynchronize(s->stream));
      CK(cudaMemcpy(s->d_pool_list, list.data(), list.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(s->d_xw_blocks, blocks.data(), blocks.size() * 4, cudaMemcpyHostToDevice));
      const int n2max = s->sorted_order[r0];
      k.pool_count = (int)list.size();
      k.tw = SATS_REF_GRID_THREADS; k.teams = 1;
      k.sm_entry_bytes = (int)entry_blob_bytes(n2max);
      std::vector<int> slot_of((size_t)Q);
      for (int slot = 0; slot < Q; slot++) slot_of[s->slot_q[slot]] = slot;
      for (int qo = 0; qo < Q; qo++) {    // one launch per query, in the batch's order: the streams carry over
        const int q = slot_of[qo];
        const int n1 = s->q_n1[q];
        k.q_first = q;
        k.sm_query_bytes = qwords_for(n1) > 2 ? SATS_K_QUERY_HDR : (int)s->q_bytes[q];
        k.sm_mapwords = qwords_for(n1) > 2 ? (n1 + 3) / 4 : n1 + 2;      // word maps carry elements -1 and n1
        k.sm_bmapwords = pp->lsoln ? (n1 + 3) / 4 : 0;
        k.sm_qmask_bytes = (int)round16((size_t)n1 * words_for(n2max) * 4);
        k.sm_team_bytes = (int)round16(k.sm_entry_bytes + (size_t)(k.sm_mapwords + k.sm_bmapwords) * k.tw * 4 + SATS_K_SCRATCH_BYTES + (size_t)(k.tw / 32) * k.sm_qmask_bytes);
        size_t smem = SATS_K_BAR_BYTES + k.sm_query_bytes + SATS_K_ZTAB_BYTES + k.sm_team_bytes;
        if (smem > (size_t)kMaxSmem) return sats_fail(SATS_ERR_ARG, "query %d x entry order %d needs %zu B of shared memory", q, n2max, smem);
        kernel_fn fn = pick_kernel(qwords_for(n1), words_for(n2max), pp->lorder != 0, true, pp->lsoln != 0);
        rc = allow_full_smem(s, fn);
        if (rc) return rc;
        fn<<<dim3((unsigned)blocks.size(), 1), k.tw, smem, s->stream>>>(k);
        CK(cudaGetLastError());
        s->launches++;
      }
    } else {
      const size_t ncounters = (size_t)Q * (sizeof kBucketBounds / sizeof kBucketBounds[0]);
      if (ncounters > s->counter_cap) {
        cudaFree(s->d_counters); s->d_counters = nullptr; s->counter_cap = 0;
        CK(cudaMalloc(&s->d_counters, ncounters * sizeof(int)));
        s->counter_cap = ncounters;
      }
      size_t counter_base = 0;
      std::vector<LaunchDesc> plan;
      int tw_max = std::min(128, ((pp->restarts + 31) / 32) * 32);
      if (const char *e = getenv("SATS_TW")) { int t = atoi(e); if (t == 32 || t == 64 || t == 128) tw_max = std::min(tw_max, t); }
      int teams_cap = SATS_K_MAXTHREADS / 32;
      if (const char *e = getenv("SATS_TEAMS")) teams_cap = std::max(1, std::min(teams_cap, atoi(e)));
      // runs of consecutive queries with the same mask width share launches (grid.y)
      for (int q0 = 0; q0 < Q;) {
        int q1 = q0 + 1;
        const int w1 = qwords_for(s->q_n1[q0]);
        while (q1 < Q && query_class(s->q_n1[q1]) == query_class(s->q_n1[q0]) && q1 - q0 < 65535) q1++;   // gridDim.y limit
        int n1max = 0; uint32_t qbmax = 0;
        for (int q = q0; q < q1; q++) { n1max = std::max(n1max, s->q_n1[q]); qbmax = std::max(qbmax, s->q_bytes[q]); }
        k.q_first = q0;
        k.sm_query_bytes = w1 > 2 ? SATS_K_QUERY_HDR : (int)qbmax;
        k.sm_mapwords = w1 > 2 ? (n1max + 3) / 4 : n1max + 2;      // word maps carry elements -1 and n1
        k.sm_bmapwords = pp->lsoln ? (n1max + 3) / 4 : 0;
        // launch shape for a bucket whose largest entry has order n2max: team width (threads sharing one entry; results do
        // not depend on it) and teams per CTA -- whatever keeps the most warps resident per SM, a narrower team only when
        // it buys strictly more.  The choice depends only on the kernel variant and the shared-memory sizes: remember it (the
        // occupancy queries would otherwise cost a few hundred microseconds of host time per search).
        struct Shape { kernel_fn fn; int tw, teams, team_bytes, ctas, entry_bytes, qmask_bytes; };
        auto shape_for = [&](int n2max, Shape *out) -> int {
          Shape sh;
          sh.entry_bytes = (int)entry_blob_bytes(n2max);
          sh.qmask_bytes = (int)round16((size_t)n1max * words_for(n2max) * 4);
          sh.fn = pick_kernel(w1, words_for(n2max), pp->lorder != 0, false, pp->lsoln != 0);
          int rc2 = allow_full_smem(s, sh.fn);
          if (rc2) return rc2;
          sh.tw = sh.teams = sh.team_bytes = sh.ctas = 0;
          const std::array<int, 8> cfg_key = {w1, words_for(n2max) * 4 + (pp->lorder != 0) * 2 + (pp->lsoln != 0), k.sm_query_bytes,
                                              sh.entry_bytes, k.sm_mapwords, k.sm_bmapwords, tw_max * 64 + teams_cap, n1max};
          auto hit = s->launch_cfg.find(cfg_key);
          if (hit != s->launch_cfg.end()) {
            sh.tw = hit->second[0]; sh.teams = hit->second[1]; sh.team_bytes = hit->second[2]; sh.ctas = hit->second[3];
          } else {
            int best_warps = -1;
            for (int tw = tw_max; tw >= 32; tw >>= 1) {
              if (tw & 31) continue;             // 96 -> 48: not a whole number of warps
              const int team_bytes = (int)round16(sh.entry_bytes + (size_t)(k.sm_mapwords + k.sm_bmapwords) * tw * 4 + SATS_K_SCRATCH_BYTES + (size_t)(tw / 32) * sh.qmask_bytes);
              const int teams_max = std::min(SATS_K_MAXTHREADS / tw, teams_cap);
              for (int teams = teams_max; teams >= 1; teams--) {
                size_t smem = SATS_K_BAR_BYTES + k.sm_query_bytes + SATS_K_ZTAB_BYTES + (size_t)teams * team_bytes;
                if (smem > (size_t)kMaxSmem) continue;
                int ctas = 0;
                CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, sh.fn, teams * tw, smem));
                if (ctas < 1) continue;            // does not fit at all
                int warps = ctas * teams * tw / 32;
                if (warps > best_warps) { best_warps = warps; sh.teams = teams; sh.tw = tw; sh.team_bytes = team_bytes; sh.ctas = ctas; }
              }
            }
            if (sh.teams) s->launch_cfg[cfg_key] = {sh.tw, sh.teams, sh.team_bytes, sh.ctas};
          }
          if (sh.teams == 0) return sats_fail(SATS_ERR_ARG, "query order %d x entry order %d does not fit in shared memory", n1max, n2max);
          *out = sh;
          return SATS_OK;
        };
        // end of the size class (kBucketBounds) that the entry at list position b belongs to (the list is decreasing)
        auto class_end = [&](int b) {
          int lowbound = 0;
          for (int bb : kBucketBounds) { if (bb >= s->sorted_order[b]) break; lowbound = bb; }
          int e = b;
          while (e < r1 && s->sorted_order[e] > lowbound) e++;
          return e;
        };
        static const bool no_merge = getenv("SATS_NO_MERGE") != nullptr;
        int b0 = r0;
        while (b0 < r1) {
          // A bucket = one launch, its shared memory sized for its largest entry.  It starts as one size class and then
          // absorbs the following (smaller) classes for as long as a launch of their own would not put more warps on an SM
          // and they run the same kernel variant: one work queue over more entries balances the persistent teams better than
          // several launches with a tail each -- which matters once a GPU holds only a 1/8 shard.
          const int n2max = s->sorted_order[b0];
          Shape sh;
          rc = shape_for(n2max, &sh);
          if (rc) return rc;
          int b1 = class_end(b0);
          while (!no_merge && b1 < r1 && words_for(s->sorted_order[b1]) == words_for(n2max)) {
            Shape nx;
            rc = shape_for(s->sorted_order[b1], &nx);
            if (rc) return rc;
            if (nx.ctas * nx.teams * nx.tw > sh.ctas * sh.teams * sh.tw) break;
            b1 = class_end(b1);
          }
          kernel_fn fn = sh.fn;
          k.sm_entry_bytes = sh.entry_bytes;
          k.sm_qmask_bytes = sh.qmask_bytes;
          const int tw = sh.tw, best_ctas = sh.ctas;
          k.tw = tw;
          k.teams = sh.teams;
          k.sm_team_bytes = sh.team_bytes;
          k.item_first = b0; k.item_count = b1 - b0;
          size_t smem = SATS_K_BAR_BYTES + k.sm_query_bytes + SATS_K_ZTAB_BYTES + (size_t)k.teams * k.sm_team_bytes;
          // persistent teams: no more CTAs than fit on the device at once (per query); each team keeps claiming entries
          k.counters = s->d_counters + counter_base;
          counter_base += (size_t)(q1 - q0);
          const int resident = std::max(1, best_ctas) * s->num_sms;
          dim3 grid((unsigned)std::min((k.item_count + k.teams - 1) / k.teams, resident), (unsigned)(q1 - q0));
          LaunchDesc ld;
          memset(&ld, 0, sizeof ld);
          ld.fn = fn; ld.grid_x = grid.x; ld.grid_y = grid.y; ld.threads = (unsigned)(k.teams * tw); ld.smem = smem; ld.k = k;
          plan.push_back(ld);
          b0 = b1;
        }
        q0 = q1;
      }
      // enqueue: counter reset, fork, the bucket kernels round-robin over the side streams (so that the tail of one bucket
      // overlaps the head of the next), join
      auto enqueue = [&]() -> int {
        CK(cudaMemsetAsync(s->d_counters, 0, ncounters * sizeof(int), s->stream));
        CK(cudaEventRecord(s->fork, s->stream));
        for (int i = 0; i < sats_searcher::kSide; i++) CK(cudaStreamWaitEvent(s->side[i], s->fork, 0));
        for (size_t n = 0; n < plan.size(); n++) {
          const LaunchDesc &ld = plan[n];
          ld.fn<<<dim3(ld.grid_x, ld.grid_y), ld.threads, ld.smem, s->side[n % sats_searcher::kSide]>>>(ld.k);
          CK(cudaGetLastError());
        }
        for (int i = 0; i < sats_searcher::kSide; i++) {
          CK(cudaEventRecord(s->side_done[i], s->side[i]));
          CK(cudaStreamWaitEvent(s->stream, s->side_done[i], 0));
        }
        return SATS_OK;
      };
      static const bool no_graph = getenv("SATS_NO_GRAPH") != nullptr;
      if (no_graph) {
        rc = enqueue();
        if (rc) return rc;
      } else {
        const bool same = s->graph_exec && s->graph_counters == ncounters && s->graph_plan.size() == plan.size() &&
                          (plan.empty() || memcmp(s->graph_plan.data(), plan.data(), plan.size() * sizeof(LaunchDesc)) == 0);
        if (!same) {
          if (s->graph_exec) { cudaGraphExecDestroy(s->graph_exec); s->graph_exec = nullptr; }
          cudaGraph_t graph = nullptr;
          CK(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
          rc = enqueue();
          cudaError_t ce = cudaStreamEndCapture(s->stream, &graph);
          if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
          CK(ce);
          ce = cudaGraphInstantiate(&s->graph_exec, graph, 0);
          cudaGraphDestroy(graph);
          CK(ce);
          s->graph_plan = plan;
          s->graph_counters = ncounters;
        }
        CK(cudaGraphLaunch(s->graph_exec, s->stream));
        key.ptr[8] = s->d_counters;          // may have been (re)allocated above
        memcpy(&s->graph_key, &key, sizeof key);      // byte copy: the comparison above is a memcmp, padding included
        s->graph_key_valid = true;
      }
      s->launches += (long long)plan.size();
    }
  }
  if (elapsed_ms) {
    CK(cudaEventRecord(s->ev1, s->stream));
    CK(cudaEventSynchronize(s->ev1));
    CK(cudaEventElapsedTime(elapsed_ms, s->ev0, s->ev1));
  }
  return SATS_OK;
}
SATS_CATCH_ALL

// enqueue the device -> pinned host copies of the last launch's results (no wait): lets a multi-GPU caller start every
// GPU's copy before it waits for the first
extern "C" int sats_search_collect_begin(sats_searcher *s)
try {
  if (!s) return sats_fail(SATS_ERR_ARG, "sats_search_collect_begin: null argument");
  CK(cudaSetDevice(s->device));
  const int D = (int)s->sorted_orig.size(), Q = s->last_q;
  if (D == 0 || Q < 1) return SATS_OK;
  size_t n = (size_t)Q * D;
  CK(cudaMemcpyAsync(s->h_scores, s->d_scores, n * 4, cudaMemcpyDeviceToHost, s->stream));
  if (s->last_lsoln) CK(cudaMemcpyAsync(s->h_maps, s->d_maps, n * SATS_K_MAPROW, cudaMemcpyDeviceToHost, s->stream));
  s->collect_pending = true;
  return SATS_OK;
}
SATS_CATCH_ALL

extern "C" int sats_search_collect(sats_searcher *s, int32_t *scores, int32_t *maps)
try {
  if (!s || !scores) return sats_fail(SATS_ERR_ARG, "sats_search_collect: null argument");
  if (s->last_lsoln && !maps) return sats_fail(SATS_ERR_ARG, "sats_search_collect: maps buffer required with lsoln");
  CK(cudaSetDevice(s->device));
  const int D = (int)s->sorted_orig.size(), Q = s->last_q;
  if (D == 0) return SATS_OK;
  if (!s->collect_pending) { int rc = sats_search_collect_begin(s); if (rc) return rc; }
  s->collect_pending = false;
  CK(cudaStreamSynchronize(s->stream));
  for (int q = 0; q < Q; q++) {
    const int n1 = s->q_n1[q];
    for (int kpos = 0; kpos < D; kpos++) {
      int v = s->h_scores[(size_t)q * D + kpos];
      if (v == kScoreSentinel) continue;
      size_t o = (size_t)s->slot_q[q] * s->db_count + s->sorted_orig[kpos];
      scores[o] = v;
      if (s->last_lsoln) {
        const int8_t *row = s->h_maps + ((size_t)q * D + kpos) * SATS_K_MAPROW;
        int32_t *dst = maps + o * SATS_MAP_STRIDE;
        for (int i = 0; i < n1; i++) dst[i] = row[i];
      }
    }
  }
  return SATS_OK;
}
SATS_CATCH_ALL

// Results of the last launch where they are: for callers that gather the shards' scores with their own collective (NCCL)
extern "C" int sats_search_device_results(sats_searcher *s, const int32_t **d_scores, int *qcount, int *entries, int32_t *slot_query)
{
  if (!s || !d_scores) return sats_fail(SATS_ERR_ARG, "sats_search_device_results: null argument");
  if (s->last_q < 1 || !s->d_scores) return sats_fail(SATS_ERR_ARG, "sats_search_device_results: no search has been launched");
  *d_scores = s->d_scores;
  if (qcount) *qcount = s->last_q;
  if (entries) *entries = (int)s->sorted_orig.size();
  if (slot_query) for (int q = 0; q < s->last_q; q++) slot_query[q] = s->slot_q[q];
  return SATS_OK;
}

extern "C" int sats_searcher_entry_index(const sats_searcher *s, int32_t *index)
{
  if (!s || !index) return sats_fail(SATS_ERR_ARG, "sats_searcher_entry_index: null argument");
  for (size_t k = 0; k < s->sorted_orig.size(); k++) index[k] = s->sorted_orig[k];
  return SATS_OK;
}

// ------------------------------------------------------------------------------------------------ top-k
// One CTA per query slot selects the k best-scoring entries of that query on the device, so only k (position, score)
// pairs cross PCIe instead of one score per database entry.  Scores are small integers, so an exact counting select
// does it: block max, histogram of (max - score), threshold bin, then an ordered pick -- everything above the threshold
// plus the first `need` entries AT the threshold in device order (= decreasing structure order, then file order).
#define SATS_TOPK_THREADS 1024
#define SATS_TOPK_BINS 4096

__device__ __forceinline__ int block_exclusive_scan(int v, int *scratch, int *total)
{
  // Hillis-Steele over SATS_TOPK_THREADS values in shared memory
  const int t = threadIdx.x;
  scratch[t] = v;
  __syncthreads();
  for (int off = 1; off < SATS_TOPK_THREADS; off <<= 1) {
    int add = t >= off ? scratch[t - off] : 0;
    __syncthreads();
    scratch[t] += add;
    __syncthreads();
  }
  const int incl = scratch[t];
  *total = scratch[SATS_TOPK_THREADS - 1];
  __syncthreads();
  return incl - v;
}

__global__ void __launch_bounds__(SATS_TOPK_THREADS)
sats_topk_kernel(const int32_t *scores, int stride, int count, int k, int32_t *out_pos, int32_t *out_score, int32_t *out_n)
{
  __shared__ int hist[SATS_TOPK_BINS];
  __shared__ int scratch[SATS_TOPK_THREADS];
  __shared__ int s_max, s_min, s_thr, s_need;
  const int q = blockIdx.x, t = threadIdx.x;
  const int32_t *row = scores + (size_t)q * stride;
  const int32_t none = (int32_t)0x80808080;
  // (a) maximum over computed entries
  int mx = INT_MIN, mn = INT_MAX;
  for (int e = t; e < count; e += SATS_TOPK_THREADS) { int v = row[e]; if (v != none) { mx = max(mx, v); mn = min(mn, v); } }
  mx = __reduce_max_sync(0xffffffffu, mx);
  mn = __reduce_min_sync(0xffffffffu, mn);
  if ((t & 31) == 0) { scratch[t >> 5] = mx; scratch[32 + (t >> 5)] = mn; }
  for (int b = t; b < SATS_TOPK_BINS; b += SATS_TOPK_THREADS) hist[b] = 0;
  __syncthreads();
  if (t == 0) {
    int m = INT_MIN, n = INT_MAX;
    for (int w = 0; w < SATS_TOPK_THREADS / 32; w++) { m = max(m, scratch[w]); n = min(n, scratch[32 + w]); }
    s_max = m; s_min = n;
  }
  __syncthreads();
  mx = s_max;
  // (b) histogram of max - score (scores further than BINS-1 below the maximum share the last bin)
  for (int e = t; e < count; e += SATS_TOPK_THREADS) {
    int v = row[e];
    if (v != none) atomicAdd(&hist[min(mx - v, SATS_TOPK_BINS - 1)], 1);
  }
  __syncthreads();
  // (c) threshold bin: the first bin at which the running count reaches k
  if (t == 0) {
    int run = 0, thr = SATS_TOPK_BINS - 1;
    for (int b = 0; b < SATS_TOPK_BINS; b++) { if (run + hist[b] >= k) { thr = b; break; } run += hist[b]; }
    s_thr = thr;
    s_need = k - run;              // how many entries of the threshold bin are still wanted (may exceed what exists)
  }
  __syncthreads();
  const int thr = s_thr, need = s_need;
  // scores spread over more than BINS values and the cut falls into the shared last bin: cannot order it here
  if (thr == SATS_TOPK_BINS - 1 && (long long)mx - (long long)s_min >= SATS_TOPK_BINS - 1) {
    if (t == 0) out_n[q] = -1;
    return;
  }
  // (d) ordered pick: thread t owns a contiguous chunk of the (device-ordered) entries
  const int chunk = (count + SATS_TOPK_THREADS - 1) / SATS_TOPK_THREADS;
  const int lo = min(count, t * chunk), hi = min(count, lo + chunk);
  int nsel = 0, ntie = 0;
  for (int e = lo; e < hi; e++) {
    int v = row[e];
    if (v == none) continue;
    int b = min(mx - v, SATS_TOPK_BINS - 1);
    nsel += b < thr;
    ntie += b == thr;
  }
  int total;
  const int ties_before = block_exclusive_scan(ntie, scratch, &total);
  const int quota = max(0, min(ntie, need - ties_before));
  const int base = block_exclusive_scan(nsel + quota, scratch, &total);
  if (t == 0) out_n[q] = min(total, k);
  int w = base, taken = 0;
  for (int e = lo; e < hi; e++) {
    int v = row[e];
    if (v == none) continue;
    int b = min(mx - v, SATS_TOPK_BINS - 1);
    bool take = b < thr || (b == thr && taken < quota);
    if (b == thr && taken < quota) taken++;
    if (take && w < k) { out_pos[(size_t)q * k + w] = e; out_score[(size_t)q * k + w] = v; w++; }
  }
}

extern "C" int sats_search_topk(sats_searcher *s, int k, int32_t *index_out, int32_t *score_out)
try {
  if (!s || !index_out || !score_out || k < 1) return sats_fail(SATS_ERR_ARG, "sats_search_topk: bad argument");
  if (s->last_q < 1) return sats_fail(SATS_ERR_ARG, "sats_search_topk: no search has been launched");
  CK(cudaSetDevice(s->device));
  const int D = (int)s->sorted_orig.size(), Q = s->last_q;
  size_t n = (size_t)Q * k;
  if (n > s->topk_cap) {
    cudaFree(s->d_topk); cudaFreeHost(s->h_topk);
    s->d_topk = nullptr; s->h_topk = nullptr; s->topk_cap = 0;
    CK(cudaMalloc(&s->d_topk, (2 * n + Q) * 4));
    CK(cudaMallocHost(&s->h_topk, (2 * n + Q) * 4));
    s->topk_cap = n;
  }
  int32_t *d_pos = s->d_topk, *d_sc = s->d_topk + n, *d_n = s->d_topk + 2 * n;
  sats_topk_kernel<<<Q, SATS_TOPK_THREADS, 0, s->stream>>>(s->d_scores, std::max(1, D), D, k, d_pos, d_sc, d_n);
  CK(cudaGetLastError());
  s->launches++;
  CK(cudaMemcpyAsync(s->h_topk, s->d_topk, (2 * n + Q) * 4, cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  const int32_t *h_pos = s->h_topk, *h_sc = s->h_topk + n, *h_n = s->h_topk + 2 * n;
  std::vector<int> order((size_t)k);
  std::vector<int32_t> rowbuf, rowpos;
  for (int q = 0; q < Q; q++) {
    int got = h_n[q];
    if (got < 0) {
      // rare fallback (see kernel): select this query's row on the host with the same ordering rule
      rowbuf.resize((size_t)D);
      CK(cudaMemcpy(rowbuf.data(), s->d_scores + (size_t)q * std::max(1, D), (size_t)D * 4, cudaMemcpyDeviceToHost));
      rowpos.clear();
      for (int e = 0; e < D; e++) if (rowbuf[e] != kScoreSentinel) rowpos.push_back(e);
      got = std::min<int>(k, (int)rowpos.size());
      std::partial_sort(rowpos.begin(), rowpos.begin() + got, rowpos.end(), [&](int a, int b) {
        if (rowbuf[a] != rowbuf[b]) return rowbuf[a] > rowbuf[b];
        return a < b;
      });
      for (int i = 0; i < k; i++) {
        index_out[(size_t)s->slot_q[q] * k + i] = i < got ? s->sorted_orig[rowpos[i]] : -1;
        score_out[(size_t)s->slot_q[q] * k + i] = i < got ? rowbuf[rowpos[i]] : INT_MIN;
      }
      continue;
    }
    std::iota(order.begin(), order.begin() + got, 0);
    std::sort(order.begin(), order.begin() + got, [&](int a, int b) {
      int sa = h_sc[(size_t)q * k + a], sb = h_sc[(size_t)q * k + b];
      if (sa != sb) return sa > sb;
      return h_pos[(size_t)q * k + a] < h_pos[(size_t)q * k + b];
    });
    for (int i = 0; i < k; i++) {
      if (i < got) {
        index_out[(size_t)s->slot_q[q] * k + i] = s->sorted_orig[h_pos[(size_t)q * k + order[i]]];
        score_out[(size_t)s->slot_q[q] * k + i] = h_sc[(size_t)q * k + order[i]];
      } else {
        index_out[(size_t)s->slot_q[q] * k + i] = -1;
        score_out[(size_t)s->slot_q[q] * k + i] = INT_MIN;
      }
    }
  }
  return SATS_OK;
}
SATS_CATCH_ALL

// ---- device-side significance cut ---------------------------------------------------------------------
// One CTA per query slot.  thr[q][n2] is the smallest raw score whose Gumbel z-score reaches the cut for a structure of
// order n2 (built on the host with the very functions the result printer uses, so the cut is exact).  Ordered compaction:
// thread t owns a contiguous chunk of the device-ordered entries, a block scan places its hits; the CTA then reserves a
// contiguous range of the shared pair buffer with one atomicAdd, so that only the hits have to travel to the host.
__global__ void __launch_bounds__(SATS_TOPK_THREADS)
sats_hits_kernel(const int32_t *scores, int stride, int count, const int32_t *orders, const int32_t *thr, int cap,
                 int32_t *out_n, int32_t *out_base, int *cursor, int2 *pairs)
{
  __shared__ int scratch[SATS_TOPK_THREADS];
  __shared__ int s_thr[SATS_MAXDIM_EXT + 1];
  __shared__ int s_base;
  const int q = blockIdx.x, t = threadIdx.x;
  const int32_t *row = scores + (size_t)q * stride;
  const int32_t none = (int32_t)0x80808080;
  for (int n = t; n <= SATS_MAXDIM_EXT; n += SATS_TOPK_THREADS) s_thr[n] = thr[(size_t)q * (SATS_MAXDIM_EXT + 1) + n];
  __syncthreads();
  const int chunk = (count + SATS_TOPK_THREADS - 1) / SATS_TOPK_THREADS;
  const int lo = min(count, t * chunk), hi = min(count, lo + chunk);
  int mine = 0;
  for (int e = lo; e < hi; e++) { int v = row[e]; mine += v != none && v >= s_thr[orders[e]]; }
  int total;
  int w = block_exclusive_scan(mine, scratch, &total);
  if (t == 0) {
    out_n[q] = total;
    s_base = atomicAdd(cursor, min(total, cap));
    out_base[q] = s_base;
  }
  __syncthreads();
  int2 *dst = pairs + s_base;
  for (int e = lo; e < hi && w < cap; e++) {
    int v = row[e];
    if (v != none && v >= s_thr[orders[e]]) dst[w++] = make_int2(e, v);
  }
}

extern "C" int sats_search_hits(sats_searcher *s, double z_min, int cap, int32_t *count_out, int32_t *index_out, int32_t *score_out)
try {
  if (!s || !count_out || !index_out || !score_out || cap < 1) return sats_fail(SATS_ERR_ARG, "sats_search_hits: bad argument");
  if (s->last_q < 1) return sats_fail(SATS_ERR_ARG, "sats_search_hits: no search has been launched");
  if (!(z_min == z_min)) return sats_fail(SATS_ERR_ARG, "sats_search_hits: z_min is NaN");
  CK(cudaSetDevice(s->device));
  const int D = (int)s->sorted_orig.size(), Q = s->last_q, T = SATS_MAXDIM_EXT + 1;
  const int kcap = std::min(cap, std::max(1, D));      // rows the device keeps per query; the caller's rows stay `cap` wide
  // one integer score threshold per (query slot, structure order): see sats_score_threshold()
  std::vector<int32_t> thr((size_t)Q * T);
  for (int q = 0; q < Q; q++)
    for (int n2 = 0; n2 < T; n2++) thr[(size_t)q * T + n2] = n2 == 0 ? INT_MAX : sats_score_threshold(z_min, s->q_n1[q], n2);
  // device / pinned buffer: [Q counts][Q bases][cursor][pad][thresholds Q x T][pairs: int2 x Q x cap]
  const size_t head = 2 * (size_t)Q + 2, pair_off = (head + thr.size() + 1) & ~(size_t)1, words = pair_off + 2 * (size_t)Q * kcap;
  if (words > s->hits_cap) {
    cudaFree(s->d_hits); cudaFreeHost(s->h_hits);
    s->d_hits = nullptr; s->h_hits = nullptr; s->hits_cap = 0;
    CK(cudaMalloc(&s->d_hits, words * 4));
    CK(cudaMallocHost(&s->h_hits, words * 4));
    s->hits_cap = words;
  }
  memset(s->h_hits, 0, head * 4);
  memcpy(s->h_hits + head, thr.data(), thr.size() * 4);
  CK(cudaMemcpyAsync(s->d_hits, s->h_hits, (head + thr.size()) * 4, cudaMemcpyHostToDevice, s->stream));
  sats_hits_kernel<<<Q, SATS_TOPK_THREADS, 0, s->stream>>>(s->d_scores, std::max(1, D), D, s->d_sorted_order, s->d_hits + head, kcap,
                                                           s->d_hits, s->d_hits + Q, reinterpret_cast<int *>(s->d_hits + 2 * Q),
                                                           reinterpret_cast<int2 *>(s->d_hits + pair_off));
  CK(cudaGetLastError());
  s->launches++;
  CK(cudaMemcpyAsync(s->h_hits, s->d_hits, head * 4, cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  const int32_t *h_n = s->h_hits, *h_base = s->h_hits + Q;
  const size_t kept = (size_t)s->h_hits[2 * Q];                       // hits actually stored, all queries together
  if (kept) CK(cudaMemcpyAsync(s->h_hits + pair_off, s->d_hits + pair_off, kept * 8, cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  const int32_t *h_pairs = s->h_hits + pair_off;
  for (int q = 0; q < Q; q++) {
    const size_t row = (size_t)s->slot_q[q];          // output rows follow the batch, device rows the slots
    count_out[row] = h_n[q];
    const int got = std::min(h_n[q], kcap);
    const int32_t *pr = h_pairs + 2 * (size_t)h_base[q];
    for (int i = 0; i < cap; i++) {
      index_out[row * cap + i] = i < got ? s->sorted_orig[pr[2 * i]] : -1;
      score_out[row * cap + i] = i < got ? pr[2 * i + 1] : INT_MIN;
    }
  }
  return SATS_OK;
}
SATS_CATCH_ALL

// ---- streaming hits (second half) ----------------------------------------------------------------------
extern "C" int sats_search_bind_cut(sats_searcher *s, double z_min)
{
  if (!s) return sats_fail(SATS_ERR_ARG, "sats_search_bind_cut: null searcher");
  s->cut_bound = z_min == z_min;         // NaN unbinds
  s->cut_z = z_min;
  return SATS_OK;
}

extern "C" int sats_search_streamed_hits(sats_searcher *s, int cap, int32_t *count_out, int32_t *index_out, int32_t *score_out,
                                         int64_t *d2h_bytes)
try {
  if (!s || !count_out || !index_out || !score_out || cap < 1) return sats_fail(SATS_ERR_ARG, "sats_search_streamed_hits: bad argument");
  if (!s->stream_valid) return sats_fail(SATS_ERR_ARG, "sats_search_streamed_hits: the last launch ran without a bound cut (sats_search_bind_cut)");
  CK(cudaSetDevice(s->device));
  const int Q = s->last_q;
  unsigned total = 0;
  CK(cudaMemcpyAsync(&total, s->d_stream_cursor, sizeof total, cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  if (d2h_bytes) *d2h_bytes = (int64_t)sizeof total;
  if ((size_t)total > s->stream_list_cap) {
    // more hits than the list holds (a cut that keeps nearly everything): the dense scores are still on the device
    if (d2h_bytes) *d2h_bytes += (int64_t)Q * 8 + (int64_t)std::min<long long>((long long)total, (long long)Q * cap) * 8;
    return sats_search_hits(s, s->cut_z, cap, count_out, index_out, score_out);
  }
  if ((size_t)total > s->h_stream_cap) {
    cudaFreeHost(s->h_stream); s->h_stream = nullptr; s->h_stream_cap = 0;
    CK(cudaMallocHost(&s->h_stream, (size_t)total * sizeof(int2)));
    s->h_stream_cap = total;
  }
  if (total) {
    CK(cudaMemcpyAsync(s->h_stream, s->d_stream_list, (size_t)total * sizeof(int2), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    if (d2h_bytes) *d2h_bytes += (int64_t)total * (int64_t)sizeof(int2);
  }
  // the kernel appends in completion order: bring every query's hits into device order (decreasing structure order, then
  // file order) -- the order sats_search_hits() and the -z printer use
  std::vector<std::vector<std::pair<int32_t, int32_t>>> per((size_t)Q);
  for (unsigned k = 0; k < total; k++) {
    const int2 h = s->h_stream[k];
    const int slot = (int)((uint32_t)h.y & 0xffffu), score = h.y >> 16;          // arithmetic shift: the score is signed
    if (slot >= Q) return sats_fail(SATS_ERR_CUDA, "corrupt hit record");
    per[(size_t)slot].emplace_back(h.x, score);
  }
  for (int q = 0; q < Q; q++) {
    auto &v = per[(size_t)q];
    std::sort(v.begin(), v.end());
    const size_t row = (size_t)s->slot_q[q];
    count_out[row] = (int32_t)v.size();
    for (int i = 0; i < cap; i++) {
      const bool have = (size_t)i < v.size();
      index_out[row * cap + i] = have ? s->sorted_orig[v[(size_t)i].first] : -1;
      score_out[row * cap + i] = have ? v[(size_t)i].second : INT_MIN;
    }
  }
  return SATS_OK;
}
SATS_CATCH_ALL

extern "C" int sats_search(sats_searcher *s, const sats_db *queries, int qfirst, int qcount, const sats_params *params,
                           uint32_t query_index_base, int32_t *scores, int32_t *maps)
{
  int rc = sats_search_upload(s, queries, qfirst, qcount);
  if (rc) return rc;
  rc = sats_search_launch(s, params, query_index_base, nullptr);
  if (rc) return rc;
  return sats_search_collect(s, scores, maps);
}
