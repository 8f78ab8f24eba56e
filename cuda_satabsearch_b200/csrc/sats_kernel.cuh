// sats_kernel.cuh -- the simulated-annealing tableau-search kernel for sm_100a.
//
// Replaces the body of sa_tabsearch_gpu (reference nvcc_src_current/cudaSaTabsearch_kernel.cu:804-1236)
// and its helpers tscord (:306), tmscord (:396), deltasd (:502), thinit (:588), randtypeind (:677).
// It is a new design, not a translation:
//
//   * One LANE runs one restart chain.  A TEAM of TW (32/64/128) threads shares one database entry and covers
//     `restarts` chains in rounds of TW; a CTA holds several teams that share one copy of the query.  Teams are
//     persistent: each claims entries of the launch's size-sorted list from a global counter and synchronises with one
//     named barrier per entry (the arg-max), behind which its leader starts the next entry's copy.
//   * The query blob and every entry blob are brought into shared memory by TMA 1-D bulk copies (cp.async.bulk ...
//     mbarrier::complete_tx) issued by one thread, each team on its own mbarrier.
//   * Chain state is bit masks in registers (mapped query SSEs, occupied entry SSEs) plus a map of one 32-bit word per
//     query SSE (the partner's index) in lane-private, bank-conflict-free shared memory.  The LORDER window and the candidate list of the reference (linear scans, kernel.cu:1053-1083, :677-714)
//     become O(1) mask arithmetic: bfind/ffs for the neighbouring mapped SSEs, (type mask & ~occupied & range mask) for
//     the candidates, popc/select-nth for the random pick.  deltasd walks only the *mapped* SSEs (set bits), reading one
//     8-byte {distance, code} cell per operand; zeta is a 128-byte shared-memory table indexed by the XOR of two codes;
//     a missing side of a move reads a row of NaN distances, so the loop body is branch-free.
//   * LSOLN: the best map is kept as (dirty mask, first-change saves) and completed once per chain, never copied
//     wholesale.
//   * No per-move float conversions: every decision the reference takes on a uniform u = unit(x) of 32 random bits x
//     is a monotone function of x, so the host tabulates it as integer cut-offs with the reference's exact fp32 / fp64
//     arithmetic -- the SSE pick (int)((u - 1.1e-7) * n1) (kernel.cu:1042) becomes umulhi(x, n1) corrected by one
//     shared-memory threshold, the Metropolis test expf((float)delta / T_m) > u (kernel.cu:1166) becomes x < cut[m][-delta].
//     In DEVICE_FAST mode (validation streams only) the test uses the fast-math intrinsics of the reference's GPU build.
//   * Uniforms: Philox4x32-10 at static positions (production: one block per two moves, one per seeding pass) or the
//     reference's XORWOW grid streams consumed in the reference's order (validation).
//   * Restart arg-max: redux.sync inside each warp, then across the team's warps through shared
//     memory, with the reference's tie-break (kernel.cu:1205-1221).
#ifndef SATS_KERNEL_CUH
#define SATS_KERNEL_CUH

#include <cstdint>
#include <cuda_runtime.h>

#include "sats.h"
#include "sats_kparams.h"

namespace satsk {

// ------------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  uint32_t done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void team_sync(int team, int tw)
{
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(tw) : "memory");
}

// ------------------------------------------------------------------------------------------------ bit sets
// Bit b of a W-word mask, as seen from word w: 1 << (b - 32w) when that lands inside the word, else 0.  PTX shl.b32
// clamps shift amounts above 31 to a zero result, and a negative offset is a huge unsigned amount, so one shift does it
// -- and, unlike "if ((b >> 5) == w)", it cannot be turned into a dynamically indexed (local-memory) array access.
__device__ __forceinline__ uint32_t bit_in_word(int b, int w)
{
  uint32_t r;
  asm("shl.b32 %0, 1, %1;" : "=r"(r) : "r"((uint32_t)(b - 32 * w)));
  return r;
}
template <int W> __device__ __forceinline__ bool bit_test(const uint32_t (&m)[W], int b)
{
  if (W == 1) return (m[0] >> (b & 31)) & 1u;
  uint32_t hit = 0u;
#pragma unroll
  for (int w = 0; w < W; w++) hit |= m[w] & bit_in_word(b, w);
  return hit != 0u;
}
template <int W> __device__ __forceinline__ void bit_set(uint32_t (&m)[W], int b)
{
  if (W == 1) { m[0] |= 1u << (b & 31); return; }
#pragma unroll
  for (int w = 0; w < W; w++) m[w] |= bit_in_word(b, w);
}
template <int W> __device__ __forceinline__ void bit_clear(uint32_t (&m)[W], int b)
{
  if (W == 1) { m[0] &= ~(1u << (b & 31)); return; }
#pragma unroll
  for (int w = 0; w < W; w++) m[w] &= ~bit_in_word(b, w);
}
// mask of bit positions < x within one word (x may be <= 0 or >= 32).  PTX shl.b32 clamps shift amounts above 31
// (the result is 0), so ~(~0 << max(x, 0)) is exact on the whole range without branches.
__device__ __forceinline__ uint32_t below(int x)
{
  uint32_t r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(0xffffffffu), "r"((uint32_t)max(x, 0)));
  return ~r;
}

// the same for x >= 0: no clamp needed
__device__ __forceinline__ uint32_t below_nonneg(int x)
{
  uint32_t r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(0xffffffffu), "r"((uint32_t)x));
  return ~r;
}

// index of the highest set bit (x != 0): one FLO
__device__ __forceinline__ int top_bit(uint32_t x)
{
  int r;
  asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}
// mask of bit positions < n, 0 <= n <= 32: one BMSK
__device__ __forceinline__ uint32_t bits_below(int n)
{
  uint32_t r;
  asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(r) : "r"(0), "r"(n));
  return r;
}

// highest set bit with index <= i, or -1 (0 <= i < 32 W).  bfind of 0 is -1, so one word needs no select at all.
template <int W> __device__ __forceinline__ int top_at_or_below(const uint32_t (&m)[W], int i)
{
  if (W == 1) return top_bit(m[0] & bits_below(i + 1));
  int r = -1;
#pragma unroll
  for (int w = 0; w < W; w++) {
    uint32_t x = m[w] & below(i + 1 - 32 * w);
    if (x) r = 32 * w + top_bit(x);
  }
  return r;
}
// lowest set bit with index >= i, or -1 (0 <= i <= 32 W)
template <int W> __device__ __forceinline__ int low_at_or_above(const uint32_t (&m)[W], int i)
{
  if (W == 1) return __ffs(m[0] & ~bits_below(i)) - 1;
  int r = -1;
#pragma unroll
  for (int w = W - 1; w >= 0; w--) {
    uint32_t x = m[w] & ~below(i - 32 * w);
    if (x) r = 32 * w + __ffs(x) - 1;
  }
  return r;
}
// index of the idx-th (0-based) set bit; idx < popcount
template <int W> __device__ __forceinline__ int select_nth(const uint32_t (&c)[W], int idx)
{
  uint32_t word = c[0];
  int base = 0;
  if (W > 1) {
    bool found = false;
#pragma unroll
    for (int w = 0; w < W; w++) {
      int pc = __popc(c[w]);
      if (!found) {
        if (idx < pc) { word = c[w]; base = 32 * w; found = true; }
        else idx -= pc;
      }
    }
  }
  // kept rolled: idx is 0..2 almost always (two or three candidates), and the unrolled-by-four form ptxas makes of this loop
  // runs ~25 instructions of remainder handling for the handful of lanes that get here
#pragma unroll 1
  for (; idx > 0; idx--) word &= word - 1u;
  return base + __ffs(word) - 1;
}

// ------------------------------------------------------------------------------------------------ uniforms
__device__ __forceinline__ float unit_from_bits(uint32_t x)
{
  // cuRAND's _curand_uniform: (0, 1]; the product with 2^-32 is exact so FMA contraction cannot change it
  return (float)x * 2.3283064e-10f + 1.1641532e-10f;
}
// (u - 1.1e-7) * n in double, truncated (kernel.cu:67, :1042, :710)
__device__ __forceinline__ int scaled_index(float u, int n)
{
  // u >= 2^-33, so the product is > -1 and the truncating conversion already maps the negative sliver to 0
  return (int)(((double)u - 1.1e-7) * (double)n);
}

// Philox4x32-10 (Salmon et al., SC'11).  rk = the ten round keys (k0 + r * 0x9E3779B9, k1 + r * 0xBB67AE85), precomputed on
// the host into the kernel parameters: they reach the XORs as uniform operands instead of costing two adds per round in
// every thread (6 instead of 8 instructions per round).  The products stay IMAD.HI + IMAD: one IMAD.WIDE each (mul.wide.u32)
// measured 0.5 % slower (register pairs under the 56-register cap; profiles/r02_experiments.txt).
__device__ __forceinline__ void mul_wide(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo)
{
  hi = __umulhi(a, b); lo = a * b;
}
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               const uint32_t (&rk)[20], uint32_t (&out)[4])
{
#ifndef SATS_EXP_ROUNDS
#define SATS_EXP_ROUNDS 10      // experiments only (profiles/r02_experiments.txt): anything else is not Philox4x32-10
#endif
#pragma unroll
  for (int r = 0; r < SATS_EXP_ROUNDS; r++) {
    uint32_t h0, l0, h1, l1;
    mul_wide(0xD2511F53u, c0, h0, l0);
    mul_wide(0xCD9E8D57u, c2, h1, l1);
    c0 = h1 ^ c1 ^ rk[2 * r]; c1 = l1; c2 = h0 ^ c3 ^ rk[2 * r + 1]; c3 = l0;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Xorwow {
  uint32_t d, v0, v1, v2, v3, v4;
  __device__ __forceinline__ void load(const uint32_t *p) { d = p[0]; v0 = p[1]; v1 = p[2]; v2 = p[3]; v3 = p[4]; v4 = p[5]; }
  __device__ __forceinline__ void store(uint32_t *p) const { p[0] = d; p[1] = v0; p[2] = v1; p[3] = v2; p[4] = v3; p[5] = v4; }
  __device__ __forceinline__ uint32_t next()      // the 32 bits curand_uniform turns into a float
  {
    uint32_t t = v0 ^ (v0 >> 2);
    v0 = v1; v1 = v2; v2 = v3; v3 = v4;
    v4 = (v4 ^ (v4 << 4)) ^ (t ^ (t << 1));
    d += 362437u;
    return v4 + d;
  }
};

// The SSE pick of a move, (int)((unit(x) - 1.1e-7) * n1) in double (kernel.cu:1042), straight from the 32 random bits:
// unit() is monotone and the pick lies at most one below floor(x * n1 / 2^32) (the shift by 1.1e-7 is less than one step
// for every n1 <= 111), so one multiply-high and one compare with the host-tabulated exact boundary of that step do it.
// pick_cut[k] = the smallest x whose pick is >= k (pick_cut[0] = 0): n1 words behind the SSE types in the query header.
__device__ __forceinline__ int pick_index(uint32_t x, int n1, uint32_t pick_cut)
{
  const uint32_t s0 = __umulhi(x, (uint32_t)n1);
  uint32_t t;
  asm("ld.shared.b32 %0, [%1];" : "=r"(t) : "r"(pick_cut + s0 * 4u));
  return (int)s0 - (x < t ? 1 : 0);
}
// the production streams' candidate draw: Fibonacci hash of the pick word (stream layout v2)
__device__ __forceinline__ uint32_t candidate_bits(uint32_t x1) { return x1 * 0x9E3779B9u; }

// ------------------------------------------------------------------------------------------------ scoring
// tscord (kernel.cu:306-332) as a 128-byte table in shared memory.  A device cell carries its tableau code as
// (first letter << 4) | second letter (letters 0..4), so q.code ^ e.code has a zero high field iff the first letters agree
// and a zero low field iff the second letters agree, and zeta = table[q.code ^ e.code].  The table is 128-byte aligned and
// exactly 32 words long: every word sits in its own bank, so lookups never conflict, and its address can be OR-ed into the
// staged query cells once per CTA -- then zeta costs one LOP3 and one LDS.S8 per operand and nothing on the XU pipe
// (AND + POPC + select on one-hot codes, the previous form, kept that pipe the second busiest).
__device__ __forceinline__ void fill_zeta_table(int8_t *tab, int tid, int nthreads)
{
  for (int x = tid; x < 128; x += nthreads) {
    const int hits = ((x >> 4) == 0) + ((x & 15) == 0);
    tab[x] = (int8_t)(hits ? hits : -2);
  }
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
  uint32_t v;
  asm("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr)
{
  uint2 v;
  asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
  return v;
}
// q.y = table address | query code (see above), e.y = entry code
__device__ __forceinline__ int gated(uint2 q, uint2 e)
{
  int z;
  asm("ld.shared.s8 %0, [%1];" : "=r"(z) : "r"(q.y ^ e.y) : "memory");
  float gap = fabsf(__uint_as_float(q.x) - __uint_as_float(e.x));
  return gap <= 4.0f ? z : 0;
}
// once per CTA, after the query blob has landed: OR the table address into the code word of every staged query cell
template <int W1> __device__ __forceinline__ void stamp_query_cells(uint8_t *sq, int n1, uint32_t ztab)
{
  if (W1 <= 2) {
    uint32_t *cell = reinterpret_cast<uint32_t *>(sq + SATS_K_QUERY_HDR);
    for (int c = threadIdx.x; c < n1 * n1; c += blockDim.x) cell[2 * c + 1] |= ztab;
  }
  __syncthreads();
}
// a query cell read from global memory (queries of more than 64 SSEs) still lacks the table address
template <bool PATCH> __device__ __forceinline__ uint2 with_table(uint2 q, uint32_t ztab)
{
  if (PATCH) q.y |= ztab;
  return q;
}

// Per-team view of shared memory (shared-window byte addresses)
struct TeamView {
  const uint2 *qcell_g;    // W1 == 4 only: the query's n1 x n1 cells in global memory
  uint32_t qcell;          // n1 x n1 {distance bits, code} (W1 <= 2)
  const uint8_t *qtype;    // n1
  uint32_t pick_cut;       // n1 words: exact boundaries of the SSE pick (pick_index)
  uint32_t ecell;          // n2 x n2, preceded by "row -1": n2 cells of NaN distance (the missing side of a move)
  uint32_t ztab;           // the zeta table (128-byte aligned)
  const uint32_t *tmask;   // [4][4] type -> 128-bit mask of entry SSEs of that type
  uint32_t qmask;          // [n1][W2] per query SSE: the mask of entry SSEs of its type (built per entry, one load per move)
  uint32_t smap;           // this lane's live map (Map<W1 <= 2>)
  uint32_t bmap;           // this lane's best map (Map<false>)
  uint32_t mstride;        // tw * 4: consecutive lanes own consecutive banks, so lane-private accesses never conflict
  int n1, n2;
};

// Lane-private maps (query SSE -> partner entry SSE) in two representations, both laid out so that consecutive lanes own
// consecutive banks (stride = tw * 4 bytes between a lane's successive words):
//   Map<true>   one 32-bit word per query SSE holding the partner (-1 = unmapped): one IMAD to address an element, and one
//               IMAD turns the loaded partner into the address of its 8-byte cell in a row (holding 8 * partner instead cost
//               two shifts per move in the window logic).  Used for the live map of queries of <= 64 SSEs.
//   Map<false>  one byte per query SSE (0xff = unmapped), four to a word.  A quarter of the shared memory, three more
//               instructions per access: used for the live map of larger queries and for every best-so-far map.
template <bool WIDE> struct Map {
  static __device__ __forceinline__ uint32_t addr(uint32_t base, int k, uint32_t stride)
  {
    if (WIDE) return base + (uint32_t)k * stride;
    return (uint32_t)(k >> 2) * stride + (base | (uint32_t)(k & 3));      // the lane's slot is 4-byte aligned
  }
  // 8 * partner of a MAPPED query SSE: the byte offset of the partner's cell within a row of the entry matrix
  static __device__ __forceinline__ uint32_t off8(uint32_t base, int k, uint32_t stride)
  {
    uint32_t v;
    if (WIDE) asm("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr(base, k, stride)) : "memory");
    else asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr(base, k, stride)) : "memory");
    return v << 3;          // folds into the IMAD that adds the row address
  }
  // partner of a query SSE, -1 if unmapped
  static __device__ __forceinline__ int get(uint32_t base, int k, uint32_t stride)
  {
    int v;
    if (WIDE) { asm("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr(base, k, stride)) : "memory"); return v; }
    asm("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"(addr(base, k, stride)) : "memory");
    return v;
  }
  static __device__ __forceinline__ void put(uint32_t base, int k, uint32_t stride, int j)      // j = -1 unmaps
  {
    if (WIDE) asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr(base, k, stride)), "r"(j) : "memory");
    else asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr(base, k, stride)), "r"(j) : "memory");
  }
  static __device__ __forceinline__ int words(int n1) { return WIDE ? n1 : (n1 + 3) >> 2; }
  static __device__ __forceinline__ void clear(uint32_t base, int n1, uint32_t stride)
  {
#pragma unroll 1
    for (int w = 0; w < words(n1); w++)
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + (uint32_t)w * stride), "r"(-1) : "memory");
  }
};
template <int W1, int W2, bool LORDER, bool XORWOW, bool LSOLN>
struct Chain {
  typedef Map<(W1 <= 2)> LiveMap;
  uint32_t mq[W1];   // query SSEs currently mapped
  uint32_t md[W2];   // entry SSEs currently occupied
  int score;
  // LSOLN: the best map is never copied wholesale when the best score improves.  An improving move has d > 0 and is
  // therefore accepted, so right after it best map == live map.  From then on the first accepted change of a query SSE
  // saves that SSE's partner *as it was at the best moment* into v.bmap and marks the SSE dirty; so at any time
  //     best map[k] = dirty[k] ? v.bmap[k] : live map[k],
  // and the clean entries are filled in once, when the chain ends.  `owned`: this chain holds the lane's best so far
  // (otherwise v.bmap still carries an earlier chain's best map and must not be touched).
  uint32_t dirty[W1];
  bool owned;

  __device__ __forceinline__ void best_is_live()
  {
    owned = true;
#pragma unroll
    for (int w = 0; w < W1; w++) dirty[w] = 0u;
  }
  __device__ __forceinline__ void finish_best_map(const TeamView &v)
  {
    if (!owned) return;
#pragma unroll 1
    for (int k = 0; k < v.n1; k++)
      if (!bit_test<W1>(dirty, k)) Map<false>::put(v.bmap, k, v.mstride, LiveMap::get(v.smap, k, v.mstride));
  }

  __device__ __forceinline__ void seed_clear(const TeamView &v)
  {
#pragma unroll
    for (int w = 0; w < W1; w++) mq[w] = 0u;
#pragma unroll
    for (int w = 0; w < W2; w++) md[w] = 0u;
    LiveMap::clear(v.smap, v.n1, v.mstride);
  }
  // match query SSE i to the next free entry SSE of its type at or after next_j; false = none left (thinit returns)
  __device__ __forceinline__ bool seed_pair(const TeamView &v, int i, int &next_j)
  {
    uint32_t cand[W2];
#pragma unroll
    for (int w = 0; w < W2; w++) cand[w] = lds32(v.qmask + (uint32_t)(i * W2 + w) * 4u);
    const int j = low_at_or_above<W2>(cand, next_j);
    if (j < 0) return false;
    LiveMap::put(v.smap, i, v.mstride, j);
    bit_set<W1>(mq, i);
    bit_set<W2>(md, j);
    next_j = j + 1;
    return true;
  }
  // thinit (kernel.cu:588-648): walk the query, with probability 1/2 match SSE i to the next entry SSE of its type.
  // Validation streams: one draw per query SSE until the entry runs out of SSEs of the wanted type.
  template <class Attempt> __device__ __forceinline__ void seed(const TeamView &v, Attempt &&attempt)
  {
    seed_clear(v);
    int next_j = 0;
    for (int i = 0; i < v.n1; i++)
      if (attempt(i) && !seed_pair(v, i, next_j)) break;
  }
  // Production streams: the seeding pass owns one random bit per query SSE (att: the Philox seeding block), so only the
  // SSEs that do attempt a match are visited.
  __device__ __forceinline__ void seed_bits(const TeamView &v, const uint32_t (&att)[4])
  {
    seed_clear(v);
    int next_j = 0;
    bool alive = true;
#pragma unroll
    for (int w = 0; w < (W1 < 4 ? W1 : 4); w++) {
      uint32_t a = att[w] & below(v.n1 - 32 * w);
      while (a && alive) {
        const int i = 32 * w + __ffs(a) - 1;
        a &= a - 1u;
        alive = seed_pair(v, i, next_j);
      }
    }
  }

  // tmscord (kernel.cu:396-440) over mapped pairs only
  __device__ __forceinline__ int full_score(const TeamView &v) const
  {
    int total = 0;
#pragma unroll
    for (int wi = 0; wi < W1; wi++) {
      uint32_t bi = mq[wi];
      while (bi) {
        int i = 32 * wi + __ffs(bi) - 1;
        bi &= bi - 1u;
        const uint32_t erow = v.ecell + LiveMap::off8(v.smap, i, v.mstride) * (uint32_t)v.n2;
        const uint32_t qrow = v.qcell + (uint32_t)(i * v.n1) * 8u;
        const uint2 *qrow_g = v.qcell_g + i * v.n1;
#pragma unroll
        for (int wk = 0; wk < W1; wk++) {
          if (wk < wi) continue;
          uint32_t bk = mq[wk];
          if (wk == wi) bk &= ~below((i & 31) + 1);
          while (bk) {
            int k = 32 * wk + __ffs(bk) - 1;
            bk &= bk - 1u;
            total += gated(W1 > 2 ? with_table<true>(__ldg(qrow_g + k), v.ztab) : lds64(qrow + (uint32_t)k * 8u),
                           lds64(erow + LiveMap::off8(v.smap, k, v.mstride)));
          }
        }
      }
    }
    return total;
  }

  // deltasd (kernel.cu:502-535) over mapped SSEs only.  Branch-free body: a missing `from` / `to` side reads the row
  // of NaN distances, whose gate never opens, so lanes with one-sided and two-sided moves run the same instructions.
  // The mapped SSEs are walked from the top bit down (one FLO + one BMSK per term; the sum does not care about order).
  __device__ __forceinline__ int delta(const TeamView &v, int i, int from, int to) const
  {
    int d = 0;
    const uint2 *qrow_g = v.qcell_g + i * v.n1;                         // W1 == 4: query cells live in global memory
    uint32_t qrow = W1 > 2 ? 0u : v.qcell + (uint32_t)(i * v.n1) * 8u;  // row base addresses, hoisted by hand
    uint32_t frow = v.ecell + (uint32_t)(from * v.n2) * 8u;         // from / to = -1: the NaN row in front of the matrix
    uint32_t trow = v.ecell + (uint32_t)(to * v.n2) * 8u;
    asm volatile("" : "+r"(qrow), "+r"(frow), "+r"(trow));       // keep the compiler from re-folding them into the loop
#pragma unroll
    for (int w = 0; w < W1; w++) {
      uint32_t b = mq[w];
      if (W1 == 1 || (i >> 5) == w) b &= ~(1u << (i & 31));
      while (b) {
        const int z = top_bit(b);
        b &= bits_below(z);
        const int k = 32 * w + z;
        const uint32_t l8 = LiveMap::off8(v.smap, k, v.mstride);
        const uint2 q = W1 > 2 ? with_table<true>(__ldg(qrow_g + k), v.ztab) : lds64(qrow + (uint32_t)k * 8u);
        const uint2 ef = lds64(frow + l8), et = lds64(trow + l8);
        d += gated(q, et) - gated(q, ef);
      }
    }
    return d;
  }

  // One Metropolis move (kernel.cu:1032-1191) on query SSE i = pick_index(first draw of the move).  u2/u3 are callables
  // returning 32 random bits, so that a sequential generator is advanced exactly when the reference would draw (u2 only
  // with >= 2 candidates).
  template <class U2, class U3>
  __device__ __forceinline__ void move(const TeamView &v, int m, const SatsKParams &p, int &best, int &best_tag, int tag,
                                       const int i, U2 &&u2, U3 &&u3)
  {
    const bool was_mapped = bit_test<W1>(mq, i);
    int lo, hi, from;
    if (LORDER) {
      int kp = top_at_or_below<W1>(mq, i);
      // word maps keep n2 in "element -1" (written once per entry), so a missing lower neighbour needs no select
      lo = (W1 <= 2) ? LiveMap::get(v.smap, kp, v.mstride) : (kp >= 0 ? LiveMap::get(v.smap, kp, v.mstride) : v.n2);
      from = was_mapped ? lo : -1;
      if (W1 <= 2) {
        // word maps: a sentinel bit "SSE n1" above the last one (their masks always have room for it) whose map element
        // holds 0 -- an upper bound that empties the window like the reference's -1 --, so the upper neighbour always
        // exists and the empty case needs no select
        uint32_t ms[W1];
#pragma unroll
        for (int w = 0; w < W1; w++) ms[w] = mq[w] | bit_in_word(v.n1, w);
        hi = LiveMap::get(v.smap, low_at_or_above<W1>(ms, i + 1), v.mstride);
      } else {
        const int kn = low_at_or_above<W1>(mq, i + 1);            // -1 for the last query SSE: nothing above it
        hi = kn >= 0 ? LiveMap::get(v.smap, kn, v.mstride) : -1;
      }
      if (i == v.n1 - 1) hi = v.n2;
    } else {
      lo = 0; hi = v.n2;
      from = was_mapped ? LiveMap::get(v.smap, i, v.mstride) : -1;
    }
    // randtypeind (kernel.cu:677-714): unoccupied entry SSEs of the right type inside [lo, hi)
    uint32_t cand[W2];
    int ncand = 0;
#pragma unroll
    for (int w = 0; w < W2; w++) {
      // one entry word and word maps: lo and hi are never negative (the map's sentinel elements hold n2 and 0), so the range
      // masks need no clamping
      if (W2 == 1 && W1 <= 2) cand[w] = lds32(v.qmask + (uint32_t)i * 4u) & ~md[0] & below_nonneg(hi) & ~below_nonneg(lo);
      else cand[w] = lds32(v.qmask + (uint32_t)(i * W2 + w) * 4u) & ~md[w] & below(hi - 32 * w) & ~below(lo - 32 * w);
      ncand += __popc(cand[w]);
    }
    int to = -1;
    if (ncand == 1) to = select_nth<W2>(cand, 0);
    else if (ncand > 1) to = select_nth<W2>(cand, scaled_index(unit_from_bits(u2()), ncand));

    int d = 0;
    if (from >= 0 || to >= 0) d = delta(v, i, from, to);
    const int cand_score = score + d;
    const bool improved = cand_score > best;
    if (improved) {
      best = cand_score;
      best_tag = tag;
      if (LSOLN) best_is_live();
    }
    // Metropolis: expf((float)d / T_m) > unit(x) (kernel.cu:1166) as x < cut[m][-d], cut tabulated on the host from the
    // reference's exact fp32 thresholds (d > 0 always passes: the threshold exceeds 1 >= unit(x))
    const uint32_t x = u3();
    bool accept;
    if (XORWOW && p.accept_mode == SATS_ACCEPT_DEVICE_FAST) {
      accept = __expf(__fdividef((float)d, __ldg(p.temps + m))) > unit_from_bits(x);
    } else if (d > 0) {
      accept = true;
    } else if (d == 0) {
      accept = x < p.accept_cut0;
    } else {
      int nd = -d;
      accept = nd <= SATS_K_DCLAMP && x < __ldg(p.accept_cut + m * (SATS_K_DCLAMP + 1) + nd);
    }
    if (accept) {
      score = cand_score;
      if (W2 == 1) {
        md[0] = (md[0] & ~bit_in_word(from, 0)) | bit_in_word(to, 0);       // shl.b32 clamps: "bit -1" is no bit, no selects
      } else {
        if (from >= 0) bit_clear<W2>(md, from);
        if (to >= 0) bit_set<W2>(md, to);
      }
      if (to >= 0) bit_set<W1>(mq, i);
      else bit_clear<W1>(mq, i);
      if (from >= 0 || to >= 0) {
        if (LSOLN && !improved && owned && !bit_test<W1>(dirty, i)) {
          Map<false>::put(v.bmap, i, v.mstride, from);       // what SSE i was matched to when the best score was reached
          bit_set<W1>(dirty, i);
        }
        LiveMap::put(v.smap, i, v.mstride, to);
      }
    }
  }
};

// Runs every chain this thread owns for one (query, entry) pair, then the team arg-max and the output.
// `red` is this entry's arg-max scratch (the callers alternate between two, so that the one team barrier per entry is
// enough); `before_barrier` runs on every thread after its chains and before that barrier, `after_barrier` right after it
// (from then on nobody reads the entry blob any more: the persistent loop claims and fetches the next entry there).
template <int W1, int W2, bool LORDER, bool XORWOW, bool LSOLN, class Before, class After>
__device__ __forceinline__ void anneal_entry(const SatsKParams &p, const TeamView &v, int team, int tl, uint64_t *red,
                                             uint32_t entry_orig, uint32_t query_index, Xorwow &xw,
                                             int out_slot, int entry_sorted, Before &&before_barrier, After &&after_barrier)
{
  // per query SSE, the entry SSEs of its type: every warp keeps its own copy, so a warp barrier is all it takes
  for (int k = tl & 31; k < v.n1; k += 32) {
    const uint32_t *tm = v.tmask + 4 * v.qtype[k];
#pragma unroll
    for (int w = 0; w < W2; w++) asm volatile("st.shared.b32 [%0], %1;" ::"r"(v.qmask + (uint32_t)(k * W2 + w) * 4u), "r"(tm[w]) : "memory");
  }
  if (W1 <= 2) {                                                   // see Chain::move: the window bounds of "nothing mapped below / above"
    Map<true>::put(v.smap, -1, v.mstride, v.n2);
    Map<true>::put(v.smap, v.n1, v.mstride, 0);        // "nothing mapped above": upper bound 0 = empty window (K:1064-1077 gives -1)
  }
  __syncwarp();
  Chain<W1, W2, LORDER, XORWOW, LSOLN> ch;
  int best = SATS_K_NEG_INIT;
  int best_tag = tl;                                  // XORWOW: thread id; Philox: restart index of the best chain
  const int chains = XORWOW ? ((p.restarts + p.tw - 1) / p.tw) * p.tw : p.restarts;

  for (int r = tl; r < chains; r += p.tw) {
    const int tag = XORWOW ? tl : r;
    if (XORWOW) {
      ch.seed(v, [&](int) { return xw.next() < p.seed_cut; });
    } else {
      uint32_t att[4];
      philox4x32_10(0x80000000u, (uint32_t)r, entry_orig, query_index, p.rk, att);
      ch.seed_bits(v, att);
    }
    ch.score = ch.full_score(v);
    if (ch.score > best) {
      best = ch.score;
      best_tag = tag;
      if (LSOLN) ch.best_is_live();
    } else {
      ch.owned = false;
    }
    if (XORWOW) {
      for (int m = 0; m < SATS_K_MOVES; m++)
        ch.move(v, m, p, best, best_tag, tag, pick_index(xw.next(), v.n1, v.pick_cut), [&] { return xw.next(); }, [&] { return xw.next(); });
    } else {
      // static draw positions: Philox block g feeds moves 2g (words 0, 1) and 2g + 1 (words 2, 3); per move the first word
      // picks the SSE (and, hashed, the candidate), the second is the Metropolis draw
      for (int g = 0; g < SATS_K_MOVES / 4; g++) {
        uint32_t a[4], b[4];
        philox4x32_10(2u * g + 0u, (uint32_t)r, entry_orig, query_index, p.rk, a);
        philox4x32_10(2u * g + 1u, (uint32_t)r, entry_orig, query_index, p.rk, b);
        const int m = 4 * g;
        // the SSE picks depend on the draws only, not on the chains' state: issued together, off the moves' critical path
        const int i0 = pick_index(a[0], v.n1, v.pick_cut), i1 = pick_index(a[2], v.n1, v.pick_cut);
        const int i2 = pick_index(b[0], v.n1, v.pick_cut), i3 = pick_index(b[2], v.n1, v.pick_cut);
        ch.move(v, m + 0, p, best, best_tag, tag, i0, [&] { return candidate_bits(a[0]); }, [&] { return a[1]; });
        ch.move(v, m + 1, p, best, best_tag, tag, i1, [&] { return candidate_bits(a[2]); }, [&] { return a[3]; });
        ch.move(v, m + 2, p, best, best_tag, tag, i2, [&] { return candidate_bits(b[0]); }, [&] { return b[1]; });
        ch.move(v, m + 3, p, best, best_tag, tag, i3, [&] { return candidate_bits(b[2]); }, [&] { return b[3]; });
      }
    }
    if (LSOLN) ch.finish_best_map(v);
  }

  // ---- arg-max over the team: highest score, lowest tag (kernel.cu:1205-1221 scans thread 0..127 with '>')
  const unsigned full = 0xffffffffu;
  int wbest = __reduce_max_sync(full, best);
  unsigned wtag = __reduce_min_sync(full, best == wbest ? (unsigned)best_tag : 0xffffffffu);
  const int warp_in_team = tl >> 5, warps = p.tw >> 5;
  if ((tl & 31) == 0) red[warp_in_team] = ((uint64_t)(uint32_t)(wbest + 0x40000000) << 32) | (uint32_t)(~wtag);
  before_barrier();
  team_sync(team, p.tw);
  after_barrier();
  uint64_t key = red[0];
  for (int w = 1; w < warps; w++) key = red[w] > key ? red[w] : key;
  const int team_best = (int)(uint32_t)(key >> 32) - 0x40000000;
  const unsigned team_tag = ~(uint32_t)key;
  if (tl == 0) {
    p.out_scores[(size_t)out_slot * p.out_stride + entry_sorted] = team_best;
    if (p.hit_thr != nullptr && team_best >= __ldg(p.hit_thr + out_slot * (SATS_MAXDIM_EXT + 1) + v.n2)) {
      const unsigned pos = atomicAdd(p.hit_cursor, 1u);
      if (pos < p.hit_cap) p.hit_list[pos] = make_int2(entry_sorted, (int)(((uint32_t)team_best << 16) | (uint32_t)out_slot));
    }
  }
  if (LSOLN && best == team_best && (unsigned)best_tag == team_tag) {
    int8_t *row = p.out_maps + ((size_t)out_slot * p.out_stride + entry_sorted) * SATS_K_MAPROW;
#pragma unroll 1
    for (int w = 0; w < Map<false>::words(v.n1); w++) {
      uint32_t x;
      asm("ld.shared.b32 %0, [%1];" : "=r"(x) : "r"(v.bmap + (uint32_t)w * v.mstride) : "memory");
      reinterpret_cast<uint32_t *>(row)[w] = x;
    }
  }
}

}  // namespace satsk

// Shared-memory layout of a CTA:
//   [0, 128)                    mbarriers: one for the query, one per team
//   then 256 B                  the 128-byte zeta table at the first 128-byte aligned address
//   then sm_query_bytes         query blob (header + SSE types only when W1 == 4)
//   then per team: entry blob (sm_entry_bytes) | live maps (mapwords*tw*4) | best maps (bmapwords*tw*4) | 80 B scratch (2 arg-max buffers, 2 claim slots) | one qmask (n1 x W2 words) per warp
template <int W1, int W2, bool LORDER, bool XORWOW, bool LSOLN>
__global__ void __launch_bounds__(SATS_K_MAXTHREADS, SATS_K_MINBLOCKS) sats_anneal_kernel(const SatsKParams p)
{
  using namespace satsk;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem);       // bar[0]: query (and XORWOW entries); bar[1 + team]: that team's entries
  const uint32_t zpad = (0u - (smem_u32(smem) + SATS_K_BAR_BYTES)) & 127u;       // up to the next 128-byte aligned address
  int8_t *sz = reinterpret_cast<int8_t *>(smem + SATS_K_BAR_BYTES + zpad);
  uint8_t *sq = smem + SATS_K_BAR_BYTES + SATS_K_ZTAB_BYTES;
  const int team = threadIdx.x / p.tw, tl = threadIdx.x - team * p.tw;
  uint8_t *steam = smem + SATS_K_BAR_BYTES + SATS_K_ZTAB_BYTES + p.sm_query_bytes + (size_t)team * p.sm_team_bytes;
  uint8_t *se = steam;
  uint8_t *smaps = se + p.sm_entry_bytes;
  uint8_t *bmaps = smaps + p.sm_mapwords * p.tw * 4;
  uint64_t *red = reinterpret_cast<uint64_t *>(bmaps + p.sm_bmapwords * p.tw * 4);

  const int qi = p.q_first + blockIdx.y;     // query slot of this batch: selects the blob and the output row
  if (threadIdx.x == 0)
    for (int t = 0; t <= p.teams; t++) mbar_init(bar + t, 1);
  fill_zeta_table(sz, threadIdx.x, blockDim.x);
  __syncthreads();

  TeamView v;
  v.mstride = (uint32_t)p.tw * 4u;
  v.qtype = sq + 16;
  v.pick_cut = smem_u32(sq + SATS_K_QUERY_PICK);
  // Queries of more than 64 SSEs (W1 == 4) keep their n1 x n1 cells in global memory (read through L1 with ld.global.nc):
  // an 82 KB query copy per CTA would leave room for one CTA per SM.  Only header + SSE types are staged then.
  v.qcell_g = reinterpret_cast<const uint2 *>(p.qblobs + p.qblob_off[qi] + SATS_K_QUERY_HDR);
  v.qcell = smem_u32(sq + SATS_K_QUERY_HDR);
  v.ztab = smem_u32(sz);
  v.tmask = reinterpret_cast<const uint32_t *>(se + 16);
  const uint32_t ecell0 = smem_u32(se + SATS_K_ENTRY_HDR);      // the matrix starts one row (8 n2 bytes) further: set per entry
  v.smap = smem_u32(smaps + tl * 4) + (W1 <= 2 ? (uint32_t)p.tw * 4u : 0u);      // word maps: element -1 exists
  v.bmap = smem_u32(bmaps + tl * 4);
  // team scratch: red[2][4] (two alternating arg-max buffers) | claim[2] | per-warp qmask copies
  volatile int *claim = reinterpret_cast<volatile int *>(red + 8);
  v.qmask = smem_u32(reinterpret_cast<uint8_t *>(red) + SATS_K_SCRATCH_BYTES + (tl >> 5) * p.sm_qmask_bytes);
  Xorwow xw;

  if (!XORWOW) {
    // Persistent teams: the query is staged once per CTA; after that every team works on its own -- its leader claims
    // the next entry of this launch's (decreasing-order) list from a global counter, brings the blob in with one TMA
    // bulk copy on the team's own mbarrier, the team anneals it, and so on until the list is exhausted.  No CTA-wide
    // synchronisation after the prologue, so a slow entry holds up only its own team.
    if (threadIdx.x == 0) {
      const uint32_t qbytes = W1 > 2 ? (uint32_t)SATS_K_QUERY_HDR : p.qblob_bytes[qi];
      mbar_expect_tx(bar, qbytes);
      tma_load_1d(sq, p.qblobs + p.qblob_off[qi], qbytes, bar);
    }
    mbar_wait(bar, 0);
    const int32_t *qh = reinterpret_cast<const int32_t *>(sq);
    const int32_t *eh = reinterpret_cast<const int32_t *>(se);
    v.n1 = qh[0];
    stamp_query_cells<W1>(sq, v.n1, v.ztab);
    uint64_t *tbar = bar + 1 + team;
    int *counter = p.counters + blockIdx.y;
    // claim an entry (leader only): take the next list position, publish it in claim[slot]; fetch it if there is one
    auto claim_next = [&](int slot) {
      const int idx = atomicAdd(counter, 1);
      claim[slot] = idx;
      return idx;
    };
    auto fetch = [&](int idx) {
      if (idx < p.item_count) {
        const int e = p.item_first + idx;
        const uint32_t bytes = p.blob_bytes[e];
        mbar_expect_tx(tbar, bytes);
        tma_load_1d(se, p.blobs + p.blob_off[e], bytes, tbar);
      }
    };
    if (tl == 0) fetch(claim_next(0));
    team_sync(team, p.tw);
    uint32_t phase = 0;
    // One team barrier per entry (the arg-max).  The leader takes the next list position once its own chains are done,
    // before that barrier, so the barrier also publishes it; right after the barrier nobody needs the entry blob any more
    // and the leader starts the next TMA copy, which then overlaps the arg-max and the output.
    for (int par = 0;; par ^= 1) {
      const int idx = claim[par];
      if (idx >= p.item_count) break;
      mbar_wait(tbar, phase);
      phase ^= 1u;
      v.n2 = eh[0];
      v.ecell = ecell0 + 8u * (uint32_t)v.n2;
      anneal_entry<W1, W2, LORDER, false, LSOLN>(p, v, team, tl, red + 4 * par, (uint32_t)eh[1], p.q_index_base + (uint32_t)qh[1], xw, qi, p.item_first + idx,
                                                 [&] { if (tl == 0) claim_next(par ^ 1); },
                                                 [&] { if (tl == 0) fetch(claim[par ^ 1]); });
    }
  } else {
    // validation: this CTA is reference block b; one team of 128 threads; entries b, b+128, ... in pool order
    const int b = p.xw_blocks[blockIdx.x];
    uint32_t *st = p.xw_states + ((size_t)b * SATS_REF_GRID_THREADS + tl) * 6;
    xw.load(st);
    uint32_t phase = 0;
    if (threadIdx.x == 0) {
      const uint32_t qbytes = W1 > 2 ? (uint32_t)SATS_K_QUERY_HDR : p.qblob_bytes[qi];
      mbar_expect_tx(bar, qbytes);
      tma_load_1d(sq, p.qblobs + p.qblob_off[qi], qbytes, bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
    const int32_t *qh = reinterpret_cast<const int32_t *>(sq);
    v.n1 = qh[0];
    stamp_query_cells<W1>(sq, v.n1, v.ztab);
    for (int pos = b; pos < p.pool_count; pos += SATS_REF_GRID_BLOCKS) {
      const int e = p.pool_list[pos];
      if (threadIdx.x == 0) {
        mbar_expect_tx(bar, p.blob_bytes[e]);
        tma_load_1d(se, p.blobs + p.blob_off[e], p.blob_bytes[e], bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1u;
      const int32_t *eh = reinterpret_cast<const int32_t *>(se);
      v.n2 = eh[0];
      v.ecell = ecell0 + 8u * (uint32_t)v.n2;
      anneal_entry<W1, W2, LORDER, true, LSOLN>(p, v, 0, tl, red, (uint32_t)eh[1], p.q_index_base + (uint32_t)qh[1], xw, qi, e, [] {}, [] {});
      __syncthreads();       // the entry buffer and the arg-max scratch are reused by the next entry
    }
    xw.store(st);
  }
}

#endif  // SATS_KERNEL_CUH
