// sats_kparams.h -- launch parameters and layout constants shared by the kernel translation units
// (sats_kernels.cu, compiled once per query mask width) and the dispatcher (sats_device.cu).
#ifndef SATS_KPARAMS_H
#define SATS_KPARAMS_H

#include <cstdint>
#include <vector_types.h>

#define SATS_K_MOVES 100
#define SATS_K_DCLAMP 229          // expf(-d/T) <= 2^-33 (smallest uniform) for every d >= 229 and T <= 10
#define SATS_K_NEG_INIT (-99999)
#define SATS_K_ENTRY_HDR 80        // 16 B header + 4 types x 4 words of type masks; then one row of NaN cells, then the matrix
#define SATS_K_QUERY_PICK 128      // query header: 16 B (n1, position in the batch) + 112 B of SSE types, then ...
#define SATS_K_QUERY_HDR 576       // ... 112 words of SSE-pick boundaries (pick_index() in sats_kernel.cuh); then the cells
#define SATS_K_BAR_BYTES 128       // shared-memory header: 1 + teams mbarriers (teams <= 12)
#define SATS_K_ZTAB_BYTES 256      // shared-memory room for the 128-byte zeta table at a 128-byte aligned address
#define SATS_K_SCRATCH_BYTES 80     // per team: two alternating arg-max buffers (4 x 8 B each) and two claim slots
#define SATS_K_MAPROW 112          // bytes per (query, entry) row of the device map output

struct SatsKParams {
  // database (this GPU's shard), entries sorted by decreasing order
  const uint8_t *blobs;            // entry blobs, each 16-byte aligned
  const uint64_t *blob_off;        // byte offset of entry k
  const uint32_t *blob_bytes;      // size of entry k's blob (multiple of 16)
  // queries of this launch: blockIdx.y selects one
  const uint8_t *qblobs;
  const uint64_t *qblob_off;
  const uint32_t *qblob_bytes;
  int q_first;                     // index into qblob_* of blockIdx.y == 0
  // work: Philox -> entries [item_first, item_first + item_count) of the sorted list, claimed one at a time by the
  //       teams of all CTAs through counters[blockIdx.y] (zeroed before the launch)
  //       XORWOW -> CTA c is reference block xw_blocks[c]; it walks pool_list[b], pool_list[b+128], ...
  int item_first, item_count;
  int *counters;
  const int32_t *pool_list;        // XORWOW: pool position -> sorted entry index
  int pool_count;
  const int32_t *xw_blocks;
  uint32_t *xw_states;             // 16384 x 6 words (d, v0..v4)
  // geometry / shared-memory carve-up (bytes)
  int tw;                          // threads per team
  int teams;                       // teams per CTA
  int sm_query_bytes;              // room for the largest query blob of this launch
  int sm_entry_bytes;              // room for the largest entry blob of this launch
  int sm_mapwords;                 // 32-bit words per live chain map: n1max (queries of <= 64 SSEs) or ceil(n1max / 4)
  int sm_bmapwords;                // 32-bit words per best map: ceil(n1max / 4) with lsoln, else 0
  int sm_qmask_bytes;              // per warp: room for n1max x W2 words (16-byte multiple)
  int sm_team_bytes;               // total per team
  // search parameters
  int restarts, lsoln, accept_mode;
  uint32_t rk[20];                 // Philox round keys: rk[2r] = seed_lo + r * 0x9E3779B9, rk[2r + 1] = seed_hi + r * 0xBB67AE85
  const uint32_t *accept_cut;      // [SATS_K_MOVES][SATS_K_DCLAMP + 1]: accept a move of score change -d at step m iff draw < cut
  uint32_t accept_cut0;            // the cut-off for d == 0 (unit(x) < 1.0f), the same at every step
  uint32_t seed_cut;               // validation streams: the seeding pass attempts a match iff draw < seed_cut (unit(x) < 0.5)
  uint32_t q_index_base;           // Philox query index = q_index_base + the query's position in the batch (header word 1)
  const float *temps;              // [SATS_K_MOVES]: T_m = 10 * 0.95^m accumulated in fp32 like kernel.cu:1189
  // streaming hits: while a significance cut is bound, the arg-max epilogue also appends every entry whose
  // score reaches hit_thr[query slot][entry order] to one device list, so that only the hits ever travel to the host
  const int32_t *hit_thr;          // [query slot][SATS_MAXDIM_EXT + 1], nullptr = no cut bound
  unsigned *hit_cursor;            // hits appended so far (may run past hit_cap: the overflow is detected on the host)
  int2 *hit_list;                  // (sorted entry index, score << 16 | query slot): |score| <= 12210 and slot < 65536 (gridDim.y)
  unsigned hit_cap;
  // outputs, indexed [query slot][sorted entry index]
  int32_t *out_scores;
  int8_t *out_maps;                // rows of SATS_K_MAPROW bytes, or nullptr
  int out_stride;                  // entries per query slot
};

#ifndef SATS_K_MAXTHREADS
#define SATS_K_MAXTHREADS 384
#define SATS_K_MINBLOCKS 3
#endif

typedef void (*sats_kernel_fn)(const SatsKParams);
// one definition per query mask width W1 (32-bit words: 1, 2 or 4), each in its own translation unit
sats_kernel_fn sats_pick_kernel_w1(int w2, bool lorder, bool xorwow, bool lsoln);
sats_kernel_fn sats_pick_kernel_w2(int w2, bool lorder, bool xorwow, bool lsoln);
sats_kernel_fn sats_pick_kernel_w4(int w2, bool lorder, bool xorwow, bool lsoln);

#endif  // SATS_KPARAMS_H
