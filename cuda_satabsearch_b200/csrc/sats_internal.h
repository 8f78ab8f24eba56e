// Internal definitions shared by the host (sats_host.cpp) and device (sats_device.cu) halves of libsats.
#ifndef SATS_INTERNAL_H
#define SATS_INTERNAL_H

#include <cstdarg>
#include <cstdint>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "sats.h"

// Host database: structures in original file order, each stored as the lower triangle (diagonal
// included, row-major: cell (i, j<=i) at i*(i+1)/2 + j) of its symmetric tableau and distance matrix.
struct sats_db {
  std::vector<int32_t> order;
  std::vector<char> names;        // count x 9, NUL padded
  std::vector<int64_t> tri_off;   // count + 1 prefix sums of order*(order+1)/2
  std::vector<uint8_t> tab;
  std::vector<float> dmat;

  int count() const { return (int)order.size(); }
  const char *name(int i) const { return names.data() + (size_t)i * 9; }
  static inline int64_t tri(int i, int j) { return i >= j ? (int64_t)i * (i + 1) / 2 + j : (int64_t)j * (j + 1) / 2 + i; }
  uint8_t code(int e, int i, int j) const { return tab[tri_off[e] + tri(i, j)]; }
  float dist(int e, int i, int j) const { return dmat[tri_off[e] + tri(i, j)]; }
  void append(const char *nm, int n, const uint8_t *tri_tab, const float *tri_dmat);
};

int sats_fail(int status, const char *fmt, ...) __attribute__((format(printf, 2, 3)));

// C++ exceptions must not cross the C ABI: every entry point that allocates is a function-try-block ending in this
#define SATS_CATCH_ALL                                                                                         \
  catch (const std::bad_alloc &) { return sats_fail(SATS_ERR_NOMEM, "out of memory"); }                        \
  catch (const std::exception &ex) { return sats_fail(SATS_ERR_ARG, "internal error: %s", ex.what()); }

// Cost model used by the partitioner, sats_partition(): greedy longest-processing-time-first over the size-sorted list
// (SURVEY 8e shape a + b*n2, recalibrated on a B200: per-entry kernel time by size
// bucket in profiles/r01h_launches.txt is 0.107 us at order ~4 rising linearly to 0.43 us at order ~56, i.e.
// proportional to 13 + order for the bench query; the shape is what matters for balancing).
static inline double sats_entry_cost(int order) { return 13.0 + (double)order; }

#endif
