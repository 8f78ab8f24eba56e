// sats_kernels.cu -- instantiations of sats_anneal_kernel for one query mask width (compile with -DSATS_W1=1|2|4;
// three objects, built in parallel).  sats_device.cu picks a kernel through sats_pick_kernel_w<W1>().
#include "sats_kernel.cuh"

#ifndef SATS_W1
#error "compile with -DSATS_W1=1, 2 or 4"
#endif

namespace {
template <int W2, bool LORDER, bool XORWOW> sats_kernel_fn pick_lsoln(bool lsoln)
{
  return lsoln ? sats_anneal_kernel<SATS_W1, W2, LORDER, XORWOW, true> : sats_anneal_kernel<SATS_W1, W2, LORDER, XORWOW, false>;
}
template <int W2> sats_kernel_fn pick_flags(bool lorder, bool xorwow, bool lsoln)
{
  if (xorwow) return lorder ? pick_lsoln<W2, true, true>(lsoln) : pick_lsoln<W2, false, true>(lsoln);
  return lorder ? pick_lsoln<W2, true, false>(lsoln) : pick_lsoln<W2, false, false>(lsoln);
}
}  // namespace

#define SATS_PICK_NAME2(w) sats_pick_kernel_w##w
#define SATS_PICK_NAME(w) SATS_PICK_NAME2(w)
sats_kernel_fn SATS_PICK_NAME(SATS_W1)(int w2, bool lorder, bool xorwow, bool lsoln)
{
  switch (w2) {
    case 1: return pick_flags<1>(lorder, xorwow, lsoln);
    case 2: return pick_flags<2>(lorder, xorwow, lsoln);
    default: return pick_flags<4>(lorder, xorwow, lsoln);
  }
}
