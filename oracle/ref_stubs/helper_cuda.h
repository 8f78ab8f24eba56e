/* Stand-in for the CUDA-samples helper_cuda.h, which this image does not ship.
 * TEST INFRASTRUCTURE ONLY: lets oracle/build_ref.sh compile the unmodified
 * reference sources (they include <helper_cuda.h> for checkCudaErrors). */
#ifndef SATS_STUB_HELPER_CUDA_H
#define SATS_STUB_HELPER_CUDA_H
#include <stdio.h>
#include <stdlib.h>
#include <cuda_runtime.h>
#define checkCudaErrors(call)                                                   \
  do {                                                                          \
    cudaError_t sats_stub_err_ = (call);                                        \
    if (sats_stub_err_ != cudaSuccess) {                                        \
      fprintf(stderr, "CUDA error %d (%s) at %s:%d\n", (int)sats_stub_err_,     \
              cudaGetErrorString(sats_stub_err_), __FILE__, __LINE__);          \
      exit(1);                                                                  \
    }                                                                           \
  } while (0)
#endif
