/* Stand-in for the CUDA-samples multithreading.h (absent from this image).
 * TEST INFRASTRUCTURE ONLY: the reference only needs the CUT_THREADPROC tag. */
#ifndef SATS_STUB_MULTITHREADING_H
#define SATS_STUB_MULTITHREADING_H
#define CUT_THREADPROC void
#endif
