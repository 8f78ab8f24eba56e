/* Stand-in for the CUDA-samples helper_timer.h (absent from this image).
 * TEST INFRASTRUCTURE ONLY.  A monotonic-clock stopwatch exposing the six
 * sdk*Timer calls the reference host code uses; values are milliseconds. */
#ifndef SATS_STUB_HELPER_TIMER_H
#define SATS_STUB_HELPER_TIMER_H
#include <time.h>
#include <stdlib.h>
struct StopWatchInterface {
  struct timespec t0;
  double acc_ms;
  int running;
};
static inline double sats_stub_now_ms(const struct timespec *a) {
  struct timespec b;
  clock_gettime(CLOCK_MONOTONIC, &b);
  return (b.tv_sec - a->tv_sec) * 1e3 + (b.tv_nsec - a->tv_nsec) * 1e-6;
}
static inline bool sdkCreateTimer(StopWatchInterface **t) {
  *t = (StopWatchInterface *)calloc(1, sizeof(StopWatchInterface));
  return *t != NULL;
}
static inline bool sdkDeleteTimer(StopWatchInterface **t) {
  free(*t); *t = NULL; return true;
}
static inline bool sdkResetTimer(StopWatchInterface **t) {
  (*t)->acc_ms = 0.0;
  if ((*t)->running) clock_gettime(CLOCK_MONOTONIC, &(*t)->t0);
  return true;
}
static inline bool sdkStartTimer(StopWatchInterface **t) {
  clock_gettime(CLOCK_MONOTONIC, &(*t)->t0); (*t)->running = 1; return true;
}
static inline bool sdkStopTimer(StopWatchInterface **t) {
  if ((*t)->running) { (*t)->acc_ms += sats_stub_now_ms(&(*t)->t0); (*t)->running = 0; }
  return true;
}
static inline float sdkGetTimerValue(StopWatchInterface **t) {
  double v = (*t)->acc_ms;
  if ((*t)->running) v += sats_stub_now_ms(&(*t)->t0);
  return (float)v;
}
#endif
