/*
 * TESTING BUILDS ONLY.  Entry points that exist in libvlitefast_testing.so (the product sources compiled
 * with -DVF_TESTING) and not in libvlitefast.so: round 1's channelisers behind vf_config.k1_threads
 * (1: pipelined, two-for-one FFT; 320 / 512 / 640: monolithic) and the self-checks below.  tests/ load this library for A/B comparisons only.
 */
#ifndef VF_TESTING_H
#define VF_TESTING_H
#include "vlitefast.h"
#ifdef __cplusplus
extern "C" {
#endif
/* the normaliser divides with a packed, branch-free sequence; this runs it beside CUDA's correctly
 * rounded division on n (even) operand pairs p / b so that a test can compare the bits */
int vf_debug_division (vf_handle *h, const float *p, const float *b, float *q_packed, float *q_ref, size_t n);
/* clock64 stamps of the normaliser's chunk pipeline, first CTA: [2 streams][4096 chunks][6]:
 * 0 iteration top, 1 rows landed and prepared, 2 bandpass received, 3 recursion done, 4 handed on, 5 chunk finished.
 * First call (out == NULL) arms the trace, a later call copies it out. */
int vf_debug_k2_trace (vf_handle *h, long long *out);
#ifdef __cplusplus
}
#endif
#endif
