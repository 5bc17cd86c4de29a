/*
 * 12500-point complex FFT in shared memory, used as a two-for-one real FFT:
 * z[n] = pol0[n] + i pol1[n]  ->  Z[k];  X0[k] = (Z[k] + conj Z[N-k]) / 2,
 * X1[k] = (Z[k] - conj Z[N-k]) / 2i.  It replaces the reference's batched
 * cuFFT R2C (cufftPlan1d (NFFT, CUFFT_R2C, 2048), src/process_baseband.cu:597-598,
 * exec :1222-1224): same unnormalised forward transform, both polarisations
 * of one time step in a single pass, and only the channels the filterbank keeps
 * (CHANMIN..CHANMAX, src/process_baseband.h:53-54) ever leave the SM.
 *
 * N = 12500 = 25 * 25 * 20, three Stockham autosort passes with the radix
 * butterflies held in registers:
 *
 *   pass A  radix 25, 500 butterflies  in  x[p + 500 j]        (8-bit samples)
 *                                      out y[25 p + k] * w_N^(p k)
 *   pass B  radix 25, 500 butterflies  in  x[i + 500 j],  i = q + 25 p
 *                                      out y[q + 625 p + 25 k] * w_500^(p k)
 *   pass C  radix 20, 625 butterflies  in  x[q + 625 j]
 *                                      out Z[q + 625 k]        (natural order)
 *
 * Every function here is __host__ __device__ and takes the butterfly index
 * explicitly, so tests/fft_hosttest.cu can run the exact index arithmetic of
 * the kernel on the CPU (one loop per pass where the kernel has a barrier).
 */
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#if defined(__CUDACC__)
#define VF_HD __host__ __device__ __forceinline__
#else
#define VF_HD inline
#endif

/* compiler-only fence: keeps ptxas from hoisting every twiddle load of a
 * butterfly above its arithmetic (register pressure at 640 threads per SM) */
#if defined(__CUDA_ARCH__)
#define VF_SCHED_FENCE() asm volatile ("" ::: "memory")
#else
#define VF_SCHED_FENCE() do { } while (0)
#endif

#define VF_NFFT      12500
#define VF_NA        500     /* butterflies in passes A and B */
#define VF_NC        625     /* butterflies in pass C */

/* v *= (wr + i wi), compile-time constant */
#define VF_CMULC(v, wr, wi) do { float _x = (v).x, _y = (v).y; \
  (v).x = fmaf (_x, (wr), -(_y * (wi))); (v).y = fmaf (_x, (wi), _y * (wr)); } while (0)

#include "vf_fft_consts.h"

VF_HD float2 vf_cmul (float2 a, float2 b)
{
  return make_float2 (fmaf (a.x, b.x, -(a.y * b.y)), fmaf (a.x, b.y, a.y * b.x));
}

/* forward radix-5 butterfly, in place: a_k <- sum_j a_j exp(-2 pi i j k / 5) */
VF_HD void vf_r5 (float2 &a0, float2 &a1, float2 &a2, float2 &a3, float2 &a4)
{
  const float c1 = 0.30901699437494745f, c2 = -0.80901699437494745f;
  const float s1 = 0.95105651629515353f, s2 = 0.58778525229247314f;
  float2 t1 = make_float2 (a1.x + a4.x, a1.y + a4.y);
  float2 t2 = make_float2 (a2.x + a3.x, a2.y + a3.y);
  float2 t3 = make_float2 (a1.x - a4.x, a1.y - a4.y);
  float2 t4 = make_float2 (a2.x - a3.x, a2.y - a3.y);
  float2 m1 = make_float2 (fmaf (c2, t2.x, fmaf (c1, t1.x, a0.x)), fmaf (c2, t2.y, fmaf (c1, t1.y, a0.y)));
  float2 m2 = make_float2 (fmaf (c1, t2.x, fmaf (c2, t1.x, a0.x)), fmaf (c1, t2.y, fmaf (c2, t1.y, a0.y)));
  float2 u1 = make_float2 (fmaf (s2, t4.x, s1 * t3.x), fmaf (s2, t4.y, s1 * t3.y));
  float2 u2 = make_float2 (fmaf (-s1, t4.x, s2 * t3.x), fmaf (-s1, t4.y, s2 * t3.y));
  a0 = make_float2 (a0.x + t1.x + t2.x, a0.y + t1.y + t2.y);
  a1 = make_float2 (m1.x + u1.y, m1.y - u1.x);
  a4 = make_float2 (m1.x - u1.y, m1.y + u1.x);
  a2 = make_float2 (m2.x + u2.y, m2.y - u2.x);
  a3 = make_float2 (m2.x - u2.y, m2.y + u2.x);
}

/* forward radix-4 butterfly, in place */
VF_HD void vf_r4 (float2 &a0, float2 &a1, float2 &a2, float2 &a3)
{
  float2 e0 = make_float2 (a0.x + a2.x, a0.y + a2.y), e1 = make_float2 (a0.x - a2.x, a0.y - a2.y);
  float2 o0 = make_float2 (a1.x + a3.x, a1.y + a3.y), o1 = make_float2 (a1.x - a3.x, a1.y - a3.y);
  a0 = make_float2 (e0.x + o0.x, e0.y + o0.y);
  a2 = make_float2 (e0.x - o0.x, e0.y - o0.y);
  a1 = make_float2 (e1.x + o1.y, e1.y - o1.x);
  a3 = make_float2 (e1.x - o1.y, e1.y + o1.x);
}

/* 25-point DFT of v[j], j = a + 5 b, in two halves so that callers can
 * consume one output row at a time: after vf_dft25_cols + vf_dft25_row (v, d)
 * X[5 c + d] sits in v[c + 5 d]. */
VF_HD void vf_dft25_cols (float2 (&v)[25])
{
#pragma unroll
  for (int a = 0; a < 5; ++a) vf_r5 (v[a], v[a + 5], v[a + 10], v[a + 15], v[a + 20]);
  VF_TW25 (v);                       /* v[a + 5 d] *= W_25^(a d) */
}
#define vf_dft25_row(v, d) vf_r5 ((v)[5 * (d)], (v)[5 * (d) + 1], (v)[5 * (d) + 2], (v)[5 * (d) + 3], (v)[5 * (d) + 4])

/* 20-point DFT of v[j], j = a + 4 b.  On return X[5 c + d] sits in v[c + 4 d]. */
VF_HD void vf_dft20 (float2 (&v)[20])
{
#pragma unroll
  for (int a = 0; a < 4; ++a) vf_r5 (v[a], v[a + 4], v[a + 8], v[a + 12], v[a + 16]);
  VF_TW20 (v);                       /* v[a + 4 d] *= W_20^(a d) */
#pragma unroll
  for (int d = 0; d < 5; ++d) vf_r4 (v[4 * d], v[4 * d + 1], v[4 * d + 2], v[4 * d + 3]);
}

/* 8-bit sample -> voltage, src/pb_kernels.cu:23-33: 0 -> 0, else u/128 - 1.
 * 2^23 + u is built in the mantissa (no I2F); (2^23 + u)/128 - 65537 is exact. */
VF_HD float vf_unpack (unsigned u)
{
#if defined(__CUDA_ARCH__)
  float v = __uint_as_float (0x4B000000u | u);
#else
  float v = 8388608.0f + (float) u;
#endif
  float x = fmaf (v, 0.0078125f, -65537.0f);
  return u ? x : 0.0f;
}

/* Twiddle tables (built on the host in double, rounded to float):
 *   tw1[p]  = w_12500^p        p < 500
 *   tw5[p]  = w_12500^(5 p)    p < 500
 *   tw500[m] = w_500^m         m < 500                                        */
struct vf_fft_tables {
  const float2 *tw1, *tw5, *tw500;
};

/* pass A: butterfly p in [0,500).  b0/b1 point at sample 0 of this FFT block
 * for pol 0 / pol 1.  Sub-block j of 500 samples (the kurtosis block,
 * src/pb_kernels.cu:243-295) is exactly input j of every butterfly, so the
 * excision mask is applied by dropping inputs. */
VF_HD void vf_pass_a (int p, const uint8_t *b0, const uint8_t *b1, uint32_t zero_mask,
                      const vf_fft_tables &tb, float2 *W)
{
  float2 v[25];
#pragma unroll
  for (int j = 0; j < 25; ++j) {
    float x0 = vf_unpack (b0[p + 500 * j]), x1 = vf_unpack (b1[p + 500 * j]);
    bool z = (zero_mask >> j) & 1u;
    v[j] = make_float2 (z ? 0.0f : x0, z ? 0.0f : x1);
  }
  vf_dft25_cols (v);
  /* external twiddle w_N^(p k), k = 5 c + d:  (w^5)^c * w^d */
  float2 w[5], q[5];
  w[1] = tb.tw1[p];
  w[2] = vf_cmul (w[1], w[1]);
  w[3] = vf_cmul (w[2], w[1]);
  w[4] = vf_cmul (w[2], w[2]);
  q[1] = tb.tw5[p];
  q[2] = vf_cmul (q[1], q[1]);
  q[3] = vf_cmul (q[2], q[1]);
  q[4] = vf_cmul (q[2], q[2]);
  float2 *o = W + 25 * p;
  vf_dft25_row (v, 0);
  o[0] = v[0];
#pragma unroll
  for (int c = 1; c < 5; ++c) o[5 * c] = vf_cmul (v[c], q[c]);
#pragma unroll
  for (int d = 1; d < 5; ++d) {
    VF_SCHED_FENCE ();
    vf_dft25_row (v, d);
    o[d] = vf_cmul (v[5 * d], w[d]);
#pragma unroll
    for (int c = 1; c < 5; ++c) o[5 * c + d] = vf_cmul (v[c + 5 * d], vf_cmul (q[c], w[d]));
  }
}

/* pass B, split at the barrier the in-place update needs */
VF_HD void vf_pass_b_load (int i, const float2 *W, float2 (&v)[25])
{
#pragma unroll
  for (int j = 0; j < 25; ++j) v[j] = W[i + 500 * j];
}

VF_HD void vf_pass_b_store (int i, float2 (&v)[25], const vf_fft_tables &tb, float2 *W)
{
  vf_dft25_cols (v);
  const int p = i / 25, q = i - 25 * p;
  float2 *o = W + q + 625 * p;
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    vf_dft25_row (v, d);
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      const int k = 5 * c + d;
      if (k) o[25 * k] = vf_cmul (v[c + 5 * d], tb.tw500[p * k]);
      else o[0] = v[0];
    }
    VF_SCHED_FENCE ();
  }
}

/* pass C */
VF_HD void vf_pass_c_load (int q, const float2 *W, float2 (&v)[20])
{
#pragma unroll
  for (int j = 0; j < 20; ++j) v[j] = W[q + 625 * j];
}

/* stores Z[q + 625 k] for the indices detection reads: [lo, hi] */
VF_HD void vf_pass_c_store (int q, float2 (&v)[20], float2 *W, int lo, int hi)
{
  vf_dft20 (v);
#pragma unroll
  for (int d = 0; d < 5; ++d)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int idx = q + 625 * (5 * c + d);
      if (idx >= lo && idx <= hi) W[idx] = v[c + 4 * d];
    }
}

/* detection of output channel ch (FFT bin k = ch + chanmin) for both pols:
 * |X0[k]|^2 and |X1[k]|^2 from Z[k], Z[N-k]. */
VF_HD float2 vf_detect (int k, const float2 *W)
{
  const float2 a = W[k], b = W[VF_NFFT - k];
  const float sr = a.x + b.x, si = a.y - b.y;     /* 2 X0 */
  const float dr = a.y + b.y, di = a.x - b.x;     /* 2 X1 = (dr, -di) */
  return make_float2 (0.25f * fmaf (sr, sr, si * si), 0.25f * fmaf (dr, dr, di * di));
}
