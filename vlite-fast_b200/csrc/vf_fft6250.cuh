/*
 * Real 12500-point FFT of ONE polarisation as a 6250-point complex FFT in shared memory plus a split pass:
 *   z[m] = x[2m] + i x[2m+1]  ->  Z[k] (M = 6250 points),
 *   X[k] = (Z[k] + conj Z[M-k]) / 2  -  (i/2) w_N^k (Z[k] - conj Z[M-k]),   N = 12500, Z[M] = Z[0].
 * Same transform as the reference's cuFFT R2C (cufftPlan1d (NFFT, CUFFT_R2C, 2048), src/process_baseband.cu:597-598,
 * exec :1222-1224).  Against the two-for-one complex FFT of vf_fft12500.cuh (both polarisations in one 12500-point
 * transform) the arithmetic is about the same, but a transform needs half the shared memory, so the two polarisations
 * of a time step are transformed by two INDEPENDENT thread groups whose load, arithmetic, store and barrier phases
 * interleave on the SM's schedulers instead of marching in step.
 *
 * M = 6250 = 25 * 25 * 10, decimation in frequency, in place.  n = 250 j + 10 j' + p',  k = k1 + 25 k2 + 625 k3:
 *   pass 1  250 butterflies (p = 10 j' + p'), radix 25 over j
 *           in   sample byte pairs (x[2n], x[2n+1]), n = 250 j + p   (unpack fused, mask = drop input j: a 500-sample
 *                kurtosis block, src/pb_kernels.cu:243-295, is exactly input j of every butterfly)
 *           out  W[S k1 + p]  *=  w_M^(p k1)
 *   pass 2  250 butterflies (k1, p'), radix 25 over j':  W[S k1 + 10 j' + p'] -> W[S k1 + 10 k2 + p'] * w_250^(p' k2)
 *   pass 3  625 butterflies (k1, k2), radix 10 over p':  W[S k1 + 10 k2 + p'] -> W[S k1 + 10 k2 + k3]
 * so Z[k] ends at pos (k) = S (k mod 25) + 10 ((k / 25) mod 25) + k / 625.  S = 265: odd multiple structure as in
 * vf_fft12500.cuh (S mod 16 = 9, 25 S = 1 mod 16), threads that walk k1 fall on distinct bank pairs.
 *
 * Every function is __host__ __device__ with the butterfly index explicit: csrc/vf_fft6250_hosttest.cu runs them on the
 * CPU, one loop per barrier interval (tests/test_fft_host.py).
 */
#pragma once
#include "vf_fft12500.cuh"

#define VF6_M        6250
#define VF6_NA       250     /* butterflies in passes 1 and 2 */
#define VF6_NC       625     /* butterflies in pass 3 */
#define VF6_WS       265     /* padded block stride of W */
#define VF6_WLEN     (25 * VF6_WS)

/* Twiddle tables (host, double -> float):
 *   tw1[p] = w_M^p, tw5[p] = w_M^(5 p)           p < 250                 (pass 1, products as in vf_dft25_rows_store)
 *   tw250[(k2 - 1) * 10 + p'] = w_250^(p' k2)    p' < 10, 1 <= k2 < 25   (pass 2)
 *   tws[k] = w_N^k                               k < M                   (split pass)                               */
struct vf6_tables {
  const float2 *tw1, *tw5, *tw250, *tws;
};

VF_HD void vf_r2 (float2 &a0, float2 &a1)
{
  const float2 s = vf_add2 (a0, a1), d = vf_sub2 (a0, a1);
  a0 = s; a1 = d;
}

/* 10-point DFT of v[j], j = a + 2 b.  On return X[5 c + d] sits in v[c + 2 d]. */
VF_HD void vf_dft10 (float2 (&v)[10])
{
  vf_r5 (v[0], v[2], v[4], v[6], v[8]);
  vf_r5 (v[1], v[3], v[5], v[7], v[9]);
  VF_TW10 (v);                       /* v[1 + 2 d] *= W_10^d */
#pragma unroll
  for (int d = 0; d < 5; ++d) vf_r2 (v[2 * d], v[2 * d + 1]);
}

/* two consecutive samples of one polarisation as (re, im), at HALF the reference's scale: the bytes are sanitised
 * (0 -> 128), 2^23 + u is built in the mantissa and (2^23 + u) / 256 - 32768.5 = (u - 128) / 256 is exact
 * (src/pb_kernels.cu:23-33 has u / 128 - 1).  The transform is linear and scaling by a power of two is exact, so
 * every Z is exactly half of what the full scale gives, and the split pass's X = (s - i t) / 2 needs no further factor:
 * the powers are bit-identical to 0.25 |s - i t|^2 at full scale. */
VF_HD float2 vf6_unpack_pair (const uint8_t *b)
{
#if defined(__CUDA_ARCH__)
  const unsigned u = *reinterpret_cast<const unsigned short *> (b);
  const float2 v = make_float2 (__uint_as_float (__byte_perm (u, 0x4B000000u, 0x7540)), __uint_as_float (__byte_perm (u, 0x4B000000u, 0x7541)));
#else
  const float2 v = make_float2 (8388608.0f + (float) b[0], 8388608.0f + (float) b[1]);
#endif
  return vf_fma2 (v, vf_bc (0.00390625f), vf_bc (-32768.5f));
}

/* pass 1: butterfly p in [0,250).  b points at (sanitised) sample 0 of this FFT block of one polarisation (2-byte
 * aligned); bit j of zero_mask drops input j. */
template <bool MASKED>
VF_HD void vf6_pass1 (int p, const uint8_t *b, uint32_t zero_mask, const vf6_tables &tb, float2 *W)
{
  float2 v[25];
#pragma unroll
  for (int j = 0; j < 25; ++j) {
    if (MASKED && ((zero_mask >> j) & 1u)) v[j] = make_float2 (0.0f, 0.0f);
    else v[j] = vf6_unpack_pair (b + 2 * p + 500 * j);
  }
  vf_dft25_cols (v);
  vf_dft25_rows_store (v, tb.tw1[p], tb.tw5[p], W + p, VF6_WS);
}

/* second half of a radix-25 pass with one table look-up per output: tw[(k - 1) * 10 + pk] = w_250^(pk k) */
VF_HD void vf6_dft25_rows_store_tab (float2 (&v)[25], const float2 *tw, int pk, float2 *o)
{
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    vf_dft25_row (v, d);
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      const int k = 5 * c + d;
      if (k) o[k * 10] = vf_cmul (v[c + 5 * d], tw[(k - 1) * 10 + pk]);
      else o[0] = v[0];
    }
    VF_SCHED_FENCE ();
  }
}

/* pass 2: butterfly b in [0,250): offset p' = b / 25, block k1 = b % 25; in place */
VF_HD void vf6_pass2 (int b, const vf6_tables &tb, float2 *W)
{
  const int pp = b / 25, k1 = b - 25 * pp;
  float2 *o = W + VF6_WS * k1 + pp;
  float2 v[25];
#pragma unroll
  for (int j = 0; j < 25; ++j) v[j] = o[10 * j];
  vf_dft25_cols (v);
  vf6_dft25_rows_store_tab (v, tb.tw250, pp, o);
}

/* pass 3: butterfly m in [0,625): k1 = m % 25, k2 = m / 25; in place.  Every output is needed: the channels the filterbank
 * keeps are k = 2155..6250 and the split pass pairs them with M - k = 0..4095. */
VF_HD void vf6_pass3 (int m, float2 *W)
{
  const int k2 = m / 25, k1 = m - 25 * k2;
  float2 *o = W + VF6_WS * k1 + 10 * k2;
  float2 v[10];
#pragma unroll
  for (int j = 0; j < 10; ++j) v[j] = o[j];
  vf_dft10 (v);
#pragma unroll
  for (int d = 0; d < 5; ++d)
#pragma unroll
    for (int c = 0; c < 2; ++c) o[5 * c + d] = v[c + 2 * d];
}

/* location of Z[k] after pass 3, k < M */
VF_HD int vf6_zpos (int k)
{
  const int k3 = k / 625, r = k - 625 * k3, k2 = r / 25, k1 = r - 25 * k2;
  return VF6_WS * k1 + 10 * k2 + k3;
}

/* split pass: |X[k]|^2 from a = Z[k mod M], b = Z[(M - k) mod M] and w = w_N^k.
 *   s = a + conj b,  d = a - conj b,  t = w d,  2 X = s - i t = (s.x + t.y, s.y - t.x) */
VF_HD float vf6_split_power (float2 a, float2 b, float2 w)
{
  const float2 s = vf_add2 (a, make_float2 (b.x, -b.y));
  const float2 d = vf_add2 (a, make_float2 (-b.x, b.y));
  const float2 t = vf_cmul (d, w);
  const float2 u = vf_add2 (s, make_float2 (t.y, -t.x));
  return fmaf (u.x, u.x, u.y * u.y);             /* Z at half scale (vf6_unpack_pair): u = X */
}

/* the same with the field amplitude scaled by amp (FRB injection, src/pb_kernels.cu:38-66) */
VF_HD float vf6_split_power_amp (float2 a, float2 b, float2 w, float amp)
{
  const float2 s = vf_add2 (a, make_float2 (b.x, -b.y));
  const float2 d = vf_add2 (a, make_float2 (-b.x, b.y));
  const float2 t = vf_cmul (d, w);
  const float2 u = vf_add2 (s, make_float2 (t.y, -t.x));
  const float xr = u.x * amp, xi = u.y * amp;
  return fmaf (xr, xr, xi * xi);
}

VF_HD float vf6_detect (int c, const float2 *W, const float2 *tws)
{
  const int k = 2155 + c;
  const float2 w = (k < VF6_M) ? tws[k] : make_float2 (-1.0f, 0.0f);      /* w_N^M = -1 */
  return vf6_split_power (W[vf6_zpos (k % VF6_M)], W[vf6_zpos ((VF6_M - k) % VF6_M)], w);
}

/* ---- pass 3 fused with the split pass -------------------------------------------------------------------------------
 * Z[k] and Z[M - k] come out of two pass-3 butterflies that are each other's mirror image: with k = kb + 625 k3
 * (kb = k1 + 25 k2 the butterfly, k3 its output), M - k = (625 - kb) + 625 (9 - k3) for kb > 0.  A thread that
 * transforms butterflies kb and 625 - kb therefore holds, in registers, every pair the split pass needs -- the 20
 * outputs never go back to shared memory (saves the 6250 stores of pass 3, the 8192 loads of a separate split pass
 * and a barrier).  One pair (a = Z[k], b = Z[M - k], w = w_N^k) yields both mirror channels:
 *     s = a + conj b,  d = a - conj b,  t = w d,   |2 X[k]|^2 = |s - i t|^2,   |2 X[M - k]|^2 = |s + i t|^2
 * (the transform runs at half scale, so the factor 2 is already in Z)
 * (w_N^(M - k) = -conj w_N^k turns the second into the conjugate of s + i t).
 * Kept channels are k = 2155..6250, channel index c = k - 2155: of a pair, X[k] is kept when k3 >= 4 or (k3 = 3 and
 * kb >= 280), X[M - k] (c = 4095 - k) when k3 <= 5 or (k3 = 6 and kb <= 345): compile-time except for two outputs.
 * Units: u = 1..312 is the pair (u, 625 - u); unit 0 is butterfly 0, which mirrors onto itself (k3 <-> 10 - k3,
 * Z[M] = Z[0]) and also yields X[M], the last channel.  out points at channel 0 of this polarisation in a tile of
 * (pol 0, pol 1) pairs: channel c at out[2 c].  tws[k] = w_N^k, k < M. */
#define VF6_NU       313     /* work units of the fused pass */
#define VF6_IDX(k3)  ((k3) / 5 + 2 * ((k3) % 5))      /* where vf_dft10 leaves output k3 */

VF_HD void vf6_mirror_powers (float2 a, float2 b, float2 w, float &pk, float &pm)
{
  const float2 s = vf_add2 (a, make_float2 (b.x, -b.y));
  const float2 d = vf_add2 (a, make_float2 (-b.x, b.y));
  const float2 t = vf_cmul (d, w);
  const float2 um = vf_add2 (s, make_float2 (t.y, -t.x));      /* s - i t */
  const float2 up = vf_add2 (s, make_float2 (-t.y, t.x));      /* s + i t */
  pk = fmaf (um.x, um.x, um.y * um.y);           /* Z at half scale (vf6_unpack_pair): um = X[k], up = conj X[M - k] */
  pm = fmaf (up.x, up.x, up.y * up.y);
}

VF_HD void vf6_pass3_split (int u, const float2 *W, const float2 *tws, float *out)
{
  /* unit 0 runs through the same code as the pairs: its partner is butterfly 0 itself with the outputs rotated by
   * one (Z[M - 625 k3] = Z[625 ((9 - k3) + 1)]), which is the transform of the inputs times W_10^j */
  const int kb = u, kc = u ? 625 - u : 0;
  const int k2 = kb / 25, k1 = kb - 25 * k2, k2c = kc / 25, k1c = kc - 25 * k2c;
  const float2 *o1 = W + VF6_WS * k1 + 10 * k2, *o2 = W + VF6_WS * k1c + 10 * k2c;
  float2 v1[10], v2[10];
#pragma unroll
  for (int j = 0; j < 10; ++j) { v1[j] = o1[j]; v2[j] = o2[j]; }
  if (u == 0) VF_ROT10 (v2);
  float *const ok = out + 2 * (kb - 2155), *const om = out + 2 * (4095 - kb);      /* channel of X[kb], of X[M - kb] */
  const float2 *const twk = tws + kb;
  vf_dft10 (v1);
  vf_dft10 (v2);
#pragma unroll
  for (int k3 = 0; k3 < 10; ++k3) {
    float pk, pm;
    vf6_mirror_powers (v1[VF6_IDX (k3)], v2[VF6_IDX (9 - k3)], twk[625 * k3], pk, pm);
    if (k3 >= 4 || (k3 == 3 && kb >= 280)) ok[2 * 625 * k3] = pk;
    /* unit 0: its mirror outputs are its own (written from the other side), except X[M] */
    if ((k3 <= 5 || (k3 == 6 && kb <= 345)) && (u > 0 || k3 == 0)) om[-2 * 625 * k3] = pm;
  }
}
