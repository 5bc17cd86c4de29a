/*
 * 12500-point complex FFT in shared memory, used as a two-for-one real FFT:
 * z[n] = pol0[n] + i pol1[n]  ->  Z[k];  X0[k] = (Z[k] + conj Z[N-k]) / 2,
 * X1[k] = (Z[k] - conj Z[N-k]) / 2i.  It replaces the reference's batched
 * cuFFT R2C (cufftPlan1d (NFFT, CUFFT_R2C, 2048), src/process_baseband.cu:597-598,
 * exec :1222-1224): same unnormalised forward transform, both polarisations
 * of one time step in a single pass, and only the channels the filterbank keeps
 * (CHANMIN..CHANMAX, src/process_baseband.h:53-54) ever leave the SM.
 *
 * N = 12500 = 25 * 25 * 20, decimation in frequency, IN PLACE: every butterfly
 * reads and writes the same shared-memory locations, so one barrier per pass
 * is enough.  With n = 500 j + 20 j' + p'  and  k = k1 + 25 k2 + 625 k3:
 *
 *   pass 1  500 butterflies (p = 20 j' + p'), radix 25 over j
 *           in   sample bytes x[500 j + p]          (unpack fused, mask = drop input j)
 *           out  W[S k1 + p]  *=  w_N^(p k1)
 *   pass 2  500 butterflies (k1, p'), radix 25 over j'
 *           in/out  W[S k1 + 20 j' + p']  ->  W[S k1 + 20 k2 + p'] * w_500^(p' k2)
 *   pass 3  625 butterflies (k1, k2), radix 20 over p'
 *           in/out  W[S k1 + 20 k2 + p']  ->  W[S k1 + 20 k2 + k3]
 *
 * so Z[k] ends at  pos (k) = S (k mod 25) + 20 ((k / 25) mod 25) + k / 625.
 * S = 505: the 500-element blocks are padded so that threads which walk k1
 * (passes 2 and 3, detection) fall on distinct shared-memory banks: a float2
 * occupies 2 of the 32 banks, S mod 16 = 9 is odd, and 25 S = 1 (mod 16), so in
 * pass 2 (thread b -> k1 = b mod 25, p' = b / 25) the step from (k1 = 24, p')
 * to (k1 = 0, p' + 1) continues the same progression of banks.
 * A 500-sample kurtosis block (src/pb_kernels.cu:243-295) is exactly input j of
 * every pass-1 butterfly, so excision is "drop input j".
 *
 * Every function is __host__ __device__ and takes the butterfly index
 * explicitly: csrc/vf_fft_hosttest.cu runs the same index arithmetic on the
 * CPU (one loop per pass where the kernel has a barrier).
 */
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#if defined(__CUDACC__)
#define VF_HD __host__ __device__ __forceinline__
#else
#define VF_HD inline
#endif

/* compiler-only fence: keeps ptxas from hoisting every load of a butterfly
 * above its arithmetic (register pressure) */
#if defined(__CUDA_ARCH__)
#define VF_SCHED_FENCE() asm volatile ("" ::: "memory")
#else
#define VF_SCHED_FENCE() do { } while (0)
#endif

#define VF_NFFT      12500
#define VF_NA        500     /* butterflies in passes 1 and 2 */
#define VF_NC        625     /* butterflies in pass 3 */
#define VF_WS        505     /* padded block stride of W (see below) */
#define VF_WLEN      (25 * VF_WS)

/* ---- packed fp32 arithmetic ---------------------------------------------- *
 * sm_100 has two-wide fp32 instructions (PTX add/mul/fma .f32x2, SASS FADD2 /
 * FMUL2 / FFMA2) whose operands take a register pair with an optional swap
 * (LO_HI), per-lane negation, or a broadcast scalar.  A complex number is such
 * a pair, so one instruction updates re and im together and the +-i rotations
 * and complex multiplies of the butterflies need no data movement: the FFT
 * issues about half the instructions of the scalar form.  Each lane is an
 * independent IEEE operation, so the results are bit-identical to the scalar
 * expressions of the host fallback (which the CPU emulation test runs). */
#if defined(__CUDA_ARCH__)
VF_HD float2 vf_add2 (float2 a, float2 b)
{
  float2 r;
  asm ("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
       : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
VF_HD float2 vf_mul2 (float2 a, float2 b)
{
  float2 r;
  asm ("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
       : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
VF_HD float2 vf_fma2 (float2 a, float2 b, float2 c)
{
  float2 r;
  asm ("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
       : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
#else
VF_HD float2 vf_add2 (float2 a, float2 b) { return make_float2 (a.x + b.x, a.y + b.y); }
VF_HD float2 vf_mul2 (float2 a, float2 b) { return make_float2 (a.x * b.x, a.y * b.y); }
VF_HD float2 vf_fma2 (float2 a, float2 b, float2 c) { return make_float2 (fmaf (a.x, b.x, c.x), fmaf (a.y, b.y, c.y)); }
#endif
VF_HD float2 vf_bc (float s) { return make_float2 (s, s); }
VF_HD float2 vf_sub2 (float2 a, float2 b) { return vf_add2 (a, make_float2 (-b.x, -b.y)); }
/* a + (-i) u  and  a + i u */
VF_HD float2 vf_add_mi (float2 a, float2 u) { return vf_add2 (a, make_float2 (u.y, -u.x)); }
VF_HD float2 vf_add_pi (float2 a, float2 u) { return vf_add2 (a, make_float2 (-u.y, u.x)); }

/* a b = (a.x b.x - a.y b.y, a.y b.x + a.x b.y) in exactly two packed instructions and no data movement:
 *   t = b.y * (a.y, a.x)            FMUL2  t, b.y (scalar), a (halves swapped)
 *   r = b.x * (a.x, a.y) + (-t.x, t.y)   FFMA2  r, b.x (scalar), a, t (first half negated)
 * The operand modifiers of the packed instructions (swap of the halves, negation of one half, a scalar
 * register or an immediate broadcast to both halves) cover every step; written as
 * fma (bc (a.x), b, mul (bc (a.y), (-b.y, b.x))) the same product costs two more instructions (a MOV and
 * a negating FADD to build the rotated copy of b).  a is the data, b the twiddle. */
VF_HD float2 vf_cmul (float2 a, float2 b)
{
  const float2 t = vf_mul2 (vf_bc (b.y), make_float2 (a.y, a.x));
  return vf_fma2 (vf_bc (b.x), a, make_float2 (-t.x, t.y));
}

/* v *= (wr + i wi), compile-time constant: both factors are immediates of the two instructions
 *   t = wi * (-v.y, v.x);  v = wr * v + t */
#define VF_CMULC(v, wr, wi) do { (v) = vf_fma2 ((v), vf_bc (wr), vf_mul2 (make_float2 (-(v).y, (v).x), vf_bc (wi))); } while (0)

#include "vf_fft_consts.h"

/* forward radix-5 butterfly, in place: a_k <- sum_j a_j exp(-2 pi i j k / 5) */
VF_HD void vf_r5 (float2 &a0, float2 &a1, float2 &a2, float2 &a3, float2 &a4)
{
  const float c1 = 0.30901699437494745f, c2 = -0.80901699437494745f;
  const float s1 = 0.95105651629515353f, s2 = 0.58778525229247314f;
  const float2 t1 = vf_add2 (a1, a4), t2 = vf_add2 (a2, a3);
  const float2 t3 = vf_sub2 (a1, a4), t4 = vf_sub2 (a2, a3);
  const float2 m1 = vf_fma2 (vf_bc (c2), t2, vf_fma2 (vf_bc (c1), t1, a0));
  const float2 m2 = vf_fma2 (vf_bc (c1), t2, vf_fma2 (vf_bc (c2), t1, a0));
  const float2 u1 = vf_fma2 (vf_bc (s2), t4, vf_mul2 (vf_bc (s1), t3));
  const float2 u2 = vf_fma2 (vf_bc (-s1), t4, vf_mul2 (vf_bc (s2), t3));
  a0 = vf_add2 (vf_add2 (a0, t1), t2);
  a1 = vf_add_mi (m1, u1);
  a4 = vf_add_pi (m1, u1);
  a2 = vf_add_mi (m2, u2);
  a3 = vf_add_pi (m2, u2);
}

/* forward radix-4 butterfly, in place */
VF_HD void vf_r4 (float2 &a0, float2 &a1, float2 &a2, float2 &a3)
{
  const float2 e0 = vf_add2 (a0, a2), e1 = vf_sub2 (a0, a2);
  const float2 o0 = vf_add2 (a1, a3), o1 = vf_sub2 (a1, a3);
  a0 = vf_add2 (e0, o0);
  a2 = vf_sub2 (e0, o0);
  a1 = vf_add_mi (e1, o1);
  a3 = vf_add_pi (e1, o1);
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
 * The staged bytes are "sanitised" first (0 -> 128, which is the same voltage
 * 0.0), so the conversion is branch free: 2^23 + u is built in the mantissa
 * (no I2F) and (2^23 + u)/128 - 65537 is exact. */
VF_HD float vf_unpack_s (unsigned u)
{
#if defined(__CUDA_ARCH__)
  float v = __uint_as_float (0x4B000000u | u);
#else
  float v = 8388608.0f + (float) u;
#endif
  return fmaf (v, 0.0078125f, -65537.0f);
}

/* both polarisations of one sample in one packed FMA */
VF_HD float2 vf_unpack2_s (unsigned u0, unsigned u1)
{
#if defined(__CUDA_ARCH__)
  const float2 v = make_float2 (__uint_as_float (0x4B000000u | u0), __uint_as_float (0x4B000000u | u1));
#else
  const float2 v = make_float2 (8388608.0f + (float) u0, 8388608.0f + (float) u1);
#endif
  return vf_fma2 (v, vf_bc (0.0078125f), vf_bc (-65537.0f));
}

/* 0 -> 128 in every byte of a word */
VF_HD uint32_t vf_sanitise_word (uint32_t w)
{
  const uint32_t z = ~((((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w) | 0x7F7F7F7Fu);   /* 0x80 where the byte is 0 */
  return w | z;
}

/* Second half of a radix-25 pass: rows, twiddle by w^(5c+d) = q^c w^d
 * (w, q = w^5 from tables rounded once from double), store X[5c+d] to
 * o[(5c+d) * stride].  Products are formed at depth <= 3 multiplications. */
VF_HD void vf_dft25_rows_store (float2 (&v)[25], float2 w1, float2 q1, float2 *o, int stride)
{
  float2 w[5], q[5];
  w[1] = w1;
  w[2] = vf_cmul (w1, w1);
  w[3] = vf_cmul (w[2], w1);
  w[4] = vf_cmul (w[2], w[2]);
  q[1] = q1;
  q[2] = vf_cmul (q1, q1);
  q[3] = vf_cmul (q[2], q1);
  q[4] = vf_cmul (q[2], q[2]);
  vf_dft25_row (v, 0);
  o[0] = v[0];
#pragma unroll
  for (int c = 1; c < 5; ++c) o[5 * c * stride] = vf_cmul (v[c], q[c]);
#pragma unroll
  for (int d = 1; d < 5; ++d) {
    VF_SCHED_FENCE ();
    vf_dft25_row (v, d);
    o[d * stride] = vf_cmul (v[5 * d], w[d]);
#pragma unroll
    for (int c = 1; c < 5; ++c) o[(5 * c + d) * stride] = vf_cmul (v[c + 5 * d], vf_cmul (q[c], w[d]));
  }
}

/* Same with one table look-up per output: tw[(k - 1) * 20 + pk] = w_500^(pk k), laid out so that the
 * threads of a warp (consecutive pk) read consecutive entries */
VF_HD void vf_dft25_rows_store_tab (float2 (&v)[25], const float2 *tw, int pk, float2 *o, int stride)
{
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    vf_dft25_row (v, d);
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      const int k = 5 * c + d;
      if (k) o[k * stride] = vf_cmul (v[c + 5 * d], tw[(k - 1) * 20 + pk]);
      else o[0] = v[0];
    }
    VF_SCHED_FENCE ();
  }
}

/* Twiddle tables (built on the host in double, rounded to float):
 *   tw1[p] = w_12500^p, tw5[p] = w_12500^(5 p)      p < 500   (pass 1: 12000 distinct twiddles, formed as products)
 *   tw500[(k2 - 1) * 20 + p'] = w_500^(p' k2)       p' < 20, 1 <= k2 < 25   (pass 2)                              */
struct vf_fft_tables {
  const float2 *tw1, *tw5, *tw500;
};

/* pass 1: butterfly p in [0,500).  b0/b1 point at (sanitised) sample 0 of
 * this FFT block for pol 0 / pol 1; bit j of zero_mask drops input j. */
template <bool MASKED>
VF_HD void vf_pass1 (int p, const uint8_t *b0, const uint8_t *b1, uint32_t zero_mask,
                     const vf_fft_tables &tb, float2 *W)
{
  float2 v[25];
#pragma unroll
  for (int j = 0; j < 25; ++j) {
    if (MASKED && ((zero_mask >> j) & 1u)) v[j] = make_float2 (0.0f, 0.0f);
    else v[j] = vf_unpack2_s (b0[p + 500 * j], b1[p + 500 * j]);
  }
  vf_dft25_cols (v);
  vf_dft25_rows_store (v, tb.tw1[p], tb.tw5[p], W + p, VF_WS);
}

/* pass 2: butterfly b in [0,500): offset p' = b / 25, block k1 = b % 25; in place */
VF_HD void vf_pass2 (int b, const vf_fft_tables &tb, float2 *W)
{
  const int pp = b / 25, k1 = b - 25 * pp;
  float2 *o = W + VF_WS * k1 + pp;
  float2 v[25];
#pragma unroll
  for (int j = 0; j < 25; ++j) v[j] = o[20 * j];
  vf_dft25_cols (v);
  vf_dft25_rows_store_tab (v, tb.tw500, pp, o, 20);
}

/* pass 3: butterfly m in [0,625): k1 = m % 25, k2 = m / 25; in place.  Only
 * outputs k = kb + 625 k3 (kb = k1 + 25 k2 < 625) in [LO, HI] are stored.  For a given k3 that is decided at
 * compile time unless the range boundary falls inside [625 k3, 625 k3 + 624]: with the reference's channels
 * (LO = CHANMIN = 2155, HI = N - CHANMIN) twelve of the twenty stores are unconditional, two depend on kb and
 * six (with the additions that feed them) disappear. */
template <int LO, int HI>
VF_HD void vf_pass3 (int m, float2 *W)
{
  const int k2 = m / 25, k1 = m - 25 * k2;
  float2 *o = W + VF_WS * k1 + 20 * k2;
  float2 v[20];
#pragma unroll
  for (int j = 0; j < 20; ++j) v[j] = o[j];
  vf_dft20 (v);
  const int kb = m;                      /* k1 + 25 k2 with k2 = m / 25, k1 = m % 25 */
#pragma unroll
  for (int d = 0; d < 5; ++d)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int k3 = 5 * c + d, kmin = 625 * k3, kmax = 625 * k3 + 624;
      if (kmax < LO || kmin > HI) continue;                       /* never needed */
      if (kmin >= LO && kmax <= HI) o[k3] = v[c + 4 * d];         /* always needed */
      else if (kb + kmin >= LO && kb + kmin <= HI) o[k3] = v[c + 4 * d];
    }
}

/* location of Z[k] after pass 3 */
VF_HD int vf_zpos (int k)
{
  const int k3 = k / 625, r = k - 625 * k3, k2 = r / 25, k1 = r - 25 * k2;
  return VF_WS * k1 + 20 * k2 + k3;
}

/* detection of FFT bin k for both pols: |X0[k]|^2 and |X1[k]|^2 from Z[k], Z[N-k] */
VF_HD float2 vf_detect_pair (float2 a, float2 b)
{
  const float2 s = vf_add2 (a, make_float2 (b.x, -b.y));      /* 2 X0 = (a.x + b.x, a.y - b.y)             */
  const float2 d = vf_add2 (a, make_float2 (-b.x, b.y));      /* (a.x - b.x, a.y + b.y): 2 X1 = (d.y, -d.x) */
  /* (s.x^2 + s.y^2, d.y^2 + d.x^2) / 4 */
  return vf_mul2 (vf_bc (0.25f), vf_fma2 (make_float2 (s.x, d.y), make_float2 (s.x, d.y), vf_mul2 (make_float2 (s.y, d.x), make_float2 (s.y, d.x))));
}

VF_HD float2 vf_detect (int k, const float2 *W)
{
  return vf_detect_pair (W[vf_zpos (k)], W[vf_zpos (VF_NFFT - k)]);
}
