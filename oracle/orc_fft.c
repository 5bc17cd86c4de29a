/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * Single-precision real-to-complex FFT for the CPU oracle.  The reference
 * calls cuFFT (cufftPlan1d (NFFT, CUFFT_R2C, batch) at
 * src/process_baseband.cu:597-598, exec at :1222-1224): an unnormalised
 * forward transform X[k] = sum_n x[n] exp(-2 pi i n k / N), k = 0..N/2.
 * cuFFT is a closed library (and FFTW3 is not installed in this image), so
 * this file restates the published algorithm -- a Stockham autosort
 * mixed-radix complex FFT of length N/2 over the even/odd packed samples
 * plus the standard real-input split pass -- and is pinned against
 * numpy.fft.rfft (float64) in tests/test_oracle_fft.py.
 *
 * N must be even and N/2 must factor into 2, 3, 4 and 5 (12500/2 = 2 * 5^5).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "vlite_oracle.h"

typedef struct { float re, im; } cf;

struct orc_fft_plan {
  int n;          /* real length */
  int h;          /* complex length n/2 */
  int nstage;
  int radix[32];
  cf *tw[32];     /* per-stage twiddles: tw[s][p*(r-1)+(k-1)] = w_cur^(p k) */
  cf *split;      /* exp(-2 pi i k / n), k = 0..h/2 */
};

static int factorise (int h, int *radix)
{
  int ns = 0;
  static const int cand[] = {5, 4, 3, 2};
  for (int c = 0; c < 4; ++c)
    while (h % cand[c] == 0) { radix[ns++] = cand[c]; h /= cand[c]; }
  return h == 1 ? ns : -1;
}

orc_fft_plan *orc_fft_plan_create (int n)
{
  if (n < 2 || (n & 1)) return NULL;
  orc_fft_plan *pl = (orc_fft_plan *) calloc (1, sizeof (*pl));
  pl->n = n;
  pl->h = n / 2;
  pl->nstage = factorise (pl->h, pl->radix);
  if (pl->nstage < 0) { free (pl); return NULL; }
  int cur = pl->h;
  for (int s = 0; s < pl->nstage; ++s) {
    int r = pl->radix[s], m = cur / r;
    pl->tw[s] = (cf *) malloc (sizeof (cf) * (size_t) m * (r - 1));
    for (int p = 0; p < m; ++p)
      for (int k = 1; k < r; ++k) {
        double a = -2.0 * M_PI * (double) p * (double) k / (double) cur;
        pl->tw[s][p * (r - 1) + (k - 1)].re = (float) cos (a);
        pl->tw[s][p * (r - 1) + (k - 1)].im = (float) sin (a);
      }
    cur = m;
  }
  pl->split = (cf *) malloc (sizeof (cf) * (size_t) (pl->h / 2 + 1));
  for (int k = 0; k <= pl->h / 2; ++k) {
    double a = -2.0 * M_PI * (double) k / (double) n;
    pl->split[k].re = (float) cos (a);
    pl->split[k].im = (float) sin (a);
  }
  return pl;
}

void orc_fft_plan_destroy (orc_fft_plan *pl)
{
  if (!pl) return;
  for (int s = 0; s < pl->nstage; ++s) free (pl->tw[s]);
  free (pl->split);
  free (pl);
}

size_t orc_fft_scratch_floats (const orc_fft_plan *pl)
{
  return (size_t) 4 * pl->h;   /* two complex buffers of length h */
}

static inline cf cmul (cf a, cf b)
{
  cf c = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re };
  return c;
}

/* One Stockham stage: length-cur sub-transforms at stride s.
 *   y[q + s (r p + k)] = w^(p k) * DFT_r { x[q + s (p + m j)] }_j        */
static void stage (const cf *restrict x, cf *restrict y, int cur, int s,
                   int r, const cf *restrict tw)
{
  const int m = cur / r;
  const float c51 = 0.30901699437494745f, c52 = -0.80901699437494745f;
  const float s51 = 0.95105651629515353f, s52 = 0.58778525229247314f;
  const float s3 = 0.86602540378443865f;
  for (int p = 0; p < m; ++p) {
    const cf *w = tw + (size_t) p * (r - 1);
    for (int q = 0; q < s; ++q) {
      const cf *xi = x + q + (size_t) s * p;
      cf *yo = y + q + (size_t) s * r * p;
      const size_t sm = (size_t) s * m;
      if (r == 5) {
        cf a0 = xi[0], a1 = xi[sm], a2 = xi[2 * sm], a3 = xi[3 * sm], a4 = xi[4 * sm];
        cf t1 = { a1.re + a4.re, a1.im + a4.im }, t2 = { a2.re + a3.re, a2.im + a3.im };
        cf t3 = { a1.re - a4.re, a1.im - a4.im }, t4 = { a2.re - a3.re, a2.im - a3.im };
        cf b0 = { a0.re + t1.re + t2.re, a0.im + t1.im + t2.im };
        cf m1 = { a0.re + c51 * t1.re + c52 * t2.re, a0.im + c51 * t1.im + c52 * t2.im };
        cf m2 = { a0.re + c52 * t1.re + c51 * t2.re, a0.im + c52 * t1.im + c51 * t2.im };
        cf u1 = { s51 * t3.re + s52 * t4.re, s51 * t3.im + s52 * t4.im };
        cf u2 = { s52 * t3.re - s51 * t4.re, s52 * t3.im - s51 * t4.im };
        /* forward transform: b_k = m - i*u style combination */
        cf b1 = { m1.re + u1.im, m1.im - u1.re }, b4 = { m1.re - u1.im, m1.im + u1.re };
        cf b2 = { m2.re + u2.im, m2.im - u2.re }, b3 = { m2.re - u2.im, m2.im + u2.re };
        yo[0] = b0;
        yo[s] = cmul (b1, w[0]);
        yo[2 * s] = cmul (b2, w[1]);
        yo[3 * s] = cmul (b3, w[2]);
        yo[4 * s] = cmul (b4, w[3]);
      } else if (r == 2) {
        cf a0 = xi[0], a1 = xi[sm];
        cf b0 = { a0.re + a1.re, a0.im + a1.im }, b1 = { a0.re - a1.re, a0.im - a1.im };
        yo[0] = b0;
        yo[s] = cmul (b1, w[0]);
      } else if (r == 4) {
        cf a0 = xi[0], a1 = xi[sm], a2 = xi[2 * sm], a3 = xi[3 * sm];
        cf e0 = { a0.re + a2.re, a0.im + a2.im }, e1 = { a0.re - a2.re, a0.im - a2.im };
        cf o0 = { a1.re + a3.re, a1.im + a3.im }, o1 = { a1.re - a3.re, a1.im - a3.im };
        cf b0 = { e0.re + o0.re, e0.im + o0.im }, b2 = { e0.re - o0.re, e0.im - o0.im };
        cf b1 = { e1.re + o1.im, e1.im - o1.re }, b3 = { e1.re - o1.im, e1.im + o1.re };
        yo[0] = b0;
        yo[s] = cmul (b1, w[0]);
        yo[2 * s] = cmul (b2, w[1]);
        yo[3 * s] = cmul (b3, w[2]);
      } else { /* r == 3 */
        cf a0 = xi[0], a1 = xi[sm], a2 = xi[2 * sm];
        cf t1 = { a1.re + a2.re, a1.im + a2.im }, t2 = { a1.re - a2.re, a1.im - a2.im };
        cf b0 = { a0.re + t1.re, a0.im + t1.im };
        cf mm = { a0.re - 0.5f * t1.re, a0.im - 0.5f * t1.im };
        cf b1 = { mm.re + s3 * t2.im, mm.im - s3 * t2.re };
        cf b2 = { mm.re - s3 * t2.im, mm.im + s3 * t2.re };
        yo[0] = b0;
        yo[s] = cmul (b1, w[0]);
        yo[2 * s] = cmul (b2, w[1]);
      }
    }
  }
}

/* in: n reals; out: n/2+1 complex (interleaved re,im); scratch: 4*h floats */
void orc_rfft (const orc_fft_plan *pl, const float *in, float *out, float *scratch)
{
  const int h = pl->h;
  cf *a = (cf *) scratch, *b = a + h;
  memcpy (a, in, sizeof (float) * (size_t) pl->n);   /* z[n] = x[2n] + i x[2n+1] */
  int cur = h, s = 1;
  for (int st = 0; st < pl->nstage; ++st) {
    stage (a, b, cur, s, pl->radix[st], pl->tw[st]);
    cur /= pl->radix[st];
    s *= pl->radix[st];
    cf *t = a; a = b; b = t;
  }
  /* split: X[k] = E[k] + w^k O[k],  E = (Z[k]+conj Z[h-k])/2,
   *        O = (Z[k]-conj Z[h-k])/(2i) */
  cf *X = (cf *) out;
  X[0].re = a[0].re + a[0].im;  X[0].im = 0.f;
  X[h].re = a[0].re - a[0].im;  X[h].im = 0.f;
  for (int k = 1; k <= h / 2; ++k) {
    cf zk = a[k], zc = { a[h - k].re, -a[h - k].im };
    cf e = { 0.5f * (zk.re + zc.re), 0.5f * (zk.im + zc.im) };
    cf d = { 0.5f * (zk.re - zc.re), 0.5f * (zk.im - zc.im) };
    cf o = { d.im, -d.re };                       /* d / i */
    cf wo = cmul (o, pl->split[k]);
    X[k].re = e.re + wo.re;      X[k].im = e.im + wo.im;
    X[h - k].re = e.re - wo.re;  X[h - k].im = -(e.im - wo.im);
  }
}
