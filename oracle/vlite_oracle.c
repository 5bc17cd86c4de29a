/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  See vlite_oracle.h.
 *
 * CPU restatement of the reference chain.  Every stage cites the reference
 * lines it follows (paths relative to /root/reference/).  Compile with
 * -ffp-contract=off: fused multiply-adds are written explicitly (fmaf) exactly
 * where the reference's sm_100a SASS has an FFMA, so that the arithmetic of
 * the non-FFT stages is bit-identical to the GPU reference:
 *
 *   kurtosis:  d4 = fma(b2, b2, a2*a2)          (FMUL, FFMA; a2=x[t]^2, b2=x[t+250]^2)
 *   detect:    p  = fma(re, re, im*im)          (FMUL, FFMA)
 *   bandpass:  bp = fma(bp, 1-s, s*p)           (FMUL, FFMA)
 *   tscrunch_weights: acc = fma(w, x, acc)      (FFMA)
 *
 * Divisions and square roots are IEEE correctly rounded on both sides.  The
 * one operation that is NOT bit-reproducible between this file and the GPU is
 * powf (glibc vs CUDA libdevice, a few ulp): it can move the D'Agostino
 * statistic by ~1e-7 relative, which only matters for a block whose statistic
 * sits within that distance of the 3.0 threshold.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "vlite_oracle.h"

#define DAG_THRESH 3.0              /* src/process_baseband.h:39 */
#define DAG_INF    (3.0 + 5.0 + 1)  /* src/process_baseband.h:42-43 */
#define MIN_WEIGHT 0.2              /* src/process_baseband.h:46 */

struct orc_chain {
  int T;            /* FFTs per pol per segment */
  int nbit, npol, rfi_mode, nthreads;
  size_t nsamp;     /* samples per pol per segment = T*NFFT */
  size_t nblk;      /* kurtosis blocks per pol = T*25 */
  orc_fft_plan *plan;
  float *volt;      /* [2][nsamp] converted voltages (fft_in) */
  float *volt_kur;  /* [2][nsamp] excised voltages (mode 2), else alias */
  float *pw, *kur, *dag;          /* [2][nblk] */
  float *pw_fb, *kur_fb, *dag_fb; /* [2][T] */
  float *w, *w_work;              /* [2][T] weights: as produced / as mutated */
  float *P_raw, *P_kur;           /* [2][T][NCHAN] |X|^2, later normalised in place */
  float *Pdet_raw, *Pdet_kur;     /* copies of |X|^2 kept for inspection */
  float *ave_raw, *ave_kur;       /* [npol][T/8][NCHAN] */
  float *bp_raw, *bp_kur;         /* [2][NCHAN] */
  float *scratch;                 /* per-thread FFT scratch */
  size_t scratch_stride;
  uint32_t histo[512];
};

/* ---- A5: convertarray, src/pb_kernels.cu:23-33 ------------------------- */
void orc_stage_convert (const uint8_t *u, float *x, size_t n)
{
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i)
    x[i] = (u[i] == 0) ? 0.f : (float) u[i] / 128 - 1;
}

/* ---- A4: histogram, src/pb_kernels.cu:321-336 -------------------------- */
static void stage_histogram (const uint8_t *p0, const uint8_t *p1, size_t n, uint32_t *h)
{
  memset (h, 0, 512 * sizeof (uint32_t));
  for (size_t i = 0; i < n; ++i) h[p0[i]]++;
  for (size_t i = 0; i < n; ++i) h[256 + p1[i]]++;
}

/* ---- A6: kurtosis, src/pb_kernels.cu:35-107 ---------------------------- *
 * 256-entry pairwise tree with strides 128..1 (entries 250..255 are zero);
 * the 4th-moment entry of slot t is fma(b2, b2, a2*a2).                    */
void orc_stage_kurtosis (const float *x, float *pw, float *kur, size_t nblock)
{
#pragma omp parallel for schedule(static)
  for (size_t b = 0; b < nblock; ++b) {
    const float *v = x + b * ORC_NKURTO;
    float d2[256], d4[256];
    for (int t = 0; t < 250; ++t) {
      float a2 = v[t] * v[t];
      float b2 = v[t + 250] * v[t + 250];
      d4[t] = fmaf (b2, b2, a2 * a2);
      d2[t] = a2 + b2;
    }
    for (int t = 250; t < 256; ++t) d2[t] = d4[t] = 0.f;
    for (int s = 128; s >= 1; s >>= 1)
      for (int t = 0; t < s; ++t) { d2[t] += d2[t + s]; d4[t] += d4[t + s]; }
    float p = d2[0] / ORC_NKURTO;
    pw[b] = p;
    kur[b] = d4[0] / ORC_NKURTO / (p * p);
  }
}

/* ---- A7/A9: Anscombe-Glynn transform of the sample kurtosis ------------ *
 * src/pb_kernels.cu:3-20 (constants), :109-134, :219-241.  The sample count
 * enters as a float; every literal is a double.  out = {mu1, A, Z1, Z2, Z3}. */
void orc_dagostino_constants (int nsamp, double out[5])
{
  const float n = (float) nsamp;
  const double mu1 = -6. / (n + 1);
  const double mu2 = (24. * n * (n - 2) * (n - 3)) / ((n + 1) * (n + 1) * (n + 3) * (n + 5));
  const double g1 = 6. * (n * n - 5 * n + 2) / ((n + 7) * (n + 9))
                    * sqrt ((6. * (n + 3) * (n + 5)) / (n * (n - 2) * (n - 3)));
  const double A = 6. + (8. / g1) * (2. / g1 + sqrt (1. + 4. / (g1 * g1)));
  out[0] = mu1;
  out[1] = A;
  out[2] = sqrt (4.5 * A);
  out[3] = 1 - 2. / (9 * A);
  out[4] = sqrt (2. / (mu2 * (A - 4)));
}

static inline float dag_one (float k, const double c[5])
{
  float d = (float) DAG_INF;
  if (k != 0.) {
    float t = (float) ((1 - 2. / c[1]) / (1. + (k - 3. - c[0]) * c[4]));
    if (t > 0)
      d = fabsf ((float) (c[2] * (c[3] - powf (t, (float) (1. / 3)))));
  }
  return d;
}

/* kur holds [2][n]; dag receives max over the two pols, duplicated. */
void orc_stage_dagostino (const float *kur, float *dag, size_t n, int nsamp)
{
  double c[5];
  orc_dagostino_constants (nsamp, c);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) {
    float d1 = dag_one (kur[i], c), d2 = dag_one (kur[i + n], c);
    dag[i] = dag[i + n] = fmaxf (d1, d2);
  }
}

/* ---- A8: block_kurtosis, src/pb_kernels.cu:140-212 --------------------- *
 * per FFT block: 32-entry tree (entries 25..31 zero) with strides 16..1.   */
static void stage_block_kurtosis (const float *pw, const float *kur, const float *dag,
                                  float *pw_fb, float *kur_fb, size_t nfft)
{
#pragma omp parallel for schedule(static)
  for (size_t b = 0; b < nfft; ++b) {
    float d2[32], d4[32];
    unsigned char wt[32];
    for (int t = 0; t < 32; ++t) {
      if (t > 24) { d2[t] = d4[t] = 0; wt[t] = 0; continue; }
      size_t idx = b * ORC_SUB + t;
      wt[t] = dag[idx] < DAG_THRESH;
      d2[t] = wt[t] * pw[idx];
      d4[t] = wt[t] * kur[idx] * pw[idx] * pw[idx];
    }
    for (int s = 16; s >= 1; s >>= 1)
      for (int t = 0; t < s; ++t) { d2[t] += d2[t + s]; d4[t] += d4[t + s]; wt[t] += wt[t + s]; }
    if (wt[0] > 0) {
      float p = d2[0] / wt[0];
      pw_fb[b] = p;
      kur_fb[b] = d4[0] / wt[0] / (p * p);
    } else
      pw_fb[b] = kur_fb[b] = 0;
  }
}

/* ---- A10: apply_kurtosis, src/pb_kernels.cu:243-295 -------------------- *
 * blocks with dag > 3 are zeroed; every kept block adds float(500)/12500 to
 * the weight of its FFT (all addends equal, so atomic order is irrelevant). */
static void stage_apply_kurtosis (const float *in, float *out, const float *dag,
                                  float *w, size_t nblk_total)
{
  const float inc = (float) ORC_NKURTO / ORC_NFFT;
  size_t nfft_total = nblk_total / ORC_SUB;
#pragma omp parallel for schedule(static)
  for (size_t f = 0; f < nfft_total; ++f) {
    float acc = 0.f;
    for (int s = 0; s < ORC_SUB; ++s) {
      size_t b = f * ORC_SUB + s;
      int bad = dag[b] > DAG_THRESH;
      float *o = out + b * ORC_NKURTO;
      if (bad)
        memset (o, 0, sizeof (float) * ORC_NKURTO);
      else {
        if (in != out) memcpy (o, in + b * ORC_NKURTO, sizeof (float) * ORC_NKURTO);
        acc += inc;
      }
    }
    w[f] = acc;
  }
}

/* ---- A11 + detect: FFT then |X|^2 = fma(re,re,im*im) ------------------- *
 * src/process_baseband.cu:1222-1224; detection arithmetic from
 * src/pb_kernels.cu:416 / :481 as compiled (FMUL im*im; FFMA re*re+..).    */
static void stage_fft_detect (orc_chain *c, const float *volt, float *P)
{
  const size_t nfft_total = (size_t) 2 * c->T;
#pragma omp parallel
  {
#ifdef _OPENMP
    int tid = omp_get_thread_num ();
#else
    int tid = 0;
#endif
    float *scr = c->scratch + (size_t) tid * c->scratch_stride;
    float *spec = scr + orc_fft_scratch_floats (c->plan);
#pragma omp for schedule(static)
    for (size_t f = 0; f < nfft_total; ++f) {
      orc_rfft (c->plan, volt + f * ORC_NFFT, spec, scr);
      float *p = P + f * ORC_NCHAN;
      for (int k = 0; k < ORC_NCHAN; ++k) {
        float re = spec[2 * k], im = spec[2 * k + 1];
        p[k] = fmaf (re, re, im * im);
      }
    }
  }
}

/* ---- A13: detect_and_normalize2, src/pb_kernels.cu:393-429 ------------- */
static void stage_normalise_raw (orc_chain *c, float *P, float *bp, float scale)
{
  const int T = c->T;
  const float oms = 1 - scale;
#pragma omp parallel for schedule(static)
  for (int pc = 0; pc < 2 * ORC_NCHAN; ++pc) {
    int pol = pc / ORC_NCHAN, ch = pc % ORC_NCHAN;
    float *col = P + (size_t) pol * T * ORC_NCHAN + ch;
    float b = bp[pc];
    if (0. == b) {
      for (int t = 0; t < T; ++t) b += col[(size_t) t * ORC_NCHAN];
      b /= T;
    }
    for (int t = 0; t < T; ++t) {
      float p = col[(size_t) t * ORC_NCHAN];
      b = fmaf (b, oms, scale * p);
      col[(size_t) t * ORC_NCHAN] = p / b - 1;
    }
    bp[pc] = b;
  }
}

/* ---- A14: detect_and_normalize3, src/pb_kernels.cu:431-511 ------------- */
static void stage_normalise_kur (orc_chain *c, float *P, const float *w, float *bp, float scale)
{
  const int T = c->T;
  const float oms = 1 - scale;
#pragma omp parallel for schedule(static)
  for (int pc = 0; pc < 2 * ORC_NCHAN; ++pc) {
    int pol = pc / ORC_NCHAN, ch = pc % ORC_NCHAN;
    float *col = P + (size_t) pol * T * ORC_NCHAN + ch;
    const float *wp = w + (size_t) pol * T;
    float b = bp[pc];
    if (0. == b) {
      int good = 0;
      for (int t = 0; t < T; ++t) {
        if (0. == wp[t]) continue;
        good++;
        b += col[(size_t) t * ORC_NCHAN] / wp[t];
      }
      if (0 == good) b = 1;
      else b /= good;
    }
    for (int t = 0; t < T; ++t) {
      float wt = wp[t];
      float *o = col + (size_t) t * ORC_NCHAN;
      if (0. == wt) { *o = 0; continue; }
      float p = *o / wt;
      if (p > b * 11) { *o = 10; continue; }
      b = fmaf (b, oms, scale * p);
      *o = p / b - 1;
    }
    bp[pc] = b;
  }
}

/* ---- A15: pscrunch / pscrunch_weights, src/pb_kernels.cu:514-560 ------- *
 * In place on the pol-0 half.  The weighted form rewrites the pol-0 weights;
 * the reference does so racily, which is benign because both pols always
 * carry the same weight (src/pb_kernels.cu:132); here the new weights are
 * applied after the pass.                                                   */
static void stage_pscrunch (orc_chain *c, float *P)
{
  const size_t n = (size_t) c->T * ORC_NCHAN;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i)
    P[i] = (float) (M_SQRT1_2 * (P[i] + P[i + n]));
}

static void stage_pscrunch_weights (orc_chain *c, float *P, float *w)
{
  const int T = c->T;
#pragma omp parallel for schedule(static)
  for (int t = 0; t < T; ++t) {
    float w1f = w[t], w2f = w[t + T];
    int w1 = w1f >= MIN_WEIGHT, w2 = w2f >= MIN_WEIGHT;
    float *a = P + (size_t) t * ORC_NCHAN, *b = a + (size_t) T * ORC_NCHAN;
    float neww;
    switch (w1 + w2) {
      case 2:
        for (int k = 0; k < ORC_NCHAN; ++k) a[k] = (float) (M_SQRT1_2 * (a[k] + b[k]));
        neww = (float) (0.5 * (w1f + w2f));
        break;
      case 1:
        for (int k = 0; k < ORC_NCHAN; ++k) a[k] = fmaf (a[k], (float) w1, b[k] * (float) w2);
        neww = fmaf ((float) w1, w1f, (float) w2 * w2f);
        break;
      default:
        for (int k = 0; k < ORC_NCHAN; ++k) a[k] = 0.f;
        neww = 0;
    }
    w[t] = neww;
  }
}

/* ---- A16: tscrunch / tscrunch_weights, src/pb_kernels.cu:564-630 ------- */
static void stage_tscrunch (const float *P, float *ave, size_t nout)
{
  const float scale = (float) sqrt (1. / ORC_NSCRUNCH);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < nout; ++i) {
    size_t src = i + (size_t) (ORC_NSCRUNCH - 1) * (i / ORC_NCHAN) * ORC_NCHAN;
    float acc = 0.f;
    for (int j = 0; j < ORC_NSCRUNCH; ++j, src += ORC_NCHAN) acc += P[src];
    ave[i] = acc * scale;
  }
}

static void stage_tscrunch_weights (const float *P, float *ave, const float *w, size_t nout)
{
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < nout; ++i) {
    size_t src = i + (size_t) (ORC_NSCRUNCH - 1) * (i / ORC_NCHAN) * ORC_NCHAN;
    float acc = 0.f, wsumf = 0.f;
    int wsum = 0;
    for (int j = 0; j < ORC_NSCRUNCH; ++j, src += ORC_NCHAN) {
      float wt = w[src / ORC_NCHAN];
      if (wt < MIN_WEIGHT) continue;
      wsum++;
      wsumf += wt;
      acc = fmaf (wt, P[src], acc);
    }
    if (wsumf / ORC_NSCRUNCH >= MIN_WEIGHT)
      ave[i] = acc / sqrtf ((float) wsum);
    else
      ave[i] = 0;
  }
}

/* ---- A17: sel_and_dig_{2,4,8}b, src/pb_kernels.cu:633-735 -------------- *
 * output order [time][pol][chan]; channel c <- FFT bin c + 2155.           */
void orc_digitise (const float *ave, uint8_t *out, int ntime, int npol, int nbit)
{
  const size_t nsamp = (size_t) ntime * npol * ORC_NCHANOUT;
  const int per = 8 / nbit;
  const size_t nbyte = nsamp / per;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < nbyte; ++i) {
    size_t s0 = i * per;
    int t = (int) (s0 / ((size_t) ORC_NCHANOUT * npol));
    int pol = (int) ((s0 - (size_t) t * ORC_NCHANOUT * npol) / ORC_NCHANOUT);
    int ch = (int) (s0 - (size_t) t * npol * ORC_NCHANOUT - (size_t) pol * ORC_NCHANOUT);
    const float *src = ave + (size_t) pol * ntime * ORC_NCHAN + (size_t) t * ORC_NCHAN + ch + ORC_CHANMIN;
    unsigned v = 0;
    if (nbit == 8) {
      float tmp = (float) (src[0] / 0.02957 + 127.5);
      if (tmp <= 0) v = 0;
      else if (tmp >= 255) v = 255;
      else v = (unsigned char) tmp;
    } else if (nbit == 4) {
      for (int j = 0; j < 2; ++j) {
        float tmp = (float) (src[j] / 0.3188 + 7.5);
        unsigned q;
        if (tmp <= 0) q = 0;
        else if (tmp >= 15) q = 15;
        else q = (unsigned char) tmp;
        v += q << (4 * j);
      }
    } else {
      for (int j = 0; j < 4; ++j) {
        float tmp = src[j];
        if (tmp < -0.6109) continue;
        if (tmp < 0.3970) v += 1u << (2 * j);
        else if (tmp < 1.4050) v += 2u << (2 * j);
        else v += 3u << (2 * j);
      }
    }
    out[i] = (uint8_t) v;
  }
}

/* ------------------------------------------------------------------------ */
orc_chain *orc_create (int ffts_per_seg, int nbit, int npol, int rfi_mode, int nthreads)
{
  if (ffts_per_seg <= 0 || ffts_per_seg % ORC_NSCRUNCH) return NULL;
  if (!(nbit == 2 || nbit == 4 || nbit == 8)) return NULL;
  if (!(npol == 1 || npol == 2)) return NULL;
  if (rfi_mode < 0 || rfi_mode > 2) return NULL;
  orc_chain *c = (orc_chain *) calloc (1, sizeof (*c));
  c->T = ffts_per_seg; c->nbit = nbit; c->npol = npol; c->rfi_mode = rfi_mode;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads ();
  omp_set_num_threads (nthreads);
#else
  nthreads = 1;
#endif
  c->nthreads = nthreads;
  c->nsamp = (size_t) c->T * ORC_NFFT;
  c->nblk = (size_t) c->T * ORC_SUB;
  c->plan = orc_fft_plan_create (ORC_NFFT);
  const size_t T = c->T;
  c->volt = (float *) malloc (sizeof (float) * 2 * c->nsamp);
  c->volt_kur = (rfi_mode == 2) ? (float *) malloc (sizeof (float) * 2 * c->nsamp) : c->volt;
  c->pw = (float *) calloc (6 * c->nblk, sizeof (float));
  c->kur = c->pw + 2 * c->nblk;
  c->dag = c->pw + 4 * c->nblk;
  c->pw_fb = (float *) calloc (10 * T, sizeof (float));
  c->kur_fb = c->pw_fb + 2 * T;
  c->dag_fb = c->pw_fb + 4 * T;
  c->w = c->pw_fb + 6 * T;
  c->w_work = c->pw_fb + 8 * T;
  const size_t np = 2 * T * ORC_NCHAN;
  c->P_raw = (float *) malloc (sizeof (float) * np);
  c->Pdet_raw = (float *) malloc (sizeof (float) * np);
  if (rfi_mode == 2) {
    c->P_kur = (float *) malloc (sizeof (float) * np);
    c->Pdet_kur = (float *) malloc (sizeof (float) * np);
  } else { c->P_kur = c->P_raw; c->Pdet_kur = c->Pdet_raw; }
  const size_t na = (size_t) npol * (T / ORC_NSCRUNCH) * ORC_NCHAN;
  c->ave_raw = (float *) calloc (na, sizeof (float));
  c->ave_kur = (rfi_mode == 2) ? (float *) calloc (na, sizeof (float)) : c->ave_raw;
  c->bp_raw = (float *) calloc (2 * ORC_NCHAN, sizeof (float));
  c->bp_kur = (rfi_mode == 2) ? (float *) calloc (2 * ORC_NCHAN, sizeof (float)) : c->bp_raw;
  c->scratch_stride = orc_fft_scratch_floats (c->plan) + 2 * ORC_NCHAN + 64;
  c->scratch = (float *) malloc (sizeof (float) * c->scratch_stride * (size_t) nthreads);
  return c;
}

void orc_destroy (orc_chain *c)
{
  if (!c) return;
  if (c->rfi_mode == 2) { free (c->volt_kur); free (c->P_kur); free (c->Pdet_kur); free (c->ave_kur); free (c->bp_kur); }
  free (c->volt); free (c->pw); free (c->pw_fb); free (c->P_raw); free (c->Pdet_raw);
  free (c->ave_raw); free (c->bp_raw); free (c->scratch);
  orc_fft_plan_destroy (c->plan);
  free (c);
}

void orc_reset_bandpass (orc_chain *c)
{
  memset (c->bp_raw, 0, sizeof (float) * 2 * ORC_NCHAN);
  memset (c->bp_kur, 0, sizeof (float) * 2 * ORC_NCHAN);
}

size_t orc_out_bytes (const orc_chain *c)
{
  return (size_t) (c->T / ORC_NSCRUNCH) * c->npol * ORC_NCHANOUT * c->nbit / 8;
}

/* Segment sequence of src/process_baseband.cu:1108-1354. */
int orc_process_segment (orc_chain *c, const uint8_t *pol0, const uint8_t *pol1,
                         uint8_t *fb_main, uint8_t *fb_raw)
{
  const int T = c->T, mode = c->rfi_mode;
  /* bp_scale = float(tsamp/tsmooth), src/process_baseband.cu:739-741 */
  const float bp_scale = (float) ((double) ORC_NFFT / 128000000 * ORC_NSCRUNCH / 1.0);
  const size_t np = (size_t) 2 * T * ORC_NCHAN;

  stage_histogram (pol0, pol1, c->nsamp, c->histo);                   /* :1135-1136 */
  orc_stage_convert (pol0, c->volt, c->nsamp);                        /* :1152 */
  orc_stage_convert (pol1, c->volt + c->nsamp, c->nsamp);
  if (mode) {                                                         /* :1160-1204 */
    orc_stage_kurtosis (c->volt, c->pw, c->kur, 2 * c->nblk);
    orc_stage_dagostino (c->kur, c->dag, c->nblk, ORC_NKURTO);
    stage_block_kurtosis (c->pw, c->kur, c->dag, c->pw_fb, c->kur_fb, (size_t) 2 * T);
    orc_stage_dagostino (c->kur_fb, c->dag_fb, (size_t) T, ORC_NFFT);
    stage_apply_kurtosis (c->volt, c->volt_kur, c->dag, c->w, 2 * c->nblk);
    memcpy (c->w_work, c->w, sizeof (float) * 2 * T);
  }
  if (mode == 0 || mode == 2) {                                       /* :1222, :1257-1262 */
    stage_fft_detect (c, c->volt, c->P_raw);
    memcpy (c->Pdet_raw, c->P_raw, sizeof (float) * np);
    stage_normalise_raw (c, c->P_raw, c->bp_raw, bp_scale);
  }
  if (mode == 1 || mode == 2) {                                       /* :1224, :1263-1268 */
    stage_fft_detect (c, c->volt_kur, c->P_kur);
    memcpy (c->Pdet_kur, c->P_kur, sizeof (float) * np);
    stage_normalise_kur (c, c->P_kur, c->w_work, c->bp_kur, bp_scale);
  }
  size_t maxn = (size_t) 2 * T * ORC_NCHAN;
  if (c->npol == 1) {                                                 /* :1278-1291 */
    maxn /= 2;
    if (mode == 0 || mode == 2) stage_pscrunch (c, c->P_raw);
    if (mode == 1 || mode == 2) stage_pscrunch_weights (c, c->P_kur, c->w_work);
  }
  maxn /= ORC_NSCRUNCH;                                               /* :1301-1312 */
  if (mode == 0 || mode == 2) stage_tscrunch (c->P_raw, c->ave_raw, maxn);
  if (mode == 1 || mode == 2) stage_tscrunch_weights (c->P_kur, c->ave_kur, c->w_work, maxn);
  const int ntime = T / ORC_NSCRUNCH;                                 /* :1322-1354 */
  if (mode == 0) {
    if (fb_main) orc_digitise (c->ave_raw, fb_main, ntime, c->npol, c->nbit);
  } else {
    if (fb_main) orc_digitise (c->ave_kur, fb_main, ntime, c->npol, c->nbit);
    if (mode == 2 && fb_raw) orc_digitise (c->ave_raw, fb_raw, ntime, c->npol, c->nbit);
  }
  return 0;
}

const float *orc_get_pow (const orc_chain *c) { return c->pw; }
const float *orc_get_kur (const orc_chain *c) { return c->kur; }
const float *orc_get_dag (const orc_chain *c) { return c->dag; }
const float *orc_get_pow_fb (const orc_chain *c) { return c->pw_fb; }
const float *orc_get_kur_fb (const orc_chain *c) { return c->kur_fb; }
const float *orc_get_dag_fb (const orc_chain *c) { return c->dag_fb; }
const float *orc_get_weights (const orc_chain *c) { return c->w; }
const float *orc_get_ave_main (const orc_chain *c) { return c->rfi_mode ? c->ave_kur : c->ave_raw; }
const float *orc_get_ave_raw (const orc_chain *c) { return c->ave_raw; }
const float *orc_get_power_main (const orc_chain *c) { return c->rfi_mode ? c->Pdet_kur : c->Pdet_raw; }
const float *orc_get_power_raw (const orc_chain *c) { return c->Pdet_raw; }
const float *orc_get_bp_main (const orc_chain *c) { return c->rfi_mode ? c->bp_kur : c->bp_raw; }
const float *orc_get_bp_raw (const orc_chain *c) { return c->bp_raw; }
const uint32_t *orc_get_histo (const orc_chain *c) { return c->histo; }

/* which: 0 main, 1 raw; bp [2][6251] */
void orc_set_bandpass (orc_chain *c, int which, const float *bp)
{
  float *dst = (which == 0 && c->rfi_mode) ? c->bp_kur : c->bp_raw;
  memcpy (dst, bp, sizeof (float) * 2 * ORC_NCHAN);
}
