/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * "Oracle B": replay of the reference's per-segment launch sequence
 * (src/process_baseband.cu:1108-1375) around the reference's OWN kernels.
 * This file contains no kernel: it is linked with pb_kernels.o, which
 * oracle/Makefile compiles UNMODIFIED from $(REF)/src/pb_kernels.cu (the
 * prototypes come from $(REF)/src/process_baseband.h), and with cuFFT.
 * Buffer sizes, aliasing rules and launch shapes follow the allocation block
 * src/process_baseband.cu:578-709 and the launches at :1135-1354 literally
 * (nsms*32 x 512 grids, cufftPlan1d (NFFT, CUFFT_R2C, 2048), memset of the
 * weights, RFI_MODE aliasing :606-613,661-664,684-686,703-709).
 *
 * The geometry is the reference's compile-time geometry: 1024 FFTs per pol
 * per segment.  Used (a) to generate tests/golden/ (scripts/make_golden.py),
 * (b) by the -m gpu parity tests, (c) as the "legacy CUDA" timing reported by
 * bench.py.  Never part of the product path.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime.h>
#include <cufft.h>
#include "process_baseband.h"   /* from the reference tree (-I$(REF)/src) */

#define RCHK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf (stderr, "ref_replay: %s at %s:%d\n", cudaGetErrorString (e_), __FILE__, __LINE__); return 20; } } while (0)

struct ref_chain {
  int nbit, npol, rfi_mode, nsms;
  size_t samps_per_chunk;      /* both pols */
  int fft_per_chunk;           /* 2 * 1024 */
  size_t nkurto_per_chunk;
  int scrunch, trim;
  cufftHandle plan;
  unsigned char *udat_dev;
  unsigned int *histo_dev;
  cufftReal *fft_in, *fft_in_kur;
  cufftComplex *fft_out, *fft_out_kur;
  cufftReal *pow_dev, *kur_dev, *pow_fb_dev, *kur_fb_dev, *dag_dev, *dag_fb_dev, *kur_weights_dev;
  cufftReal *weights_snapshot;  /* weights as left by apply_kurtosis (WRITE_KURTO copy, :1206-1210) */
  cufftReal *fft_ave, *fft_ave_kur;
  unsigned char *fft_trim_u_dev, *fft_trim_u_kur_dev;
  cufftReal *bp_dev, *bp_kur_dev;
  float *det_raw, *det_kur;     /* optional |X|^2 snapshots, [2][1024][6251] */
  int keep_det, do_histo;
  float *frb_delays_dev;
  cudaEvent_t ev0, ev1;
  float last_ms;
};

/* |X|^2 of a cuFFT output, written with the same expression as the reference
 * detection line (src/pb_kernels.cu:416); inspection only. */
__global__ void ref_snapshot_power (const cufftComplex *x, float *p, size_t n)
{
  for (size_t i = threadIdx.x + (size_t) blockIdx.x * blockDim.x; i < n; i += (size_t) blockDim.x * gridDim.x)
    p[i] = x[i].x * x[i].x + x[i].y * x[i].y;
}

extern "C" {

int ref_ffts_per_seg (void) { return (int) (FFTS_PER_SEG); }

int ref_create (int nbit, int npol, int rfi_mode, int keep_det, int do_histo, int inject_frb, ref_chain **out)
{
  ref_chain *c = (ref_chain *) calloc (1, sizeof (*c));
  c->nbit = nbit; c->npol = npol; c->rfi_mode = rfi_mode; c->keep_det = keep_det; c->do_histo = do_histo;
  int dev = 0;
  RCHK (cudaGetDevice (&dev));
  RCHK (cudaDeviceGetAttribute (&c->nsms, cudaDevAttrMultiProcessorCount, dev));      /* :474-475 */
  c->samps_per_chunk = 2 * VLITE_RATE / SEG_PER_SEC;                                   /* :593 */
  c->fft_per_chunk = 2 * FFTS_PER_SEG;                                                 /* :595 */
  if (cufftPlan1d (&c->plan, NFFT, CUFFT_R2C, c->fft_per_chunk) != CUFFT_SUCCESS) return 21;  /* :597-598 */
  RCHK (cudaMalloc ((void **) &c->udat_dev, c->samps_per_chunk));                      /* :581-582 */
  RCHK (cudaMalloc ((void **) &c->histo_dev, 2 * 256 * sizeof (unsigned int)));
  RCHK (cudaMalloc ((void **) &c->fft_in, sizeof (cufftReal) * c->samps_per_chunk));   /* :601-604 */
  RCHK (cudaMalloc ((void **) &c->fft_out, sizeof (cufftComplex) * c->fft_per_chunk * NCHAN));
  c->fft_in_kur = c->fft_in;                                                           /* :606-613 */
  c->fft_out_kur = c->fft_out;
  if (2 == rfi_mode) {
    RCHK (cudaMalloc ((void **) &c->fft_in_kur, sizeof (cufftReal) * c->samps_per_chunk));
    RCHK (cudaMalloc ((void **) &c->fft_out_kur, sizeof (cufftComplex) * c->fft_per_chunk * NCHAN));
  }
  c->nkurto_per_chunk = c->samps_per_chunk / NKURTO;                                   /* :618 */
  if (rfi_mode) {                                                                      /* :622-643 */
    RCHK (cudaMalloc ((void **) &c->pow_dev, 2 * sizeof (cufftReal) * c->nkurto_per_chunk));
    RCHK (cudaMalloc ((void **) &c->pow_fb_dev, 2 * sizeof (cufftReal) * c->fft_per_chunk));
    c->kur_dev = c->pow_dev + c->nkurto_per_chunk;
    c->kur_fb_dev = c->pow_fb_dev + c->fft_per_chunk;
    RCHK (cudaMalloc ((void **) &c->dag_dev, sizeof (cufftReal) * c->nkurto_per_chunk));
    RCHK (cudaMalloc ((void **) &c->dag_fb_dev, sizeof (cufftReal) * c->fft_per_chunk));
    RCHK (cudaMalloc ((void **) &c->kur_weights_dev, sizeof (cufftReal) * c->fft_per_chunk));
    RCHK (cudaMalloc ((void **) &c->weights_snapshot, sizeof (cufftReal) * c->fft_per_chunk));
  }
  int polfac = npol == 1 ? 2 : 1;                                                      /* :656-664 */
  c->scrunch = (c->fft_per_chunk * NCHAN) / (polfac * NSCRUNCH);
  RCHK (cudaMalloc ((void **) &c->fft_ave, sizeof (cufftReal) * c->scrunch));
  c->fft_ave_kur = c->fft_ave;
  if (2 == rfi_mode) RCHK (cudaMalloc ((void **) &c->fft_ave_kur, sizeof (cufftReal) * c->scrunch));
  c->trim = (c->fft_per_chunk * (CHANMAX - CHANMIN + 1)) / (polfac * NSCRUNCH);        /* :667-675 */
  if (c->trim % (8 / nbit) != 0) return 22;
  c->trim /= (8 / nbit);
  RCHK (cudaMalloc ((void **) &c->fft_trim_u_dev, c->trim));                           /* :676-686 */
  c->fft_trim_u_kur_dev = c->fft_trim_u_dev;
  if (2 == rfi_mode) RCHK (cudaMalloc ((void **) &c->fft_trim_u_kur_dev, c->trim));
  RCHK (cudaMalloc ((void **) &c->bp_dev, sizeof (cufftReal) * NCHAN * 2));            /* :700-709 */
  RCHK (cudaMemset (c->bp_dev, 0, sizeof (cufftReal) * NCHAN * 2));
  c->bp_kur_dev = c->bp_dev;
  if (2 == rfi_mode) {
    RCHK (cudaMalloc ((void **) &c->bp_kur_dev, sizeof (cufftReal) * NCHAN * 2));
    RCHK (cudaMemset (c->bp_kur_dev, 0, sizeof (cufftReal) * NCHAN * 2));
  }
  if (keep_det) {
    RCHK (cudaMalloc ((void **) &c->det_raw, sizeof (float) * c->fft_per_chunk * NCHAN));
    c->det_kur = c->det_raw;
    if (2 == rfi_mode) RCHK (cudaMalloc ((void **) &c->det_kur, sizeof (float) * c->fft_per_chunk * NCHAN));
  }
  if (inject_frb) {                                                                    /* :712-717 */
    RCHK (cudaMalloc ((void **) &c->frb_delays_dev, sizeof (float) * NCHAN));
    set_frb_delays <<< NCHAN / NTHREAD + 1, NTHREAD >>> (c->frb_delays_dev, 80);
    RCHK (cudaGetLastError ());
  }
  RCHK (cudaEventCreate (&c->ev0));
  RCHK (cudaEventCreate (&c->ev1));
  *out = c;
  return 0;
}

int ref_reset_bandpass (ref_chain *c)
{
  RCHK (cudaMemset (c->bp_dev, 0, sizeof (cufftReal) * NCHAN * 2));
  if (c->bp_kur_dev != c->bp_dev) RCHK (cudaMemset (c->bp_kur_dev, 0, sizeof (cufftReal) * NCHAN * 2));
  return 0;
}

size_t ref_out_bytes (const ref_chain *c) { return (size_t) c->trim; }

/* device part of one segment, udat_dev already filled.  inject_frb_now as in
 * the reference (:1098-1101, :1231-1251): 0 = no injection, k >= 1 = k-th
 * segment since the FRB second started. */
static int ref_run_device (ref_chain *c, int inject_frb_now)
{
  const int nsms = c->nsms, RFI_MODE = c->rfi_mode, npol = c->npol, NBIT = c->nbit;
  const size_t samps_per_chunk = c->samps_per_chunk;
  const int fft_per_chunk = c->fft_per_chunk;
  const size_t nkurto_per_chunk = c->nkurto_per_chunk;
  double tsmooth = 1;                                                                  /* :739-741 */
  double tsamp = double (NFFT) / VLITE_RATE * NSCRUNCH;
  float bp_scale = tsamp / tsmooth;
  int polfac = npol == 1 ? 2 : 1;
  size_t maxn;

  if (c->do_histo) {                                                                   /* :1135-1136 */
    RCHK (cudaMemset (c->histo_dev, 0, 2 * 256 * sizeof (unsigned int)));
    histogram <<<nsms * 32, 512>>> (c->udat_dev, c->histo_dev, samps_per_chunk);
    RCHK (cudaGetLastError ());
  }
  convertarray <<<nsms * 32, NTHREAD>>> (c->fft_in, c->udat_dev, samps_per_chunk);      /* :1152 */
  RCHK (cudaGetLastError ());
  if (RFI_MODE) {                                                                      /* :1160-1204 */
    kurtosis <<<nkurto_per_chunk, 256>>> (c->fft_in, c->pow_dev, c->kur_dev);
    RCHK (cudaGetLastError ());
    compute_dagostino <<<nsms * 32, NTHREAD>>> (c->kur_dev, c->dag_dev, nkurto_per_chunk / 2);
    RCHK (cudaGetLastError ());
    block_kurtosis <<<fft_per_chunk / 8, 256>>> (c->pow_dev, c->kur_dev, c->dag_dev, c->pow_fb_dev, c->kur_fb_dev);
    RCHK (cudaGetLastError ());
    compute_dagostino2 <<<nsms * 32, NTHREAD>>> (c->kur_fb_dev, c->dag_fb_dev, fft_per_chunk / 2);
    RCHK (cudaGetLastError ());
    RCHK (cudaMemset (c->kur_weights_dev, 0, sizeof (cufftReal) * fft_per_chunk));
    apply_kurtosis <<<nkurto_per_chunk, 256>>> (c->fft_in, c->fft_in_kur, c->dag_dev, c->dag_fb_dev, c->kur_weights_dev);
    RCHK (cudaGetLastError ());
    RCHK (cudaMemcpyAsync (c->weights_snapshot, c->kur_weights_dev, sizeof (cufftReal) * fft_per_chunk,
                           cudaMemcpyDeviceToDevice, 0));
  }
  if (cufftExecR2C (c->plan, c->fft_in, c->fft_out) != CUFFT_SUCCESS) return 21;        /* :1222-1224 */
  if (RFI_MODE > 1)
    if (cufftExecR2C (c->plan, c->fft_in_kur, c->fft_out_kur) != CUFFT_SUCCESS) return 21;
  if (inject_frb_now > 0 && c->frb_delays_dev) {                                       /* :1231-1251 */
    float frb_width = 2e-3 * SEG_PER_SEC * FFTS_PER_SEG;
    float frb_amp = 1.05;
    int nfft_since_frb = (inject_frb_now - 1) * FFTS_PER_SEG;
    inject_frb <<< NCHAN / NTHREAD + 1, NTHREAD >>> (c->fft_out, c->frb_delays_dev, nfft_since_frb, frb_width, frb_amp);
    RCHK (cudaGetLastError ());
    if (RFI_MODE > 1) {
      inject_frb <<< NCHAN / NTHREAD + 1, NTHREAD >>> (c->fft_out_kur, c->frb_delays_dev, nfft_since_frb, frb_width, frb_amp);
      RCHK (cudaGetLastError ());
    }
  }
  if (c->keep_det) {
    ref_snapshot_power <<<nsms * 32, NTHREAD>>> (c->fft_out, c->det_raw, (size_t) fft_per_chunk * NCHAN);
    if (RFI_MODE > 1)
      ref_snapshot_power <<<nsms * 32, NTHREAD>>> (c->fft_out_kur, c->det_kur, (size_t) fft_per_chunk * NCHAN);
    RCHK (cudaGetLastError ());
  }
  if (RFI_MODE == 0 || RFI_MODE == 2) {                                                /* :1257-1268 */
    detect_and_normalize2 <<<(NCHAN * 2) / NTHREAD + 1, NTHREAD>>> (c->fft_out, c->bp_dev, bp_scale);
    RCHK (cudaGetLastError ());
  }
  if (RFI_MODE == 1 || RFI_MODE == 2) {
    detect_and_normalize3 <<<(NCHAN * 2) / NTHREAD + 1, NTHREAD>>> (c->fft_out_kur, c->kur_weights_dev, c->bp_kur_dev, bp_scale);
    RCHK (cudaGetLastError ());
  }
  maxn = (fft_per_chunk * NCHAN) / polfac;                                             /* :1278-1291 */
  if (npol == 1) {
    if (RFI_MODE == 0 || RFI_MODE == 2) {
      pscrunch <<<nsms * 32, NTHREAD>>> (c->fft_out, maxn);
      RCHK (cudaGetLastError ());
    }
    if (RFI_MODE == 1 || RFI_MODE == 2) {
      pscrunch_weights <<<nsms * 32, NTHREAD>>> (c->fft_out_kur, c->kur_weights_dev, maxn);
      RCHK (cudaGetLastError ());
    }
  }
  maxn /= NSCRUNCH;                                                                    /* :1301-1312 */
  if (RFI_MODE == 0 || RFI_MODE == 2) {
    tscrunch <<<nsms * 32, NTHREAD>>> (c->fft_out, c->fft_ave, maxn);
    RCHK (cudaGetLastError ());
  }
  if (RFI_MODE == 1 || RFI_MODE == 2) {
    tscrunch_weights <<<nsms * 32, NTHREAD>>> (c->fft_out_kur, c->fft_ave_kur, c->kur_weights_dev, maxn);
    RCHK (cudaGetLastError ());
  }
  maxn = (CHANMAX - CHANMIN + 1) * (maxn / NCHAN) / (8 / NBIT);                         /* :1322-1354 */
  switch (NBIT) {
    case 4:
      sel_and_dig_4b <<<nsms * 32, NTHREAD>>> (c->fft_ave, c->fft_trim_u_dev, maxn, npol);
      if (RFI_MODE > 1) sel_and_dig_4b <<<nsms * 32, NTHREAD>>> (c->fft_ave_kur, c->fft_trim_u_kur_dev, maxn, npol);
      break;
    case 8:
      sel_and_dig_8b <<<nsms * 32, NTHREAD>>> (c->fft_ave, c->fft_trim_u_dev, maxn, npol);
      if (RFI_MODE > 1) sel_and_dig_8b <<<nsms * 32, NTHREAD>>> (c->fft_ave_kur, c->fft_trim_u_kur_dev, maxn, npol);
      break;
    default:
      sel_and_dig_2b <<<nsms * 32, NTHREAD>>> (c->fft_ave, c->fft_trim_u_dev, maxn, npol);
      if (RFI_MODE > 1) sel_and_dig_2b <<<nsms * 32, NTHREAD>>> (c->fft_ave_kur, c->fft_trim_u_kur_dev, maxn, npol);
      break;
  }
  RCHK (cudaGetLastError ());
  return 0;
}

/* one segment from HOST buffers, synchronous copies as in the reference
 * (:1117-1122 in, :1370-1375 out).  fb_main <- excised stream (modes 1,2) or
 * raw (mode 0); fb_raw <- raw stream of mode 2. */
int ref_process_segment (ref_chain *c, const uint8_t *pol0, const uint8_t *pol1,
                         uint8_t *fb_main, uint8_t *fb_raw, int inject_frb_now)
{
  RCHK (cudaEventRecord (c->ev0, 0));
  RCHK (cudaMemcpy (c->udat_dev, pol0, c->samps_per_chunk / 2, cudaMemcpyHostToDevice));
  RCHK (cudaMemcpy (c->udat_dev + c->samps_per_chunk / 2, pol1, c->samps_per_chunk / 2, cudaMemcpyHostToDevice));
  int rc = ref_run_device (c, inject_frb_now);
  if (rc) return rc;
  if (fb_main) RCHK (cudaMemcpy (fb_main, c->fft_trim_u_kur_dev, c->trim, cudaMemcpyDeviceToHost));
  if (2 == c->rfi_mode && fb_raw) RCHK (cudaMemcpy (fb_raw, c->fft_trim_u_dev, c->trim, cudaMemcpyDeviceToHost));
  RCHK (cudaEventRecord (c->ev1, 0));
  RCHK (cudaEventSynchronize (c->ev1));
  RCHK (cudaEventElapsedTime (&c->last_ms, c->ev0, c->ev1));
  return 0;
}

/* device-resident timing: nseg segments from d_in ([nseg][2][12.8M] bytes on
 * the device), kernels only (device-to-device staging of the samples). */
int ref_time_device (ref_chain *c, const uint8_t *d_in, int nseg, float *ms)
{
  RCHK (cudaEventRecord (c->ev0, 0));
  for (int s = 0; s < nseg; ++s) {
    RCHK (cudaMemcpyAsync (c->udat_dev, d_in + (size_t) s * c->samps_per_chunk, c->samps_per_chunk,
                           cudaMemcpyDeviceToDevice, 0));
    int rc = ref_run_device (c, 0);
    if (rc) return rc;
  }
  RCHK (cudaEventRecord (c->ev1, 0));
  RCHK (cudaEventSynchronize (c->ev1));
  RCHK (cudaEventElapsedTime (ms, c->ev0, c->ev1));
  return 0;
}

float ref_last_ms (const ref_chain *c) { return c->last_ms; }

/* which: 0 pow 1 kur 2 dag 3 pow_fb 4 kur_fb 5 dag_fb 6 weights (after
 * apply_kurtosis) 7 ave_main 8 ave_raw 9 det_main 10 det_raw 11 bp_main
 * 12 bp_raw 13 histo 14 weights as mutated by pscrunch_weights.
 * Returns the number of 4-byte elements copied, or -1. */
long ref_get (ref_chain *c, int which, void *out)
{
  const void *src = NULL;
  size_t n = 0;
  const int mode = c->rfi_mode;
  switch (which) {
    case 0: src = c->pow_dev; n = c->nkurto_per_chunk; break;
    case 1: src = c->kur_dev; n = c->nkurto_per_chunk; break;
    case 2: src = c->dag_dev; n = c->nkurto_per_chunk; break;
    case 3: src = c->pow_fb_dev; n = c->fft_per_chunk; break;
    case 4: src = c->kur_fb_dev; n = c->fft_per_chunk; break;
    case 5: src = c->dag_fb_dev; n = c->fft_per_chunk; break;
    case 6: src = c->weights_snapshot; n = c->fft_per_chunk; break;
    case 7: src = mode ? c->fft_ave_kur : c->fft_ave; n = c->scrunch; break;
    case 8: src = c->fft_ave; n = c->scrunch; break;
    case 9: src = mode ? c->det_kur : c->det_raw; n = (size_t) c->fft_per_chunk * NCHAN; break;
    case 10: src = c->det_raw; n = (size_t) c->fft_per_chunk * NCHAN; break;
    case 11: src = mode ? c->bp_kur_dev : c->bp_dev; n = 2 * NCHAN; break;
    case 12: src = c->bp_dev; n = 2 * NCHAN; break;
    case 13: src = c->histo_dev; n = 512; break;
    case 14: src = c->kur_weights_dev; n = c->fft_per_chunk; break;
    default: return -1;
  }
  if (!src) return -1;
  if (cudaMemcpy (out, src, n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (long) n;
}

int ref_destroy (ref_chain *c)
{
  if (!c) return 0;
  cufftDestroy (c->plan);
  if (c->fft_in_kur != c->fft_in) cudaFree (c->fft_in_kur);
  if (c->fft_out_kur != c->fft_out) cudaFree (c->fft_out_kur);
  if (c->fft_ave_kur != c->fft_ave) cudaFree (c->fft_ave_kur);
  if (c->fft_trim_u_kur_dev != c->fft_trim_u_dev) cudaFree (c->fft_trim_u_kur_dev);
  if (c->bp_kur_dev != c->bp_dev) cudaFree (c->bp_kur_dev);
  if (c->det_kur != c->det_raw) cudaFree (c->det_kur);
  cudaFree (c->udat_dev); cudaFree (c->histo_dev); cudaFree (c->fft_in); cudaFree (c->fft_out);
  cudaFree (c->pow_dev); cudaFree (c->pow_fb_dev); cudaFree (c->dag_dev); cudaFree (c->dag_fb_dev);
  cudaFree (c->kur_weights_dev); cudaFree (c->weights_snapshot); cudaFree (c->fft_ave);
  cudaFree (c->fft_trim_u_dev); cudaFree (c->bp_dev); cudaFree (c->det_raw); cudaFree (c->frb_delays_dev);
  cudaEventDestroy (c->ev0); cudaEventDestroy (c->ev1);
  free (c);
  return 0;
}

} /* extern "C" */
