/*
 * libvlitefast host side: the C ABI of include/vlitefast.h over the sm_100a
 * kernels of vf_kernels.cu.  It replaces the allocation block and the
 * per-segment loop body of the reference's main()
 * (src/process_baseband.cu:472-475, 578-709, 1108-1375).  No CPU fallback:
 * every entry point that computes needs a CUDA device and fails otherwise.
 *
 * Streams.  The handle owns one control stream and two "slots".  A slot is a
 * stream plus one set of device buffers (samples, power tiles, weights,
 * packed output).  Consecutive segments alternate between the slots so that
 * the channeliser (K1) of segment n+1 overlaps the normaliser (K2) and the
 * copies of segment n.  The only cross-segment dependency of the chain is the
 * running bandpass (src/pb_kernels.cu:406-428, state :700-709): K2 launches
 * are chained by an event so that they run in segment order.
 */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <dlfcn.h>
#include <sched.h>
#include <ctype.h>
#include <vector>
#include <cuda_runtime.h>
#include "vlitefast.h"
#include "vf_kernels.h"
#ifdef VF_TESTING
#include "vf_testing.h"
#endif

#define VF_MAX_TIMED_SEG 256

struct vf_slot {
  cudaStream_t st;
  uint8_t *d_in;              /* [n_ant][2][T*12500] */
  uint8_t *d_frames;          /* raw VDIF staging, lazily allocated */
  size_t frames_cap;
  float2 *P_raw, *P_kur;      /* [n_ant][T][4096] */
  float *w;                   /* [n_ant][T] */
  uint32_t *mask;             /* [n_ant][T] */
  uint8_t *d_out_main, *d_out_raw;   /* [n_ant][out_bytes] */
  cudaEvent_t ev_k2, ev_done;
  cudaStream_t st_k2;         /* the normaliser's launches: highest priority (see vf_enqueue_segment) */
  cudaEvent_t ev_k1done, ev_k2done;
  cudaEvent_t ev_t[5];        /* asynchronous submissions: start, before K1, after K1, after K2, end (vf_slot_elapsed_ms) */
  int t_valid;
  int pending;                /* vf_submit_*_async issued, vf_wait not yet called */
  unsigned int *d_bad, *h_bad;/* VDIF input: [0] frames outside the window, [1] frames of another second, [2] frames placed, [3] invalid-bit frames */
  uint32_t first_frame;
  uint8_t *d_in_blk;          /* depacketised samples of a block of segments (vf_submit_vdif_block_async), lazily allocated */
  uint8_t *d_out_blk[2];      /* its packed outputs (main, raw) */
  int blk_cap;                /* segments those buffers hold */
  size_t blk_expected;        /* frames a complete block has */
  unsigned int *d_work;       /* item counter of the pipelined channeliser; never reset: */
  unsigned int work_base;     /* its value when the next launch starts                   */
  /* statistics dumps of the segment this slot processed last (keep_stats / do_histo): one set per slot,
   * because the channelisers of the two slots run on independent streams */
  float *pw, *kur, *dag, *pw_fb, *kur_fb, *dag_fb;
  unsigned int *histo;
};

/* minimal NCCL surface, resolved with dlopen at vf_coadd_init */
typedef struct { char internal[128]; } vf_nccl_id;
typedef void *vf_nccl_comm;
struct vf_nccl_api {
  void *lib;
  int (*GetUniqueId) (vf_nccl_id *);
  int (*CommInitRank) (vf_nccl_comm *, int, vf_nccl_id, int);
  int (*AllReduce) (const void *, void *, size_t, int, int, vf_nccl_comm, cudaStream_t);
  int (*Reduce) (const void *, void *, size_t, int, int, int, vf_nccl_comm, cudaStream_t);
  int (*CommDestroy) (vf_nccl_comm);
  const char *(*GetErrorString) (int);
};
static vf_nccl_api g_nccl;

struct vf_handle {
  vf_config cfg;
  int T, ntime, n_ant, nsm;
  size_t nsamp;               /* per pol per segment */
  size_t out_bytes;           /* per antenna per stream per segment */
  size_t tile_elems;          /* T*4096 per antenna */
  cudaStream_t ctl;
  cudaStream_t coadd_st;      /* co-add runs beside the next segments, not in front of them */
  cudaStream_t coadd_hi;      /* highest priority: a multi-rank co-add (vf_coadd_batch) */
  vf_slot slot[2];
  int batch_cap;              /* consecutive segments the tile / weight / mask buffers of a slot hold */
  int last_n_ant;             /* antennas of the last launch (the co-add sums that many tiles) */
  long *ant_seg;              /* per handle antenna: segments enqueued so far (position in the ring of kept tiles) */
  struct vf_ant_last { int slot; size_t idx; } *ant_last;   /* per handle antenna: slot and index (in that slot's weight /
                                 mask / tile buffers) of the last segment processed for it */
  int next_slot;              /* slot of the next synchronous segment */
  cudaEvent_t ev_k2_last;     /* completion of the most recent K2 */
  int have_k2_last;
  float2 *bp_raw, *bp_kur;    /* [n_ant][4096] (pol0, pol1) */
  float2 *tw;                 /* tw1 | tw5 | tw500 */
  float2 *tw6;                /* 6250-point FFT: tw1[250] | tw5[250] | tw250[240] | tws[6250] */
  float *wtab;                /* [26] */
  float *ave_main, *ave_raw;  /* [ave_nseg][n_ant][npol][T/8][4096] */
  float *rowok;               /* [ave_nseg][n_ant][T/8]: 1 = scrunched row of the main stream kept, 0 = zeroed (co-add count) */
  int ave_nseg;               /* tiles kept (ring over consecutive segments) */
  float *frb_delays;
  int frb_nfft_since; float frb_width, frb_amp; float frb_dm;
  double dagc[5], dagc_fb[5];
  /* timing */
  cudaEvent_t ev_t0, ev_t1;
  cudaEvent_t ev_u0, ev_u1;   /* vf_timer_begin / vf_timer_end */
  cudaEvent_t ev_ka[VF_MAX_TIMED_SEG], ev_kb[VF_MAX_TIMED_SEG], ev_kc[VF_MAX_TIMED_SEG];
  int n_timed, timed_valid;
  /* co-add */
  vf_nccl_comm comm; int nranks, rank;
  struct { cudaEvent_t ev; long lo, hi; int pending; } coadd_batch[2];   /* the last two co-add batches */
  int coadd_next;
  float *coadd_sum;           /* [ave_nseg] summed tiles, then [ave_nseg][T/8] contributing-antenna counts: one reduce */
  uint8_t *coadd_out;
  int debug_sync, serial, k2_priority;
  long long *k2_trace;        /* testing builds: vf_debug_k2_trace */
  char err[512];
};

static int vf_fail (vf_handle *h, int code, const char *fmt, ...)
{
  if (h) {
    va_list ap;
    va_start (ap, fmt);
    vsnprintf (h->err, sizeof (h->err), fmt, ap);
    va_end (ap);
  }
  return code;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
  return vf_fail (h, e_ == cudaErrorMemoryAllocation ? VF_ERR_NOMEM : VF_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                  cudaGetErrorString (e_), __FILE__, __LINE__); } while (0)

extern "C" {

const char *vf_strerror (int code)
{
  switch (code) {
    case VF_OK: return "ok";
    case VF_ERR_ARG: return "bad argument or unsupported configuration";
    case VF_ERR_CUDA: return "CUDA runtime error";
    case VF_ERR_NOMEM: return "out of memory";
    case VF_ERR_STATE: return "call out of sequence";
    case VF_ERR_VDIF: return "VDIF frames outside the segment or malformed";
    case VF_ERR_NCCL: return "NCCL error";
    case VF_ERR_NODEV: return "no usable CUDA device";
    default: return "unknown error";
  }
}

const char *vf_last_error (const vf_handle *h) { return h ? h->err : "null handle"; }

int vf_config_default (vf_config *c)
{
  if (!c) return VF_ERR_ARG;
  memset (c, 0, sizeof (*c));
  c->abi_version = VF_ABI_VERSION;
  c->nfft = VF_NFFT;             /* src/process_baseband.h:20 */
  c->nscrunch = VF_NSCRUNCH;     /* :24 */
  c->ffts_per_seg = 1024;        /* :28 */
  c->nkurto = VF_NKURTO;         /* :35 */
  c->chanmin = VF_CHANMIN;       /* :53 */
  c->chanmax = VF_CHANMAX;       /* :54 */
  c->nbit = 2;                   /* src/process_baseband.cu:34 */
  c->npol = 1;                   /* :349 */
  c->rfi_mode = 2;               /* :351 */
  c->n_antennas = 1;
  c->dag_thresh = 3.0;           /* DAG_THRESH, src/process_baseband.h:42 */
  c->min_weight = 0.2;           /* MIN_WEIGHT, :45 */
  return VF_OK;
}

/* Anscombe-Glynn constants with the reference's mixed float/double macro
 * expressions, src/pb_kernels.cu:3-20: the sample count is a float, every
 * literal a double.  out = {mu1, A, Z1, Z2, Z3}. */
static void vf_dag_constants (int nsamp, double out[5])
{
  const float n = (float) nsamp;
  const double mu1 = -6. / (n + 1);
  const double mu2 = (24. * n * (n - 2) * (n - 3)) / ((n + 1) * (n + 1) * (n + 3) * (n + 5));
  const double g1 = 6. * (n * n - 5 * n + 2) / ((n + 7) * (n + 9))
                    * sqrt ((6. * (n + 3) * (n + 5)) / (n * (n - 2) * (n - 3)));
  const double A = 6. + (8. / g1) * (2. / g1 + sqrt (1. + 4. / (g1 * g1)));
  out[0] = mu1; out[1] = A; out[2] = sqrt (4.5 * A); out[3] = 1 - 2. / (9 * A);
  out[4] = sqrt (2. / (mu2 * (A - 4)));
}

/* tile, weight and mask buffers of a slot for nb consecutive segments of n_antennas antennas */
static int vf_alloc_tiles (vf_handle *h, vf_slot *s, int nb)
{
  const size_t na = (size_t) h->n_ant * nb;
  const int mode = h->cfg.rfi_mode;
  cudaFree (s->P_raw); cudaFree (s->P_kur); cudaFree (s->w); cudaFree (s->mask);
  s->P_raw = s->P_kur = NULL; s->w = NULL; s->mask = NULL;
  if (mode != 1) CK (cudaMalloc ((void **) &s->P_raw, na * h->tile_elems * sizeof (float2)));
  if (mode != 0) CK (cudaMalloc ((void **) &s->P_kur, na * h->tile_elems * sizeof (float2)));
  CK (cudaMalloc ((void **) &s->w, na * h->T * sizeof (float)));
  CK (cudaMalloc ((void **) &s->mask, na * h->T * sizeof (uint32_t)));
  CK (cudaMemset (s->w, 0, na * h->T * sizeof (float)));
  CK (cudaMemset (s->mask, 0, na * h->T * sizeof (uint32_t)));
  return VF_OK;
}

static int vf_alloc_slot (vf_handle *h, vf_slot *s)
{
  const size_t na = (size_t) h->n_ant;
  const int mode = h->cfg.rfi_mode;
  CK (cudaStreamCreateWithFlags (&s->st, cudaStreamNonBlocking));
  {
    int lo = 0, hi = 0;
    CK (cudaDeviceGetStreamPriorityRange (&lo, &hi));
    CK (cudaStreamCreateWithPriority (&s->st_k2, cudaStreamNonBlocking, hi));
    CK (cudaEventCreateWithFlags (&s->ev_k1done, cudaEventDisableTiming));
    CK (cudaEventCreateWithFlags (&s->ev_k2done, cudaEventDisableTiming));
  }
  CK (cudaMalloc ((void **) &s->d_in, na * 2 * h->nsamp));
  int rc = vf_alloc_tiles (h, s, 1);
  if (rc) return rc;
  CK (cudaMalloc ((void **) &s->d_work, sizeof (unsigned int)));
  CK (cudaMemset (s->d_work, 0, sizeof (unsigned int)));
  s->work_base = 0;
  CK (cudaMalloc ((void **) &s->d_out_main, na * h->out_bytes));
  if (mode == 2) CK (cudaMalloc ((void **) &s->d_out_raw, na * h->out_bytes));
  CK (cudaEventCreateWithFlags (&s->ev_k2, cudaEventDisableTiming));
  CK (cudaEventCreateWithFlags (&s->ev_done, cudaEventDisableTiming));
  for (int i = 0; i < 5; ++i) CK (cudaEventCreate (&s->ev_t[i]));
  return VF_OK;
}

/* segments that one launch pair may cover (vf_process_device): the statistics dumps, the histogram
 * and the FRB injection are per segment, so they keep launches per segment */
static int vf_max_batch (const vf_handle *h)
{
  const vf_config &c = h->cfg;
  if (c.keep_stats || c.do_histo || c.inject_frb) return 1;
  int m = c.max_batch_segments > 0 ? c.max_batch_segments : 16;
  /* the tiles of a batch live in both slots: keep them under ~24 GB per slot */
  const size_t per_seg = (size_t) h->n_ant * h->tile_elems * sizeof (float2) * (c.rfi_mode == 2 ? 2 : 1);
  const size_t cap = ((size_t) 24 << 30) / (per_seg ? per_seg : 1);
  if ((size_t) m > cap) m = cap ? (int) cap : 1;
  return m;
}

static int vf_ensure_batch (vf_handle *h, int nb)
{
  if (nb <= h->batch_cap) return VF_OK;
  CK (cudaDeviceSynchronize ());
  for (int i = 0; i < 2; ++i) {
    int rc = vf_alloc_tiles (h, &h->slot[i], nb);
    if (rc) return rc;
  }
  h->batch_cap = nb;
  return VF_OK;
}

int vf_destroy (vf_handle *h)
{
  if (!h) return VF_OK;
  cudaSetDevice (h->cfg.gpu_id);
  cudaDeviceSynchronize ();
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy (h->comm);
  for (int i = 0; i < 2; ++i) {
    vf_slot *s = &h->slot[i];
    cudaFree (s->d_in); cudaFree (s->d_frames); cudaFree (s->P_raw); cudaFree (s->P_kur);
    cudaFree (s->w); cudaFree (s->mask); cudaFree (s->d_out_main); cudaFree (s->d_out_raw);
    cudaFree (s->d_bad); if (s->h_bad) cudaFreeHost (s->h_bad);
    cudaFree (s->d_in_blk); cudaFree (s->d_out_blk[0]); cudaFree (s->d_out_blk[1]);
    cudaFree (s->d_work);
    cudaFree (s->pw); cudaFree (s->pw_fb); cudaFree (s->histo);
    if (s->ev_k2) cudaEventDestroy (s->ev_k2);
    if (s->ev_done) cudaEventDestroy (s->ev_done);
    if (s->ev_k1done) cudaEventDestroy (s->ev_k1done);
    if (s->ev_k2done) cudaEventDestroy (s->ev_k2done);
    if (s->st_k2) cudaStreamDestroy (s->st_k2);
    for (int i = 0; i < 5; ++i) if (s->ev_t[i]) cudaEventDestroy (s->ev_t[i]);
    if (s->st) cudaStreamDestroy (s->st);
  }
  cudaFree (h->bp_raw); cudaFree (h->bp_kur); cudaFree (h->tw); cudaFree (h->tw6); cudaFree (h->wtab);
  cudaFree (h->ave_main); cudaFree (h->ave_raw); cudaFree (h->rowok); free (h->ant_last); free (h->ant_seg);
  cudaFree (h->frb_delays); cudaFree (h->coadd_sum); cudaFree (h->coadd_out);
  if (h->ev_t0) cudaEventDestroy (h->ev_t0);
  if (h->ev_t1) cudaEventDestroy (h->ev_t1);
  if (h->ev_u0) cudaEventDestroy (h->ev_u0);
  if (h->ev_u1) cudaEventDestroy (h->ev_u1);
  if (h->ev_k2_last) cudaEventDestroy (h->ev_k2_last);
  for (int b = 0; b < 2; ++b) if (h->coadd_batch[b].ev) cudaEventDestroy (h->coadd_batch[b].ev);
  for (int i = 0; i < VF_MAX_TIMED_SEG; ++i) {
    if (h->ev_ka[i]) cudaEventDestroy (h->ev_ka[i]);
    if (h->ev_kb[i]) cudaEventDestroy (h->ev_kb[i]);
    if (h->ev_kc[i]) cudaEventDestroy (h->ev_kc[i]);
  }
  if (h->ctl) cudaStreamDestroy (h->ctl);
  if (h->coadd_st) cudaStreamDestroy (h->coadd_st);
  if (h->coadd_hi) cudaStreamDestroy (h->coadd_hi);
  free (h);
  return VF_OK;
}

int vf_create (const vf_config *cfg, vf_handle **out)
{
  if (!cfg || !out) return VF_ERR_ARG;
  *out = NULL;
  if (cfg->abi_version != VF_ABI_VERSION) return VF_ERR_ARG;
  /* geometry the kernels are written for (reference defaults,
   * src/process_baseband.h:16-55); sanity checks of :535-538, :667-672 */
  if (cfg->nfft != VF_NFFT || cfg->nscrunch != VF_NSCRUNCH || cfg->nkurto != VF_NKURTO) return VF_ERR_ARG;
  if (cfg->chanmin != VF_CHANMIN || cfg->chanmax != VF_CHANMAX) return VF_ERR_ARG;
  if (cfg->ffts_per_seg <= 0 || cfg->ffts_per_seg % VF_NSCRUNCH || cfg->ffts_per_seg > 8192) return VF_ERR_ARG;
  if (!(cfg->nbit == 2 || cfg->nbit == 4 || cfg->nbit == 8)) return VF_ERR_ARG;
  if (!(cfg->npol == 1 || cfg->npol == 2)) return VF_ERR_ARG;
  if (cfg->rfi_mode < 0 || cfg->rfi_mode > 2) return VF_ERR_ARG;
  if (cfg->n_antennas < 1 || cfg->n_antennas > 4096) return VF_ERR_ARG;
  if (cfg->max_batch_segments < 0 || cfg->max_batch_segments > 64) return VF_ERR_ARG;
#ifdef VF_TESTING
  if (!(cfg->k1_threads == 0 || cfg->k1_threads == 1 || cfg->k1_threads == 320 || cfg->k1_threads == 512 || cfg->k1_threads == 640)) return VF_ERR_ARG;
#else
  if (cfg->k1_threads != 0) return VF_ERR_ARG;      /* the monolithic channeliser exists in testing builds only */
#endif
  if (!(cfg->dag_thresh > 0.) || !(cfg->min_weight >= 0.) || cfg->min_weight > 1.) return VF_ERR_ARG;

  int ndev = 0;
  if (cudaGetDeviceCount (&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError (); return VF_ERR_NODEV; }
  if (cfg->gpu_id < 0 || cfg->gpu_id >= ndev) return VF_ERR_NODEV;

  vf_handle *h = (vf_handle *) calloc (1, sizeof (*h));
  if (!h) return VF_ERR_NOMEM;
  h->cfg = *cfg;
  *out = h;    /* handed back even on failure so that vf_last_error works; caller destroys */
  h->T = cfg->ffts_per_seg;
  h->batch_cap = 1; h->last_n_ant = 1;
  h->ntime = h->T / VF_NSCRUNCH;
  h->n_ant = cfg->n_antennas;
  h->nsamp = (size_t) h->T * VF_NFFT;
  h->out_bytes = (size_t) h->ntime * cfg->npol * VF_NCHANOUT * cfg->nbit / 8;
  h->tile_elems = (size_t) h->T * VF_NCHANOUT;
  h->frb_nfft_since = -1;
  { const char *e = getenv ("VF_DEBUG_SYNC"); h->debug_sync = e && *e == '1'; }
  /* VF_SERIAL=1: no overlap between segments, so that the per-kernel event times of
   * vf_last_elapsed_ms are pure execution times (profiling aid) */
  { const char *e = getenv ("VF_SERIAL"); h->serial = e && *e == '1'; }
  { const char *e = getenv ("VF_K2_PRIORITY"); h->k2_priority = !(e && *e == '0'); }     /* on unless VF_K2_PRIORITY=0 (A/B) */

  CK (cudaSetDevice (cfg->gpu_id));
  if (cfg->numa_pin) vf_bind_thread_to_gpu (cfg->gpu_id, NULL, 0);
  cudaDeviceProp prop;
  CK (cudaGetDeviceProperties (&prop, cfg->gpu_id));
  if (prop.major < 10)
    return vf_fail (h, VF_ERR_NODEV, "device %d is sm_%d%d; this library is built for sm_100a only",
                    cfg->gpu_id, prop.major, prop.minor);
  h->nsm = prop.multiProcessorCount;
  CK (vf_k1_configure ());
  CK (cudaStreamCreateWithFlags (&h->ctl, cudaStreamNonBlocking));
  CK (cudaStreamCreateWithFlags (&h->coadd_st, cudaStreamNonBlocking));
  {
    int lo = 0, hi = 0;
    CK (cudaDeviceGetStreamPriorityRange (&lo, &hi));
    { const char *e = getenv ("VF_COADD_PRIORITY"); if (e && *e == '0') hi = lo; }      /* A/B */
    CK (cudaStreamCreateWithPriority (&h->coadd_hi, cudaStreamNonBlocking, hi));
  }
  for (int i = 0; i < 2; ++i) {
    int rc = vf_alloc_slot (h, &h->slot[i]);
    if (rc) return rc;
  }
  const size_t na = (size_t) h->n_ant;
  CK (cudaMalloc ((void **) &h->bp_raw, na * VF_NCHANOUT * sizeof (float2)));
  CK (cudaMemset (h->bp_raw, 0, na * VF_NCHANOUT * sizeof (float2)));
  if (cfg->rfi_mode == 2) {
    CK (cudaMalloc ((void **) &h->bp_kur, na * VF_NCHANOUT * sizeof (float2)));
    CK (cudaMemset (h->bp_kur, 0, na * VF_NCHANOUT * sizeof (float2)));
  }

  /* FFT twiddles in double, rounded once to float: w_12500^p, w_12500^(5p), w_500^p (p < 500) */
  {
    std::vector<float2> tw (1500);
    for (int p = 0; p < 500; ++p) {
      const double a1 = -2.0 * M_PI * p / 12500.0, a5 = -2.0 * M_PI * 5 * p / 12500.0;
      tw[p] = make_float2 ((float) cos (a1), (float) sin (a1));
      tw[500 + p] = make_float2 ((float) cos (a5), (float) sin (a5));
      tw[1000 + p] = make_float2 (1.f, 0.f);
    }
    for (int k = 1; k < 25; ++k)                 /* tw500[(k - 1) * 20 + p'] = w_500^(p' k) */
      for (int pp = 0; pp < 20; ++pp) {
        const double a = -2.0 * M_PI * (pp * k) / 500.0;
        tw[1000 + (k - 1) * 20 + pp] = make_float2 ((float) cos (a), (float) sin (a));
      }
    CK (cudaMalloc ((void **) &h->tw, 1500 * sizeof (float2)));
    CK (cudaMemcpy (h->tw, tw.data (), 1500 * sizeof (float2), cudaMemcpyHostToDevice));
  }
  /* the same for the 6250-point real-input FFT (vf_fft6250.cuh): w_6250^p, w_6250^(5p) (p < 250), w_250^(p' k)
   * and the split twiddles w_12500^k (k < 6250) */
  {
    std::vector<float2> tw (VF_TW6_LEN);
    for (int p = 0; p < 250; ++p) {
      const double a1 = -2.0 * M_PI * p / 6250.0, a5 = -2.0 * M_PI * 5 * p / 6250.0;
      tw[p] = make_float2 ((float) cos (a1), (float) sin (a1));
      tw[250 + p] = make_float2 ((float) cos (a5), (float) sin (a5));
    }
    for (int k = 1; k < 25; ++k)
      for (int pp = 0; pp < 10; ++pp) {
        const double a = -2.0 * M_PI * (pp * k) / 250.0;
        tw[500 + (k - 1) * 10 + pp] = make_float2 ((float) cos (a), (float) sin (a));
      }
    for (int k = 0; k < VF6_M; ++k) {
      const double a = -2.0 * M_PI * k / 12500.0;
      tw[740 + k] = make_float2 ((float) cos (a), (float) sin (a));
    }
    CK (cudaMalloc ((void **) &h->tw6, VF_TW6_LEN * sizeof (float2)));
    CK (cudaMemcpy (h->tw6, tw.data (), VF_TW6_LEN * sizeof (float2), cudaMemcpyHostToDevice));
  }
  /* weight of an FFT block with k kept sub-blocks: k sequential float adds of
   * float(NKURTO)/NFFT (atomicAdd, src/pb_kernels.cu:292) */
  {
    float wt[VF_NSUB + 1];
    const float inc = (float) VF_NKURTO / VF_NFFT;
    float acc = 0.f;
    wt[0] = 0.f;
    for (int k = 1; k <= VF_NSUB; ++k) { acc += inc; wt[k] = acc; }
    CK (cudaMalloc ((void **) &h->wtab, sizeof (wt)));
    CK (cudaMemcpy (h->wtab, wt, sizeof (wt), cudaMemcpyHostToDevice));
  }
  vf_dag_constants (VF_NKURTO, h->dagc);
  vf_dag_constants (VF_NFFT, h->dagc_fb);

  for (int i = 0; i < 2; ++i) {
    vf_slot *s = &h->slot[i];
    if (cfg->keep_stats && cfg->rfi_mode) {
      const size_t nblk = (size_t) h->T * VF_NSUB;
      CK (cudaMalloc ((void **) &s->pw, na * 6 * nblk * sizeof (float)));
      s->kur = s->pw + na * 2 * nblk;
      s->dag = s->pw + na * 4 * nblk;
      CK (cudaMalloc ((void **) &s->pw_fb, na * 6 * h->T * sizeof (float)));
      s->kur_fb = s->pw_fb + na * 2 * h->T;
      s->dag_fb = s->pw_fb + na * 4 * h->T;
    }
    if (cfg->do_histo) CK (cudaMalloc ((void **) &s->histo, na * 512 * sizeof (unsigned int)));
  }
  h->ant_last = (vf_handle::vf_ant_last *) calloc (na, sizeof (*h->ant_last));
  h->ant_seg = (long *) calloc (na, sizeof (long));
  if (!h->ant_last || !h->ant_seg) return vf_fail (h, VF_ERR_NOMEM, "out of host memory");
  for (int a = 0; a < h->n_ant; ++a) h->ant_last[a].idx = (size_t) a;
  h->ave_nseg = cfg->power_segments > 0 ? cfg->power_segments : 1;
  if (cfg->keep_power) {
    const size_t n = (size_t) h->ave_nseg * na * cfg->npol * h->ntime * VF_NCHANOUT;
    CK (cudaMalloc ((void **) &h->ave_main, n * sizeof (float)));
    CK (cudaMemset (h->ave_main, 0, n * sizeof (float)));
    if (cfg->rfi_mode == 2) {
      CK (cudaMalloc ((void **) &h->ave_raw, n * sizeof (float)));
      CK (cudaMemset (h->ave_raw, 0, n * sizeof (float)));
    }
    const size_t nr = (size_t) h->ave_nseg * na * h->ntime;
    CK (cudaMalloc ((void **) &h->rowok, nr * sizeof (float)));
    CK (cudaMemset (h->rowok, 0, nr * sizeof (float)));
  }
  CK (cudaEventCreate (&h->ev_t0));
  CK (cudaEventCreate (&h->ev_t1));
  CK (cudaEventCreate (&h->ev_u0));
  CK (cudaEventCreate (&h->ev_u1));
  CK (cudaEventCreateWithFlags (&h->ev_k2_last, cudaEventDisableTiming));
  CK (cudaDeviceSynchronize ());
  return VF_OK;
}

/* Bind the calling thread to the CPUs that are local to the GPU (sysfs local_cpulist of its PCI function), so that
 * pinned buffers the thread allocates afterwards (first touch) and its copies' staging sit on the GPU's NUMA node.
 * cpulist (optional) receives the list that was applied, "" when the platform exposes none. */
int vf_bind_thread_to_gpu (int gpu_id, char *cpulist, size_t cap)
{
  if (cpulist && cap) cpulist[0] = 0;
  char bus[32];
  if (cudaDeviceGetPCIBusId (bus, sizeof (bus), gpu_id) != cudaSuccess) { cudaGetLastError (); return VF_ERR_NODEV; }
  for (char *c = bus; *c; ++c) *c = (char) tolower (*c);
  char path[128], line[1024];
  snprintf (path, sizeof (path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
  FILE *f = fopen (path, "r");
  if (!f) return VF_OK;                       /* no NUMA information: nothing to do */
  if (!fgets (line, sizeof (line), f)) line[0] = 0;
  fclose (f);
  cpu_set_t set;
  CPU_ZERO (&set);
  int n = 0;
  for (char *p = line; *p && *p != '\n';) {
    char *e;
    long a = strtol (p, &e, 10), b = a;
    if (e == p) break;
    if (*e == '-') { p = e + 1; b = strtol (p, &e, 10); }
    for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET ((int) c, &set); ++n; }
    p = (*e == ',') ? e + 1 : e;
  }
  if (n == 0) return VF_OK;
  /* keep only CPUs this process may use (cgroup / taskset) */
  cpu_set_t cur;
  if (sched_getaffinity (0, sizeof (cur), &cur) == 0) {
    cpu_set_t both;
    CPU_AND (&both, &set, &cur);
    if (CPU_COUNT (&both) == 0) return VF_OK;
    set = both;
  }
  if (sched_setaffinity (0, sizeof (set), &set) != 0) return VF_OK;
  if (cpulist && cap) { size_t l = strcspn (line, "\n"); if (l >= cap) l = cap - 1; memcpy (cpulist, line, l); cpulist[l] = 0; }
  return VF_OK;
}

size_t vf_segment_out_bytes (const vf_handle *h) { return h ? h->out_bytes : 0; }
size_t vf_segment_in_samples (const vf_handle *h) { return h ? h->nsamp : 0; }

int vf_host_alloc (void **p, size_t bytes)
{
  if (!p) return VF_ERR_ARG;
  cudaError_t e = cudaMallocHost (p, bytes);
  return e == cudaSuccess ? VF_OK : (e == cudaErrorMemoryAllocation ? VF_ERR_NOMEM : VF_ERR_CUDA);
}

int vf_host_free (void *p) { return cudaFreeHost (p) == cudaSuccess ? VF_OK : VF_ERR_CUDA; }

int vf_host_register (void *p, size_t bytes)
{
  if (!p || !bytes) return VF_ERR_ARG;
  cudaError_t e = cudaHostRegister (p, bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) cudaGetLastError ();
  return e == cudaSuccess ? VF_OK : (e == cudaErrorMemoryAllocation ? VF_ERR_NOMEM : VF_ERR_CUDA);
}

int vf_host_unregister (void *p)
{
  cudaError_t e = cudaHostUnregister (p);
  if (e != cudaSuccess) cudaGetLastError ();
  return e == cudaSuccess ? VF_OK : VF_ERR_CUDA;
}

/* ---- the launch sequence of one segment on one slot ---------------------- *
 * d_in: [n_ant][2][T*12500] on the device, antenna a of the launch being antenna ant0 + a of the handle
 * (bandpass state, statistics, kept tiles).  Outputs to d_main / d_raw ([n_ant][out_bytes]).
 * timed >= 0: record the K1/K2 events of that index; timed == -2: the slot's own events (asynchronous submissions). */
static int vf_enqueue_segment (vf_handle *h, vf_slot *s, int ant0, int n_ant, const uint8_t *d_in,
                               uint8_t *d_main, uint8_t *d_raw, int timed, int n_seg = 1)
{
  if (n_seg > h->batch_cap) return vf_fail (h, VF_ERR_STATE, "batch of %d segments, buffers hold %d", n_seg, h->batch_cap);
  const vf_config &c = h->cfg;
  const size_t nblk = (size_t) h->T * VF_NSUB;
  vf_k1_params k1;
  memset (&k1, 0, sizeof (k1));
  k1.in = d_in;
  k1.pol_stride = h->nsamp;
  k1.ant_stride = 2 * h->nsamp;
  k1.T = h->T; k1.n_ant = n_ant * n_seg; k1.rfi_mode = c.rfi_mode;   /* (segment, antenna) pairs are the channeliser's antennas */
  /* place of this launch in the slot's tile / weight / mask buffers: single-segment launches of a part of the
   * antennas sit at their handle index, so that the getters find every antenna's last segment */
  const size_t base = (n_seg == 1) ? (size_t) ant0 : 0;      /* a batched launch fills the buffers from the start */
  k1.P_raw = s->P_raw ? s->P_raw + base * h->tile_elems : NULL; k1.P_kur = s->P_kur ? s->P_kur + base * h->tile_elems : NULL;
  k1.w = s->w + base * h->T; k1.mask = s->mask + base * h->T;
  if (s->pw) {
    k1.pw = s->pw + (size_t) ant0 * 2 * nblk; k1.kur = s->kur + (size_t) ant0 * 2 * nblk; k1.dag = s->dag + (size_t) ant0 * 2 * nblk;
    k1.pw_fb = s->pw_fb + (size_t) ant0 * 2 * h->T; k1.kur_fb = s->kur_fb + (size_t) ant0 * 2 * h->T;
    k1.dag_fb = s->dag_fb + (size_t) ant0 * 2 * h->T;
  }
  if (s->histo) k1.histo = s->histo + (size_t) ant0 * 512;
  k1.tb.tw1 = h->tw; k1.tb.tw5 = h->tw + 500; k1.tb.tw500 = h->tw + 1000;
  k1.tb6.tw1 = h->tw6; k1.tb6.tw5 = h->tw6 + 250; k1.tb6.tw250 = h->tw6 + 500; k1.tb6.tws = h->tw6 + 740;
  memcpy (k1.dagc, h->dagc, sizeof (k1.dagc));
  memcpy (k1.dagc_fb, h->dagc_fb, sizeof (k1.dagc_fb));
  k1.dag_thresh = c.dag_thresh;
  k1.wtab = h->wtab;
  if (c.inject_frb && h->frb_nfft_since >= 0 && h->frb_delays) {
    k1.frb_delays = h->frb_delays;
    k1.nfft_since_frb = h->frb_nfft_since;
    k1.frb_width = h->frb_width; k1.frb_amp = h->frb_amp;
  }
  if (k1.histo) CK (cudaMemsetAsync (k1.histo, 0, (size_t) n_ant * 512 * sizeof (unsigned int), s->st));
  const int n_items = n_ant * n_seg * h->T;
  const int grid = n_items < h->nsm ? n_items : h->nsm;
  const int threads = c.k1_threads;            /* 0: the product kernel; testing builds, A/B: 1 = round 1's two-for-one pipelined kernel, 320/512/640 = monolithic kernel */
  if (threads == 0 || threads == 1) {
    k1.work_counter = s->d_work;
    k1.work_base = s->work_base;
    s->work_base += (unsigned int) (n_items + 2 * grid);   /* what this launch draws (wraps with the counter) */
  }
  if (h->serial && h->have_k2_last) CK (cudaStreamWaitEvent (s->st, h->ev_k2_last, 0));
  if (timed >= 0) CK (cudaEventRecord (h->ev_ka[timed], s->st));
  if (timed == -2) CK (cudaEventRecord (s->ev_t[1], s->st));
  CK (vf_launch_k1 (k1, grid, threads, s->st));
  if (h->debug_sync) CK (cudaStreamSynchronize (s->st));      /* VF_DEBUG_SYNC=1: attribute faults to a kernel */
  if (timed >= 0) CK (cudaEventRecord (h->ev_kb[timed], s->st));
  if (timed == -2) CK (cudaEventRecord (s->ev_t[2], s->st));

  /* The normaliser goes to a stream of its own with the highest priority: the channeliser of the NEXT submission
   * (other slot) is usually queued already, and its persistent CTAs hold an SM each for the whole launch -- the block
   * scheduler must hand the SMs that this submission's channeliser frees to this normaliser first (one wave of 128
   * CTAs), and the next channeliser then starts on the 20 SMs that wave leaves idle.  Measured (round 2, bench, 1 antenna):
   * 1118-1138 -> 1153 antenna-seconds/s. */
  cudaStream_t k2st = h->k2_priority ? s->st_k2 : s->st;
  if (h->k2_priority) {
    CK (cudaEventRecord (s->ev_k1done, s->st));
    CK (cudaStreamWaitEvent (k2st, s->ev_k1done, 0));
  }
  /* the bandpass makes K2 launches sequential in segment order */
  if (h->have_k2_last) CK (cudaStreamWaitEvent (k2st, h->ev_k2_last, 0));
  /* segments enqueued so far for these antennas (they advance together): index into the ring of kept tiles */
  const long seg0 = h->ant_seg[ant0];
  /* this launch overwrites the tiles of segments [seg0 - ave_nseg, seg0 - ave_nseg + n_seg): wait for a co-add that still reads them */
  for (int b = 0; b < 2; ++b) {
    const long victim_lo = seg0 - h->ave_nseg, victim_hi = victim_lo + n_seg;    /* [lo, hi) */
    if (h->coadd_batch[b].pending && victim_lo < h->coadd_batch[b].hi && victim_hi > h->coadd_batch[b].lo) {
      CK (cudaStreamWaitEvent (k2st, h->coadd_batch[b].ev, 0));
      h->coadd_batch[b].pending = 0;
    }
  }
  vf_k2_params k2;
  memset (&k2, 0, sizeof (k2));
  k2.P_raw = k1.P_raw; k2.P_kur = k1.P_kur; k2.w = k1.w; k2.mask = k1.mask;
  k2.bp_raw = h->bp_raw + (size_t) ant0 * VF_NCHANOUT;
  k2.bp_kur = ((c.rfi_mode == 2) ? h->bp_kur : h->bp_raw) + (size_t) ant0 * VF_NCHANOUT;
  k2.T = h->T; k2.n_ant = n_ant; k2.n_seg = n_seg; k2.rfi_mode = c.rfi_mode; k2.npol = c.npol; k2.nbit = c.nbit;
  /* bp_scale = float(tsamp/tsmooth), src/process_baseband.cu:737-741 */
  k2.bp_scale = (float) (((double) VF_NFFT / 128000000 * VF_NSCRUNCH) / 1.0);
  k2.min_weight = c.min_weight;
#ifdef VF_TESTING
  k2.trace = h->k2_trace;
#endif
  {
    /* the reference compares the float weight with the double MIN_WEIGHT (:537-538, :616-617); for a float w that
     * is w >= (smallest float >= MIN_WEIGHT) */
    float f = (float) c.min_weight;
    if ((double) f < c.min_weight) f = nextafterf (f, INFINITY);
    k2.min_weight_f = f;
  }
  k2.out_main = d_main; k2.out_raw = d_raw; k2.out_stride = h->out_bytes;
  {
    /* tile slots of these segments in the ring of kept tiles (sized for the handle's n_antennas) */
    const size_t tile1 = (size_t) c.npol * h->ntime * VF_NCHANOUT;
    k2.ave_main = h->ave_main ? h->ave_main + (size_t) ant0 * tile1 : NULL;
    k2.ave_raw = h->ave_raw ? h->ave_raw + (size_t) ant0 * tile1 : NULL;
    k2.ave_seg0 = seg0; k2.ave_nseg = h->ave_nseg;
    k2.ave_seg_elems = (size_t) h->n_ant * tile1;
    k2.rowok = h->rowok ? h->rowok + (size_t) ant0 * h->ntime : NULL;
    k2.rowok_seg_elems = (size_t) h->n_ant * h->ntime;
    for (int a = 0; a < n_ant; ++a) {
      h->ant_seg[ant0 + a] = seg0 + n_seg;
      h->ant_last[ant0 + a].slot = (int) (s - h->slot);
      h->ant_last[ant0 + a].idx = base + (size_t) (n_seg - 1) * n_ant + a;
    }
    h->last_n_ant = n_ant;
  }
  CK (vf_launch_k2 (k2, k2st));
  CK (cudaEventRecord (h->ev_k2_last, k2st));
  if (h->k2_priority) {
    CK (cudaEventRecord (s->ev_k2done, k2st));
    CK (cudaStreamWaitEvent (s->st, s->ev_k2done, 0));      /* copies out, and the slot's next channeliser, follow on the slot's stream */
  }
  if (h->debug_sync) CK (cudaStreamSynchronize (s->st));
  if (timed >= 0) CK (cudaEventRecord (h->ev_kc[timed], s->st));
  if (timed == -2) CK (cudaEventRecord (s->ev_t[3], s->st));
  h->have_k2_last = 1;
  return VF_OK;
}

static int vf_timing_begin (vf_handle *h, int n_timed)
{
  if (n_timed > VF_MAX_TIMED_SEG) n_timed = 0;   /* too many to time per kernel: total only */
  for (int i = 0; i < n_timed; ++i)
    if (!h->ev_ka[i]) {
      CK (cudaEventCreate (&h->ev_ka[i]));
      CK (cudaEventCreate (&h->ev_kb[i]));
      CK (cudaEventCreate (&h->ev_kc[i]));
    }
  h->n_timed = n_timed;
  h->timed_valid = 0;
  /* the call's clock starts where its first launch does: on the stream of the slot it uses first.  (No fork from the
   * control stream: that made every call wait for the whole of the previous one -- the control stream had joined both
   * slots -- so the channeliser of a call never ran beside the normaliser of the one before.) */
  CK (cudaEventRecord (h->ev_t0, h->slot[h->next_slot].st));
  return VF_OK;
}

static int vf_timing_end (vf_handle *h)
{
  /* join */
  for (int i = 0; i < 2; ++i) {
    CK (cudaEventRecord (h->slot[i].ev_done, h->slot[i].st));
    CK (cudaStreamWaitEvent (h->ctl, h->slot[i].ev_done, 0));
  }
  CK (cudaEventRecord (h->ev_t1, h->ctl));
  h->timed_valid = 1;
  return VF_OK;
}

static int vf_check_batch (vf_handle *h, int n_ant, size_t nsamp_per_pol)
{
  if (!h) return VF_ERR_ARG;
  if (n_ant < 1 || n_ant > h->n_ant) return vf_fail (h, VF_ERR_ARG, "n_ant %d outside 1..%d", n_ant, h->n_ant);
  if (nsamp_per_pol != h->nsamp)
    return vf_fail (h, VF_ERR_ARG, "nsamp_per_pol %zu != ffts_per_seg*12500 = %zu", nsamp_per_pol, h->nsamp);
  return VF_OK;
}

int vf_submit_async (vf_handle *h, int slot, int n_ant,
                     const uint8_t *const *pol0, const uint8_t *const *pol1, size_t nsamp_per_pol,
                     uint8_t *const *fb_main, uint8_t *const *fb_raw)
{
  int rc = vf_check_batch (h, n_ant, nsamp_per_pol);
  if (rc) return rc;
  if (slot < 0 || slot > 1 || !pol0 || !pol1 || !fb_main) return vf_fail (h, VF_ERR_ARG, "bad slot or null array");
  vf_slot *s = &h->slot[slot];
  if (s->pending) return vf_fail (h, VF_ERR_STATE, "slot %d submitted twice without vf_wait", slot);
  CK (cudaSetDevice (h->cfg.gpu_id));
  CK (cudaEventRecord (s->ev_t[0], s->st));
  for (int a = 0; a < n_ant; ++a) {                       /* H2D, src/process_baseband.cu:1117-1122 */
    if (!pol0[a] || !pol1[a] || !fb_main[a]) return vf_fail (h, VF_ERR_ARG, "null buffer for antenna %d", a);
    CK (cudaMemcpyAsync (s->d_in + (size_t) a * 2 * h->nsamp, pol0[a], h->nsamp, cudaMemcpyHostToDevice, s->st));
    CK (cudaMemcpyAsync (s->d_in + ((size_t) a * 2 + 1) * h->nsamp, pol1[a], h->nsamp, cudaMemcpyHostToDevice, s->st));
  }
  rc = vf_enqueue_segment (h, s, 0, n_ant, s->d_in, s->d_out_main, s->d_out_raw, -2);
  if (rc) return rc;
  for (int a = 0; a < n_ant; ++a) {                       /* D2H, :1370-1375 */
    CK (cudaMemcpyAsync (fb_main[a], s->d_out_main + (size_t) a * h->out_bytes, h->out_bytes, cudaMemcpyDeviceToHost, s->st));
    if (h->cfg.rfi_mode == 2 && fb_raw && fb_raw[a])
      CK (cudaMemcpyAsync (fb_raw[a], s->d_out_raw + (size_t) a * h->out_bytes, h->out_bytes, cudaMemcpyDeviceToHost, s->st));
  }
  CK (cudaEventRecord (s->ev_t[4], s->st));
  CK (cudaEventRecord (s->ev_done, s->st));
  s->t_valid = 1;
  s->pending = 1;
  return VF_OK;
}

static int vf_slot_reserve_blk (vf_handle *h, vf_slot *s, int units);

/* n_seg consecutive segments of n_ant antennas from ONE host buffer laid out as the device wants it,
 * in [n_seg][n_ant][2][nsamp] (for one antenna: the ten 100-ms segments of a second back to back, each pol 0 then
 * pol 1): one copy in, one launch pair over the block, one copy out per stream into fb_main / fb_raw
 * [n_seg][n_ant][out_bytes].  The asynchronous form of vf_process_device for host buffers; vf_wait (slot) completes it.
 * Replaces n_seg rounds of the reference's per-segment H2D / kernels / D2H (src/process_baseband.cu:1108-1375). */
int vf_submit_block_async (vf_handle *h, int slot, int n_ant, int n_seg, const uint8_t *in, uint8_t *fb_main, uint8_t *fb_raw)
{
  if (!h || !in || !fb_main) return VF_ERR_ARG;
  if (slot < 0 || slot > 1) return vf_fail (h, VF_ERR_ARG, "bad slot");
  if (n_ant < 1 || n_ant > h->n_ant) return vf_fail (h, VF_ERR_ARG, "n_ant %d outside 1..%d", n_ant, h->n_ant);
  if (n_seg < 1 || n_seg > vf_max_batch (h))
    return vf_fail (h, VF_ERR_ARG, "n_seg %d outside 1..%d (max_batch_segments; 1 with keep_stats / do_histo / inject_frb)", n_seg, vf_max_batch (h));
  vf_slot *s = &h->slot[slot];
  if (s->pending) return vf_fail (h, VF_ERR_STATE, "slot %d submitted twice without vf_wait", slot);
  CK (cudaSetDevice (h->cfg.gpu_id));
  int rc = vf_ensure_batch (h, n_seg);
  if (rc) return rc;
  rc = vf_slot_reserve_blk (h, s, n_seg * n_ant);
  if (rc) return rc;
  const size_t units = (size_t) n_seg * n_ant;
  CK (cudaEventRecord (s->ev_t[0], s->st));
  CK (cudaMemcpyAsync (s->d_in_blk, in, units * 2 * h->nsamp, cudaMemcpyHostToDevice, s->st));       /* :1117-1122 */
  rc = vf_enqueue_segment (h, s, 0, n_ant, s->d_in_blk, s->d_out_blk[0], s->d_out_blk[1], -2, n_seg);
  if (rc) return rc;
  CK (cudaMemcpyAsync (fb_main, s->d_out_blk[0], units * h->out_bytes, cudaMemcpyDeviceToHost, s->st));  /* :1370-1375 */
  if (h->cfg.rfi_mode == 2 && fb_raw)
    CK (cudaMemcpyAsync (fb_raw, s->d_out_blk[1], units * h->out_bytes, cudaMemcpyDeviceToHost, s->st));
  CK (cudaEventRecord (s->ev_t[4], s->st));
  CK (cudaEventRecord (s->ev_done, s->st));
  s->t_valid = 1;
  s->pending = 1;
  return VF_OK;
}

int vf_wait (vf_handle *h, int slot)
{
  if (!h || slot < 0 || slot > 1) return VF_ERR_ARG;
  vf_slot *s = &h->slot[slot];
  if (!s->pending) return vf_fail (h, VF_ERR_STATE, "vf_wait (%d) without vf_submit_async", slot);
  const int kind = s->pending;
  s->pending = 0;
  CK (cudaEventSynchronize (s->ev_done));
  if (kind == 2 && (s->h_bad[0] || s->h_bad[1]))
    return vf_fail (h, VF_ERR_VDIF, "%u frame(s) outside the block starting at frame %u, %u of another second: skipped (outputs complete)",
                    s->h_bad[0], s->first_frame, s->h_bad[1]);
  return VF_OK;
}

int vf_process_batch (vf_handle *h, int n_ant,
                      const uint8_t *const *pol0, const uint8_t *const *pol1, size_t nsamp_per_pol,
                      uint8_t *const *fb_main, uint8_t *const *fb_raw)
{
  if (!h) return VF_ERR_ARG;
  const int slot = h->next_slot;
  if (h->slot[slot].pending) return vf_fail (h, VF_ERR_STATE, "slot %d has an asynchronous segment in flight", slot);
  int rc = vf_timing_begin (h, 0);
  if (rc) return rc;
  rc = vf_submit_async (h, slot, n_ant, pol0, pol1, nsamp_per_pol, fb_main, fb_raw);
  if (rc) return rc;
  rc = vf_timing_end (h);
  if (rc) { h->slot[slot].pending = 0; return rc; }
  h->next_slot ^= 1;
  return vf_wait (h, slot);
}

/* one antenna of a multi-antenna handle: the same launch sequence with that antenna's bandpass,
 * statistics and kept tiles */
static int vf_submit_one_async (vf_handle *h, int slot, int antenna, const uint8_t *pol0, const uint8_t *pol1,
                                uint8_t *fb_main, uint8_t *fb_raw)
{
  vf_slot *s = &h->slot[slot];
  CK (cudaSetDevice (h->cfg.gpu_id));
  CK (cudaEventRecord (s->ev_t[0], s->st));
  CK (cudaMemcpyAsync (s->d_in, pol0, h->nsamp, cudaMemcpyHostToDevice, s->st));      /* H2D, src/process_baseband.cu:1117-1122 */
  CK (cudaMemcpyAsync (s->d_in + h->nsamp, pol1, h->nsamp, cudaMemcpyHostToDevice, s->st));
  int rc = vf_enqueue_segment (h, s, antenna, 1, s->d_in, s->d_out_main, s->d_out_raw, -2);
  if (rc) return rc;
  CK (cudaMemcpyAsync (fb_main, s->d_out_main, h->out_bytes, cudaMemcpyDeviceToHost, s->st));   /* D2H, :1370-1375 */
  if (h->cfg.rfi_mode == 2 && fb_raw)
    CK (cudaMemcpyAsync (fb_raw, s->d_out_raw, h->out_bytes, cudaMemcpyDeviceToHost, s->st));
  CK (cudaEventRecord (s->ev_t[4], s->st));
  CK (cudaEventRecord (s->ev_done, s->st));
  s->t_valid = 1;
  s->pending = 1;
  return VF_OK;
}

int vf_process_segment (vf_handle *h, int antenna,
                        const uint8_t *pol0, const uint8_t *pol1, size_t nsamp_per_pol,
                        uint8_t *fb_main, uint8_t *fb_raw, size_t *nbytes)
{
  int rc = vf_check_batch (h, 1, nsamp_per_pol);
  if (rc) return rc;
  if (antenna < 0 || antenna >= h->n_ant) return vf_fail (h, VF_ERR_ARG, "antenna %d outside 0..%d", antenna, h->n_ant - 1);
  if (!pol0 || !pol1 || !fb_main) return vf_fail (h, VF_ERR_ARG, "null buffer");
  const int slot = h->next_slot;
  if (h->slot[slot].pending) return vf_fail (h, VF_ERR_STATE, "slot %d has an asynchronous segment in flight", slot);
  rc = vf_timing_begin (h, 0);
  if (rc) return rc;
  rc = vf_submit_one_async (h, slot, antenna, pol0, pol1, fb_main, fb_raw);
  if (rc) return rc;
  rc = vf_timing_end (h);
  if (rc) { h->slot[slot].pending = 0; return rc; }
  h->next_slot ^= 1;
  rc = vf_wait (h, slot);
  if (rc == VF_OK && nbytes) *nbytes = h->out_bytes;
  return rc;
}

/* planar input and outputs of a block on slot s: units = segments x antennas */
static int vf_slot_reserve_blk (vf_handle *h, vf_slot *s, int units)
{
  if (s->blk_cap >= units) return VF_OK;
  cudaFree (s->d_in_blk); cudaFree (s->d_out_blk[0]); cudaFree (s->d_out_blk[1]);
  s->d_in_blk = NULL; s->d_out_blk[0] = s->d_out_blk[1] = NULL; s->blk_cap = 0;
  CK (cudaMalloc ((void **) &s->d_in_blk, (size_t) units * 2 * h->nsamp));
  CK (cudaMalloc ((void **) &s->d_out_blk[0], (size_t) units * h->out_bytes));
  if (h->cfg.rfi_mode == 2) CK (cudaMalloc ((void **) &s->d_out_blk[1], (size_t) units * h->out_bytes));
  s->blk_cap = units;
  return VF_OK;
}

/* device buffers of a VDIF submission of n_seg segments on slot s: the frames as they arrive, and for n_seg > 1 the
 * planar input and the outputs of the block (a single segment uses the slot's own) */
static int vf_slot_reserve_vdif (vf_handle *h, vf_slot *s, size_t bytes, int n_seg)
{
  const size_t per_pol = h->nsamp / VF_VD_DAT;
  if (s->frames_cap < bytes) {
    cudaFree (s->d_frames); s->d_frames = NULL; s->frames_cap = 0;
    size_t cap = bytes > 2 * per_pol * n_seg * VF_VD_FRM ? bytes : 2 * per_pol * n_seg * VF_VD_FRM;
    CK (cudaMalloc ((void **) &s->d_frames, cap));
    s->frames_cap = cap;
  }
  if (!s->d_bad) {
    CK (cudaMalloc ((void **) &s->d_bad, 4 * sizeof (unsigned int)));
    CK (cudaMallocHost ((void **) &s->h_bad, 4 * sizeof (unsigned int)));
  }
  if (n_seg > 1) {
    /* statistics dumps, histogram and FRB injection are per segment (vf_max_batch) */
    if (vf_max_batch (h) < n_seg) return vf_fail (h, VF_ERR_STATE, "blocks of %d segments need a handle without keep_stats / do_histo / inject_frb", n_seg);
    int rc = vf_ensure_batch (h, n_seg);
    if (rc) return rc;
    rc = vf_slot_reserve_blk (h, s, n_seg);
    if (rc) return rc;
  }
  return VF_OK;
}

/* Everything vf_submit_vdif_block_async (n_seg) would otherwise allocate on its first call, for both slots, and the
 * kernels' code on the device: called once at start-up (the reference allocates at start-up too,
 * src/process_baseband.cu:572-690) so that the first second of an observation costs what the others do. */
int vf_reserve_vdif_blocks (vf_handle *h, int n_seg)
{
  if (!h) return VF_ERR_ARG;
  if (n_seg < 1 || n_seg > 16) return vf_fail (h, VF_ERR_ARG, "n_seg %d outside 1..16", n_seg);
  if (h->nsamp % VF_VD_DAT) return vf_fail (h, VF_ERR_ARG, "the VDIF entry points need segments of whole frames (ffts_per_seg a multiple of 2)");
  CK (cudaSetDevice (h->cfg.gpu_id));
  const size_t bytes = 2 * (h->nsamp / VF_VD_DAT) * (size_t) n_seg * VF_VD_FRM;
  for (int i = 0; i < 2; ++i) {
    if (h->slot[i].pending) return vf_fail (h, VF_ERR_STATE, "slot %d has a submission in flight", i);
    int rc = vf_slot_reserve_vdif (h, &h->slot[i], bytes, n_seg);
    if (rc) return rc;
  }
  CK (cudaDeviceSynchronize ());
  return VF_OK;
}

/* Frames of n_seg consecutive segments of one antenna (n_seg = 10: a one-second block of the input ring), in any
 * order: one copy, one depacketiser launch that places every frame by (thread id, frame number) across the whole
 * block -- as the reference's host loop does across the second, src/process_baseband.cu:1015-1035 -- and ONE
 * launch pair over the n_seg segments.  Frames the stream lacks stay zero = dropped samples (the writer's fill
 * frames, src/writer.c:362,674-687); frames that belong to another second (expect_second >= 0) or fall outside
 * the block are skipped and counted, never fatal: vf_wait returns VF_ERR_VDIF as a WARNING with the outputs
 * complete, vf_vdif_report gives the counts. */
static int vf_submit_vdif_common (vf_handle *h, int slot, int antenna, const void *frames, size_t nframes,
                                  uint32_t first_frame, long expect_second, int n_seg, uint8_t *fb_main, uint8_t *fb_raw)
{
  if (!h || !frames || !fb_main) return VF_ERR_ARG;
  if (slot < 0 || slot > 1) return vf_fail (h, VF_ERR_ARG, "bad slot");
  if (antenna < 0 || antenna >= h->n_ant) return vf_fail (h, VF_ERR_ARG, "antenna %d outside 0..%d", antenna, h->n_ant - 1);
  if (n_seg < 1 || n_seg > 16) return vf_fail (h, VF_ERR_ARG, "n_seg %d outside 1..16", n_seg);
  const size_t per_pol = h->nsamp / VF_VD_DAT;            /* frames per pol per segment */
  if (h->nsamp % VF_VD_DAT) return vf_fail (h, VF_ERR_ARG, "the VDIF entry points need segments of whole frames (ffts_per_seg a multiple of 2)");
  if (nframes > 4 * per_pol * n_seg) return vf_fail (h, VF_ERR_ARG, "%zu frames for %d segment(s) of %zu", nframes, n_seg, 2 * per_pol);
  vf_slot *s = &h->slot[slot];
  if (s->pending) return vf_fail (h, VF_ERR_STATE, "slot %d submitted twice without vf_wait", slot);
  CK (cudaSetDevice (h->cfg.gpu_id));
  const size_t bytes = nframes * VF_VD_FRM;
  int rc0 = vf_slot_reserve_vdif (h, s, bytes, n_seg);
  if (rc0) return rc0;
  uint8_t *d_in = s->d_in, *d_main = s->d_out_main, *d_raw = s->d_out_raw;
  if (n_seg > 1) { d_in = s->d_in_blk; d_main = s->d_out_blk[0]; d_raw = s->d_out_blk[1]; }
  CK (cudaEventRecord (s->ev_t[0], s->st));
  CK (cudaMemcpyAsync (s->d_frames, frames, bytes, cudaMemcpyHostToDevice, s->st));
  CK (cudaMemsetAsync (d_in, 0, (size_t) n_seg * 2 * h->nsamp, s->st));
  CK (cudaMemsetAsync (s->d_bad, 0, 4 * sizeof (unsigned int), s->st));
  vf_depack_params dp;
  dp.frames = s->d_frames; dp.nframes = nframes; dp.out = d_in; dp.pol_stride = h->nsamp;
  dp.seg_stride = 2 * h->nsamp; dp.frames_per_seg = (long long) per_pol;
  dp.frame0 = first_frame; dp.nframes_per_pol = (long long) per_pol * n_seg; dp.expect_second = expect_second; dp.bad = s->d_bad;
  CK (vf_launch_depack (dp, s->st));
  CK (cudaMemcpyAsync (s->h_bad, s->d_bad, 4 * sizeof (unsigned int), cudaMemcpyDeviceToHost, s->st));
  int rc = vf_enqueue_segment (h, s, antenna, 1, d_in, d_main, d_raw, -2, n_seg);
  if (rc) return rc;
  CK (cudaMemcpyAsync (fb_main, d_main, (size_t) n_seg * h->out_bytes, cudaMemcpyDeviceToHost, s->st));
  if (h->cfg.rfi_mode == 2 && fb_raw)
    CK (cudaMemcpyAsync (fb_raw, d_raw, (size_t) n_seg * h->out_bytes, cudaMemcpyDeviceToHost, s->st));
  CK (cudaEventRecord (s->ev_t[4], s->st));
  CK (cudaEventRecord (s->ev_done, s->st));
  s->t_valid = 1;
  s->pending = 2;               /* 2: vf_wait also reports frames that were skipped */
  s->first_frame = first_frame;
  s->blk_expected = 2 * per_pol * n_seg;
  return VF_OK;
}

int vf_submit_vdif_async (vf_handle *h, int slot, int antenna, const void *frames, size_t nframes,
                          uint32_t first_frame, uint8_t *fb_main, uint8_t *fb_raw)
{
  return vf_submit_vdif_common (h, slot, antenna, frames, nframes, first_frame, -1, 1, fb_main, fb_raw);
}

int vf_submit_vdif_block_async (vf_handle *h, int slot, int antenna, const void *frames, size_t nframes,
                                uint32_t first_frame, long expect_second, int n_seg, uint8_t *fb_main, uint8_t *fb_raw)
{
  return vf_submit_vdif_common (h, slot, antenna, frames, nframes, first_frame, expect_second, n_seg, fb_main, fb_raw);
}

/* what the depacketiser of the last VDIF submission on `slot` saw (valid after vf_wait):
 * counts[0] frames outside the block, [1] frames of another second, [2] frames placed, [3] frames with the invalid
 * bit, [4] frames a complete block has */
int vf_vdif_report (vf_handle *h, int slot, unsigned int counts[5])
{
  if (!h || slot < 0 || slot > 1 || !counts) return VF_ERR_ARG;
  vf_slot *s = &h->slot[slot];
  if (!s->h_bad) return vf_fail (h, VF_ERR_STATE, "no VDIF submission on slot %d yet", slot);
  if (s->pending) return vf_fail (h, VF_ERR_STATE, "slot %d not waited for", slot);
  for (int i = 0; i < 4; ++i) counts[i] = s->h_bad[i];
  counts[4] = (unsigned int) s->blk_expected;
  return VF_OK;
}

int vf_process_vdif (vf_handle *h, int antenna, const void *frames, size_t nframes,
                     uint32_t first_frame, uint8_t *fb_main, uint8_t *fb_raw, size_t *nbytes)
{
  if (!h) return VF_ERR_ARG;
  const int slot = h->next_slot;
  if (h->slot[slot].pending) return vf_fail (h, VF_ERR_STATE, "slot %d has an asynchronous segment in flight", slot);
  int rc = vf_timing_begin (h, 0);
  if (rc) return rc;
  rc = vf_submit_vdif_async (h, slot, antenna, frames, nframes, first_frame, fb_main, fb_raw);
  if (rc) return rc;
  rc = vf_timing_end (h);
  if (rc) { h->slot[slot].pending = 0; return rc; }
  h->next_slot ^= 1;
  rc = vf_wait (h, slot);
  if (rc == VF_OK && nbytes) *nbytes = h->out_bytes;
  return rc;
}

int vf_process_device (vf_handle *h, int n_ant, int n_seg, const uint8_t *d_in,
                       uint8_t *d_fb_main, uint8_t *d_fb_raw)
{
  if (!h || !d_in || !d_fb_main || n_seg < 1) return VF_ERR_ARG;
  if (n_ant < 1 || n_ant > h->n_ant) return vf_fail (h, VF_ERR_ARG, "n_ant %d outside 1..%d", n_ant, h->n_ant);
  if (((uintptr_t) d_in) & 15) return vf_fail (h, VF_ERR_ARG, "d_in must be 16-byte aligned");
  if (h->cfg.rfi_mode == 2 && !d_fb_raw) return vf_fail (h, VF_ERR_ARG, "rfi_mode 2 needs d_fb_raw");
  if (h->slot[0].pending || h->slot[1].pending) return vf_fail (h, VF_ERR_STATE, "asynchronous segment in flight");
  CK (cudaSetDevice (h->cfg.gpu_id));
  const size_t in_seg = (size_t) n_ant * 2 * h->nsamp, out_seg = (size_t) n_ant * h->out_bytes;
  /* consecutive segments share a launch pair: the channeliser takes (segment, antenna) pairs as its
   * antennas, the normaliser walks the segments in time order inside the kernel.  Fewer, longer
   * launches: less start-up and tail per segment (cfg.max_batch_segments) */
  const int nb = n_seg < vf_max_batch (h) ? n_seg : vf_max_batch (h);
  int rc = vf_ensure_batch (h, nb);
  if (rc) return rc;
  rc = vf_timing_begin (h, (n_seg + nb - 1) / nb);       /* one (K1, K2) pair of events per launch pair */
  if (rc) return rc;
  for (int sg = 0, bi = 0; sg < n_seg; sg += nb, ++bi) {
    const int m = n_seg - sg < nb ? n_seg - sg : nb;
    vf_slot *s = &h->slot[h->next_slot];
    rc = vf_enqueue_segment (h, s, 0, n_ant, d_in + (size_t) sg * in_seg, d_fb_main + (size_t) sg * out_seg,
                             d_fb_raw ? d_fb_raw + (size_t) sg * out_seg : NULL, bi < h->n_timed ? bi : -1, m);
    if (rc) return rc;
    h->next_slot ^= 1;
  }
  return vf_timing_end (h);
}

#ifdef VF_TESTING
/* TESTING BUILDS ONLY: clock64 stamps of the normaliser's chunk pipeline (first CTA): the next launches write
 * [2 streams][4096 chunks][6 stamps] into a device buffer; copy it out with out != NULL */
int vf_debug_k2_trace (vf_handle *h, long long *out)
{
  if (!h) return VF_ERR_ARG;
  CK (cudaSetDevice (h->cfg.gpu_id));
  const size_t n = 2 * 4096 * 6 * sizeof (long long);
  if (!h->k2_trace) { CK (cudaMalloc ((void **) &h->k2_trace, n)); CK (cudaMemset (h->k2_trace, 0, n)); }
  if (out) { int rc = vf_sync (h); if (rc) return rc; CK (cudaMemcpy (out, h->k2_trace, n, cudaMemcpyDeviceToHost)); }
  return VF_OK;
}

/* TESTING BUILDS ONLY (libvlitefast_testing.so, csrc/vf_testing.h): self-check of the packed division used by the
 * normaliser: q_packed from the kernel's own routine, q_ref from CUDA's div.rn.f32, for n (even) host operand pairs */
int vf_debug_division (vf_handle *h, const float *p, const float *b, float *q_packed, float *q_ref, size_t n)
{
  if (!h || !p || !b || !q_packed || !q_ref || (n & 1)) return VF_ERR_ARG;
  CK (cudaSetDevice (h->cfg.gpu_id));
  float *d = NULL;
  CK (cudaMalloc ((void **) &d, 4 * n * sizeof (float)));
  CK (cudaMemcpy (d, p, n * sizeof (float), cudaMemcpyHostToDevice));
  CK (cudaMemcpy (d + n, b, n * sizeof (float), cudaMemcpyHostToDevice));
  CK (vf_launch_debug_div (d, d + n, d + 2 * n, d + 3 * n, n, h->ctl));
  CK (cudaStreamSynchronize (h->ctl));
  CK (cudaMemcpy (q_packed, d + 2 * n, n * sizeof (float), cudaMemcpyDeviceToHost));
  CK (cudaMemcpy (q_ref, d + 3 * n, n * sizeof (float), cudaMemcpyDeviceToHost));
  cudaFree (d);
  return VF_OK;
}
#endif

/* serial != 0: segments do not overlap (K1 of segment n+1 waits for K2 of segment n), so that the
 * per-kernel times of vf_last_elapsed_ms are pure execution times.  Also set by VF_SERIAL=1. */
int vf_set_serial (vf_handle *h, int serial)
{
  if (!h) return VF_ERR_ARG;
  h->serial = serial != 0;
  return VF_OK;
}

int vf_sync (vf_handle *h)
{
  if (!h) return VF_ERR_ARG;
  CK (cudaStreamSynchronize (h->ctl));
  CK (cudaStreamSynchronize (h->slot[0].st));
  CK (cudaStreamSynchronize (h->slot[1].st));
  CK (cudaStreamSynchronize (h->coadd_hi));
  CK (cudaStreamSynchronize (h->coadd_st));
  return VF_OK;
}

/* Device-side stopwatch over everything the handle enqueues between the two calls, on all of its streams
 * (slots and co-add): begin = an event on the control stream that every stream waits for; end = an event on the
 * control stream after it has joined every stream.  vf_timer_end blocks until that event and returns the time. */
int vf_timer_begin (vf_handle *h)
{
  if (!h) return VF_ERR_ARG;
  CK (cudaSetDevice (h->cfg.gpu_id));
  CK (cudaEventRecord (h->ev_u0, h->ctl));
  CK (cudaStreamWaitEvent (h->slot[0].st, h->ev_u0, 0));
  CK (cudaStreamWaitEvent (h->slot[1].st, h->ev_u0, 0));
  CK (cudaStreamWaitEvent (h->coadd_st, h->ev_u0, 0));
  CK (cudaStreamWaitEvent (h->coadd_hi, h->ev_u0, 0));
  return VF_OK;
}

int vf_timer_end (vf_handle *h, float *ms)
{
  if (!h || !ms) return VF_ERR_ARG;
  CK (cudaSetDevice (h->cfg.gpu_id));
  cudaStream_t sts[4] = { h->slot[0].st, h->slot[1].st, h->coadd_st, h->coadd_hi };
  for (int i = 0; i < 4; ++i) {
    CK (cudaEventRecord (h->ev_u1, sts[i]));          /* re-recorded per stream: the wait below captures this record */
    CK (cudaStreamWaitEvent (h->ctl, h->ev_u1, 0));
  }
  CK (cudaEventRecord (h->ev_u1, h->ctl));
  CK (cudaEventSynchronize (h->ev_u1));
  CK (cudaEventElapsedTime (ms, h->ev_u0, h->ev_u1));
  return VF_OK;
}

int vf_last_elapsed_ms (vf_handle *h, float *total_ms, float *k1_ms, float *k2_ms)
{
  if (!h) return VF_ERR_ARG;
  if (!h->timed_valid) return vf_fail (h, VF_ERR_STATE, "nothing timed yet");
  CK (cudaEventSynchronize (h->ev_t1));
  if (total_ms) CK (cudaEventElapsedTime (total_ms, h->ev_t0, h->ev_t1));
  float a = 0.f, b = 0.f;
  for (int i = 0; i < h->n_timed; ++i) {
    float x = 0.f, y = 0.f;
    CK (cudaEventElapsedTime (&x, h->ev_ka[i], h->ev_kb[i]));
    CK (cudaEventElapsedTime (&y, h->ev_kb[i], h->ev_kc[i]));
    a += x; b += y;
  }
  /* sums over the segments of the last vf_process_device call; k2 includes
   * any wait for the previous segment's K2 (bandpass order) */
  if (k1_ms) *k1_ms = a;
  if (k2_ms) *k2_ms = b;
  return VF_OK;
}

/* device times of the last asynchronous submission on `slot` (after vf_wait): total from the start of its copy in to
 * the end of its copy out, and the two kernels (K2 includes any wait for the bandpass order of the other slot) */
int vf_slot_elapsed_ms (vf_handle *h, int slot, float *total_ms, float *k1_ms, float *k2_ms)
{
  if (!h || slot < 0 || slot > 1) return VF_ERR_ARG;
  vf_slot *s = &h->slot[slot];
  if (!s->t_valid) return vf_fail (h, VF_ERR_STATE, "nothing submitted on slot %d yet", slot);
  if (s->pending) return vf_fail (h, VF_ERR_STATE, "slot %d not waited for", slot);
  CK (cudaEventSynchronize (s->ev_t[4]));
  if (total_ms) CK (cudaEventElapsedTime (total_ms, s->ev_t[0], s->ev_t[4]));
  if (k1_ms) CK (cudaEventElapsedTime (k1_ms, s->ev_t[1], s->ev_t[2]));
  if (k2_ms) CK (cudaEventElapsedTime (k2_ms, s->ev_t[2], s->ev_t[3]));
  return VF_OK;
}

/* ---- inspection ----------------------------------------------------------- */
static int vf_check_ant (vf_handle *h, int antenna)
{
  if (!h) return VF_ERR_ARG;
  if (antenna < 0 || antenna >= h->n_ant) return vf_fail (h, VF_ERR_ARG, "antenna %d outside 0..%d", antenna, h->n_ant - 1);
  return VF_OK;
}

int vf_get_stats (vf_handle *h, int antenna, float *pw, float *kur, float *dag,
                  float *pw_fb, float *kur_fb, float *dag_fb, float *weights, uint32_t *histo)
{
  int rc = vf_check_ant (h, antenna);
  if (rc) return rc;
  rc = vf_sync (h);
  if (rc) return rc;
  const size_t nblk = (size_t) h->T * VF_NSUB, T = h->T;
  const vf_slot *s = &h->slot[h->ant_last[antenna].slot];      /* the slot that processed this antenna last */
  if (pw || kur || dag || pw_fb || kur_fb || dag_fb) {
    if (!s->pw) return vf_fail (h, VF_ERR_STATE, "statistics need keep_stats and rfi_mode != 0");
    if (pw) CK (cudaMemcpy (pw, s->pw + (size_t) antenna * 2 * nblk, 2 * nblk * 4, cudaMemcpyDeviceToHost));
    if (kur) CK (cudaMemcpy (kur, s->kur + (size_t) antenna * 2 * nblk, 2 * nblk * 4, cudaMemcpyDeviceToHost));
    if (dag) CK (cudaMemcpy (dag, s->dag + (size_t) antenna * 2 * nblk, 2 * nblk * 4, cudaMemcpyDeviceToHost));
    if (pw_fb) CK (cudaMemcpy (pw_fb, s->pw_fb + (size_t) antenna * 2 * T, 2 * T * 4, cudaMemcpyDeviceToHost));
    if (kur_fb) CK (cudaMemcpy (kur_fb, s->kur_fb + (size_t) antenna * 2 * T, 2 * T * 4, cudaMemcpyDeviceToHost));
    if (dag_fb) CK (cudaMemcpy (dag_fb, s->dag_fb + (size_t) antenna * 2 * T, 2 * T * 4, cudaMemcpyDeviceToHost));
  }
  if (weights) {
    if (!h->cfg.rfi_mode) return vf_fail (h, VF_ERR_STATE, "weights need rfi_mode != 0");
    /* both pols always carry the same weight (src/pb_kernels.cu:132) */
    CK (cudaMemcpy (weights, s->w + h->ant_last[antenna].idx * T, T * 4, cudaMemcpyDeviceToHost));
    memcpy (weights + T, weights, T * 4);
  }
  if (histo) {
    if (!s->histo) return vf_fail (h, VF_ERR_STATE, "histogram needs do_histo");
    CK (cudaMemcpy (histo, s->histo + (size_t) antenna * 512, 512 * 4, cudaMemcpyDeviceToHost));
  }
  return VF_OK;
}

int vf_get_mask (vf_handle *h, int antenna, uint32_t *mask)
{
  int rc = vf_check_ant (h, antenna);
  if (rc) return rc;
  if (!mask) return VF_ERR_ARG;
  rc = vf_sync (h);
  if (rc) return rc;
  CK (cudaMemcpy (mask, h->slot[h->ant_last[antenna].slot].mask + h->ant_last[antenna].idx * h->T, (size_t) h->T * 4, cudaMemcpyDeviceToHost));
  return VF_OK;
}

int vf_get_power_f32 (vf_handle *h, int antenna, int which, float *out)
{
  int rc = vf_check_ant (h, antenna);
  if (rc) return rc;
  if (!out || which < 0 || which > 1) return VF_ERR_ARG;
  const float *src = which ? h->ave_raw : h->ave_main;
  if (!src) return vf_fail (h, VF_ERR_STATE, "needs keep_power (and rfi_mode 2 for which = 1)");
  rc = vf_sync (h);
  if (rc) return rc;
  const size_t n = (size_t) h->cfg.npol * h->ntime * VF_NCHANOUT;
  const size_t last = (size_t) ((h->ant_seg[antenna] + h->ave_nseg - 1) % h->ave_nseg) * h->n_ant * n;
  CK (cudaMemcpy (out, src + last + (size_t) antenna * n, n * 4, cudaMemcpyDeviceToHost));
  return VF_OK;
}

int vf_get_detected_power (vf_handle *h, int antenna, int which, float *out)
{
  int rc = vf_check_ant (h, antenna);
  if (rc) return rc;
  if (!out || which < 0 || which > 1) return VF_ERR_ARG;
  const int mode = h->cfg.rfi_mode;
  if (which == 1 && mode != 2) return vf_fail (h, VF_ERR_STATE, "raw stream beside the main one needs rfi_mode 2");
  rc = vf_sync (h);
  if (rc) return rc;
  vf_slot *s = &h->slot[h->ant_last[antenna].slot];
  const size_t n = h->tile_elems, off = h->ant_last[antenna].idx * n;
  const bool want_kur = (which == 0 && mode != 0);
  /* the device tile is [4096 / VF_PBLK][T][VF_PBLK] (VF_PBLK = 4096: plain [T][4096]): copy it out in [T][4096] order */
  std::vector<float2> blk (n), blk_raw;
  CK (cudaMemcpy (blk.data (), (want_kur ? s->P_kur : s->P_raw) + off, n * sizeof (float2), cudaMemcpyDeviceToHost));
  std::vector<uint32_t> mk;
  if (want_kur && mode == 2) {
    /* time steps with an empty mask were not re-transformed: identical to raw */
    mk.resize (h->T);
    blk_raw.resize (n);
    CK (cudaMemcpy (mk.data (), s->mask + h->ant_last[antenna].idx * h->T, (size_t) h->T * 4, cudaMemcpyDeviceToHost));
    CK (cudaMemcpy (blk_raw.data (), s->P_raw + off, n * sizeof (float2), cudaMemcpyDeviceToHost));
  }
  float2 *o2 = reinterpret_cast<float2 *> (out);
  for (int t = 0; t < h->T; ++t) {
    const float2 *src = (!mk.empty () && mk[t] == 0) ? blk_raw.data () : blk.data ();
    for (int c = 0; c < VF_NCHANOUT; ++c)
      o2[(size_t) t * VF_NCHANOUT + c] = src[(size_t) t * VF_PBLK + VF_PIDX (h->T, c)];
  }
  return VF_OK;
}

static float2 *vf_bp_of (vf_handle *h, int which)
{
  if (h->cfg.rfi_mode == 2) return which ? h->bp_raw : h->bp_kur;
  return which ? NULL : h->bp_raw;
}

int vf_get_bandpass (vf_handle *h, int antenna, int which, float *out)
{
  int rc = vf_check_ant (h, antenna);
  if (rc) return rc;
  float2 *bp = (out && which >= 0 && which <= 1) ? vf_bp_of (h, which) : NULL;
  if (!bp) return vf_fail (h, VF_ERR_ARG, "no such bandpass");
  rc = vf_sync (h);
  if (rc) return rc;
  std::vector<float2> tmp (VF_NCHANOUT);
  CK (cudaMemcpy (tmp.data (), bp + (size_t) antenna * VF_NCHANOUT, VF_NCHANOUT * sizeof (float2), cudaMemcpyDeviceToHost));
  for (int c = 0; c < VF_NCHANOUT; ++c) { out[c] = tmp[c].x; out[VF_NCHANOUT + c] = tmp[c].y; }
  return VF_OK;
}

int vf_set_bandpass (vf_handle *h, int antenna, int which, const float *in)
{
  int rc = vf_check_ant (h, antenna);
  if (rc) return rc;
  float2 *bp = (in && which >= 0 && which <= 1) ? vf_bp_of (h, which) : NULL;
  if (!bp) return vf_fail (h, VF_ERR_ARG, "no such bandpass");
  rc = vf_sync (h);
  if (rc) return rc;
  std::vector<float2> tmp (VF_NCHANOUT);
  for (int c = 0; c < VF_NCHANOUT; ++c) tmp[c] = make_float2 (in[c], in[VF_NCHANOUT + c]);
  CK (cudaMemcpy (bp + (size_t) antenna * VF_NCHANOUT, tmp.data (), VF_NCHANOUT * sizeof (float2), cudaMemcpyHostToDevice));
  return VF_OK;
}

int vf_reset_bandpass (vf_handle *h, int antenna)
{
  if (!h) return VF_ERR_ARG;
  if (antenna >= h->n_ant) return vf_fail (h, VF_ERR_ARG, "antenna %d outside 0..%d", antenna, h->n_ant - 1);
  int rc = vf_sync (h);
  if (rc) return rc;
  const size_t off = antenna < 0 ? 0 : (size_t) antenna * VF_NCHANOUT;
  const size_t n = (antenna < 0 ? (size_t) h->n_ant : 1) * VF_NCHANOUT * sizeof (float2);
  CK (cudaMemset (h->bp_raw + off, 0, n));
  if (h->bp_kur) CK (cudaMemset (h->bp_kur + off, 0, n));
  return VF_OK;
}

int vf_set_frb_injection (vf_handle *h, int nfft_since_frb, float dm, float width, float amp)
{
  if (!h) return VF_ERR_ARG;
  if (!h->cfg.inject_frb) return vf_fail (h, VF_ERR_STATE, "handle created without inject_frb");
  if (nfft_since_frb < 0) { h->frb_nfft_since = -1; return VF_OK; }
  if (!h->frb_delays || dm != h->frb_dm) {
    /* set_frb_delays, src/pb_kernels.cu:338-346, in the same double arithmetic */
    std::vector<float> d (VF_NCHAN_FFT);
    for (int i = 0; i < VF_NCHAN_FFT; ++i) {
      double freq = 0.384 - (i * 0.064) / VF_NCHAN_FFT;
      double scale = 4.15e-3 * dm * 10 * 128000000 / 10 / VF_NFFT;
      d[i] = (float) (scale / (freq * freq) - scale / (0.384 * 0.384));
    }
    int rc = vf_sync (h);
    if (rc) return rc;
    if (!h->frb_delays) CK (cudaMalloc ((void **) &h->frb_delays, VF_NCHAN_FFT * sizeof (float)));
    CK (cudaMemcpy (h->frb_delays, d.data (), VF_NCHAN_FFT * sizeof (float), cudaMemcpyHostToDevice));
    h->frb_dm = dm;
  }
  h->frb_nfft_since = nfft_since_frb; h->frb_width = width; h->frb_amp = amp;
  return VF_OK;
}

/* ---- co-add --------------------------------------------------------------- */
static int vf_nccl_load (vf_handle *h)
{
  if (g_nccl.lib) return VF_OK;
  const char *names[] = { "libnccl.so.2", "libnccl.so", NULL };
  void *lib = NULL;
  for (int i = 0; names[i] && !lib; ++i) lib = dlopen (names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return vf_fail (h, VF_ERR_NCCL, "cannot load libnccl: %s", dlerror ());
  g_nccl.GetUniqueId = (int (*) (vf_nccl_id *)) dlsym (lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*) (vf_nccl_comm *, int, vf_nccl_id, int)) dlsym (lib, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*) (const void *, void *, size_t, int, int, vf_nccl_comm, cudaStream_t)) dlsym (lib, "ncclAllReduce");
  g_nccl.Reduce = (int (*) (const void *, void *, size_t, int, int, int, vf_nccl_comm, cudaStream_t)) dlsym (lib, "ncclReduce");
  g_nccl.CommDestroy = (int (*) (vf_nccl_comm)) dlsym (lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char *(*) (int)) dlsym (lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.Reduce || !g_nccl.CommDestroy)
    return vf_fail (h, VF_ERR_NCCL, "libnccl lacks a required symbol");
  g_nccl.lib = lib;
  return VF_OK;
}

int vf_coadd_unique_id (void *out128)
{
  if (!out128) return VF_ERR_ARG;
  int rc = vf_nccl_load (NULL);
  if (rc) return rc;
  vf_nccl_id id;
  if (g_nccl.GetUniqueId (&id) != 0) return VF_ERR_NCCL;
  memcpy (out128, &id, 128);
  return VF_OK;
}

int vf_coadd_init (vf_handle *h, int nranks, int rank, const void *nccl_unique_id)
{
  if (!h || nranks < 1 || rank < 0 || rank >= nranks) return VF_ERR_ARG;
  if (!h->ave_main) return vf_fail (h, VF_ERR_STATE, "co-add needs keep_power");
  CK (cudaSetDevice (h->cfg.gpu_id));
  const size_t n = (size_t) h->cfg.npol * h->ntime * VF_NCHANOUT;
  if (!h->coadd_sum) {
    CK (cudaMalloc ((void **) &h->coadd_sum, (n + h->ntime) * h->ave_nseg * sizeof (float)));
    CK (cudaMalloc ((void **) &h->coadd_out, n * h->ave_nseg * h->cfg.nbit / 8));
    CK (cudaEventCreateWithFlags (&h->coadd_batch[0].ev, cudaEventDisableTiming));
    CK (cudaEventCreateWithFlags (&h->coadd_batch[1].ev, cudaEventDisableTiming));
  }
  h->nranks = nranks; h->rank = rank;
  if (nranks > 1) {
    if (!nccl_unique_id) return vf_fail (h, VF_ERR_ARG, "nranks > 1 needs the NCCL unique id");
    /* The reduce of a second (21 MB per rank) has a whole second of slack, but its CTAs need SMs of their own: both
     * kernels of the chain fill their SMs' registers, and the normaliser is a single wave of 128 CTAs that leaves 20 SMs
     * free.  NCCL's default channel count takes more than those (measured at 4 GPUs, round 2: 4110 antenna-seconds/s
     * with the default, 4230 with 8 channels, 4298 = 4 x the 1-GPU rate), so a late normaliser CTA stretches its launch.
     * NCCL reads the variable when the process creates its first communicator: a caller that already has one
     * (torch.distributed) sets it itself before that (bench.py does). */
    setenv ("NCCL_MAX_NCHANNELS", "8", 0);
    int rc = vf_nccl_load (h);
    if (rc) return rc;
    vf_nccl_id id;
    memcpy (&id, nccl_unique_id, 128);
    int e = g_nccl.CommInitRank (&h->comm, nranks, id, rank);
    if (e != 0) return vf_fail (h, VF_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString (e) : "?");
  }
  return VF_OK;
}

/* Co-add of the last n_seg segments.  Definition (the reference's co-adder is an external
 * program whose source is not in the tree): per output sample, the sum over antennas of the main-stream
 * pre-digitisation values divided by sqrt (number of antennas whose scrunched row was kept), the way
 * tscrunch_weights (src/pb_kernels.cu:591-630) treats the time steps of a row; 0 when no antenna kept the row.
 * Values and counts travel in ONE reduce: [n_seg tiles | n_seg x T/8 counts]. */
int vf_coadd_batch (vf_handle *h, int root, int total_antennas, int n_seg, uint8_t *fb_coadd, float *sum_f32, int wait)
{
  if (!h || total_antennas < 1 || n_seg < 1) return VF_ERR_ARG;
  if (!h->coadd_sum) return vf_fail (h, VF_ERR_STATE, "vf_coadd_init not called");
  const int na = h->last_n_ant;                 /* antennas of the last launches: their tiles are the current ones */
  const long seg_end = h->ant_seg[0];
  for (int a = 1; a < na; ++a)
    if (h->ant_seg[a] != seg_end) return vf_fail (h, VF_ERR_STATE, "antennas 0 and %d are not at the same segment", a);
  if (n_seg > h->ave_nseg || n_seg > seg_end)
    return vf_fail (h, VF_ERR_ARG, "n_seg %d exceeds the %d kept tile(s)", n_seg, h->ave_nseg);
  if (root < 0 || root >= (h->nranks ? h->nranks : 1)) return vf_fail (h, VF_ERR_ARG, "bad root");
  CK (cudaSetDevice (h->cfg.gpu_id));
  const size_t n = (size_t) h->cfg.npol * h->ntime * VF_NCHANOUT;        /* one tile */
  const size_t out1 = n * h->cfg.nbit / 8;
  float *cnt = h->coadd_sum + (size_t) n_seg * n;                        /* [n_seg][T/8] right behind the sums */
  /* Multi-rank: the co-add goes to the highest-priority stream -- the reduce couples the ranks, and behind the queued
   * CTAs of the next persistent channeliser it would start a whole launch later (2 GPUs, round 2: 2171-2177 ->
   * 2309-2333 antenna-seconds/s; with only the local sum and the reduce there and the root's digitiser at normal
   * priority: 2298).  Single rank: nothing to couple, and at highest priority the co-add's kernels sit in front of the
   * channeliser instead of in the shadow of the next normaliser (1224 -> 1182): normal priority. */
  cudaStream_t st = h->nranks > 1 ? h->coadd_hi : h->coadd_st;
  /* after every K2 that wrote the tiles */
  if (h->have_k2_last) CK (cudaStreamWaitEvent (st, h->ev_k2_last, 0));
  /* local sum and count over this handle's antennas, the segments of the batch in time order, one launch */
  {
    vf_coadd_local_params lp;
    lp.tiles = h->ave_main; lp.rowok = h->rowok;
    lp.seg0 = seg_end - n_seg; lp.nring = h->ave_nseg; lp.n_ant_total = h->n_ant;
    lp.n_ant = na; lp.ntime = h->ntime; lp.npol = h->cfg.npol; lp.n_seg = n_seg;
    lp.sum = h->coadd_sum; lp.cnt = cnt;
    CK (vf_launch_coadd_local (lp, st));
  }
  if (h->nranks > 1) {
    /* ncclFloat32 = 7, ncclSum = 0 (nccl.h); one collective for the whole batch, sums and counts */
    int e = g_nccl.Reduce (h->coadd_sum, h->coadd_sum, (n + h->ntime) * n_seg, 7, 0, root, h->comm, st);
    if (e != 0) return vf_fail (h, VF_ERR_NCCL, "ncclReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString (e) : "?");
  }
  if (h->rank == root) {
    vf_coadd_params cp;
    cp.sum = h->coadd_sum; cp.cnt = cnt;
    cp.ntime = h->ntime; cp.npol = h->cfg.npol; cp.nbit = h->cfg.nbit; cp.n_seg = n_seg; cp.out = h->coadd_out;
    CK (vf_launch_coadd (cp, st));
    if (fb_coadd) CK (cudaMemcpyAsync (fb_coadd, h->coadd_out, out1 * n_seg, cudaMemcpyDeviceToHost, st));
    if (sum_f32) CK (cudaMemcpyAsync (sum_f32, h->coadd_sum, n * n_seg * sizeof (float), cudaMemcpyDeviceToHost, st));
  }
  /* later segments overwrite the tile ring: those that hit this batch's tiles wait for it */
  {
    const int b = h->coadd_next;
    if (h->coadd_batch[b].pending) CK (cudaStreamWaitEvent (st, h->coadd_batch[b].ev, 0));   /* same stream: already ordered */
    CK (cudaEventRecord (h->coadd_batch[b].ev, st));
    h->coadd_batch[b].lo = seg_end - n_seg; h->coadd_batch[b].hi = seg_end; h->coadd_batch[b].pending = 1;
    h->coadd_next ^= 1;
  }
  if (wait) CK (cudaStreamSynchronize (st));
  (void) total_antennas;      /* kept in the signature: the count of contributing antennas is reduced with the data */
  return VF_OK;
}

int vf_coadd_segment (vf_handle *h, int root, int total_antennas, uint8_t *fb_coadd, float *sum_f32)
{
  return vf_coadd_batch (h, root, total_antennas, 1, fb_coadd, sum_f32, 1);
}

} /* extern "C" */
