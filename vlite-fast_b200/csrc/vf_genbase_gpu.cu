/*
 * libvlitegen: GPU baseband generator (include/vlitegen.h), the counterpart of
 * the reference's src/genbase.cu.  One convolution block (src/genbase.cu:366-438):
 *
 *   vfg_k_noise      noise N(0,1) of the block's buflen input samples with the
 *                    pulse profile on top (curandGenerateNormal + set_profile,
 *                    :375-384, :554-585), one kernel, no overlap buffer: a
 *                    sample is a pure function of (seed, pol, sample index)
 *   cufftExecR2C     (:391)
 *   vfg_k_chirp      multiply by the dispersion kernel (:394, :587-598; table
 *                    from vfg_k_init_kernel = init_dm_kernel, :525-552)
 *   cufftExecC2R     (:398)
 *   vfg_k_epilogue   side-band swap (:401, :651-661), RFI (:421-428, :671-687),
 *                    digitise the valid samples (:431-433, :690-708), one pass
 *   vfg_k_frames     VDIF framing of a second (:443-486) on the device
 */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime.h>
#include <cufft.h>
#include "vlitegen.h"

#define VFG_RATE      128000000ll   /* VLITE_RATE, src/def.h */
#define VFG_VD_DAT    5000
#define VFG_VD_FRM    5032
#define VFG_FRAMES    25600

struct vfg_handle {
  vfg_config cfg;
  long long buflen, n_lo, n_hi, n_dm, new_samps, period;
  float *d_f;                 /* [buflen + 2] block, real view of d_c            */
  cufftComplex *d_c;          /* [buflen/2 + 1] in-place transform               */
  cufftComplex *d_ker;        /* [buflen/2 + 1]                                  */
  uint8_t *d_out[2];          /* [new_samps] digitised valid samples of the block */
  float *d_keep[2];           /* [new_samps] pre-digitisation voltages (tests)   */
  uint8_t *d_sec[2];          /* [VFG_RATE] one second per pol (VDIF path)       */
  uint8_t *d_vdif;            /* one second of frames                            */
  cufftHandle fwd, bwd;
  int have_plans;
  long long block;            /* blocks generated so far                         */
  long long avail, rd;        /* unread samples of the current block, read offset */
  cudaStream_t st;
  char err[256];
};

static int vfg_fail (vfg_handle *h, const char *fmt, ...)
{
  if (h) {
    va_list ap;
    va_start (ap, fmt);
    vsnprintf (h->err, sizeof (h->err), fmt, ap);
    va_end (ap);
  }
  return 1;
}
#define CKG(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return vfg_fail (h, "%s: %s (%s:%d)", #x, cudaGetErrorString (e_), __FILE__, __LINE__); } while (0)
#define CKF(x) do { cufftResult r_ = (x); if (r_ != CUFFT_SUCCESS) return vfg_fail (h, "%s: cufft error %d (%s:%d)", #x, (int) r_, __FILE__, __LINE__); } while (0)

/* ---- Philox-4x32-10 (Salmon et al. 2011), counter-based ------------------- */
__host__ __device__ inline void vfg_philox (uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           uint32_t k0, uint32_t k1, uint32_t out[4])
{
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t) 0xD2511F53u * c0, p1 = (uint64_t) 0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t) p1;
    const uint32_t n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t) p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* uniform in (0, 1) with 24 bits, exactly representable */
__host__ __device__ inline float vfg_u01 (uint32_t x) { return ((float) (x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

/* noise + pulse profile.  Thread = 4 consecutive samples (one Philox call: two Box-Muller pairs).
 * base = absolute index of the block's first input sample (a multiple of 4 is not required: the
 * group of 4 is addressed by absolute index, so blocks tile the same stream). */
__global__ void vfg_k_noise (float *f, long long buflen, long long base, int pol, unsigned long long seed,
                             long long period, int skip_period, float ampl)
{
  const long long g0 = base >> 2, g1 = (base + buflen - 1) >> 2;
  for (long long g = g0 + (long long) blockIdx.x * blockDim.x + threadIdx.x; g <= g1; g += (long long) gridDim.x * blockDim.x) {
    uint32_t r[4];
    vfg_philox ((uint32_t) g, (uint32_t) ((unsigned long long) g >> 32), (uint32_t) pol, 0u,
                (uint32_t) seed, (uint32_t) (seed >> 32), r);
    float z[4];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float rad = sqrtf (-2.0f * logf (vfg_u01 (r[2 * k]))), th = 6.283185307179586f * vfg_u01 (r[2 * k + 1]);
      z[2 * k] = rad * cosf (th); z[2 * k + 1] = rad * sinf (th);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long sample = 4 * g + k, i = sample - base;
      if (i < 0 || i >= buflen) continue;
      /* set_profile, src/genbase.cu:554-585 */
      const long long phasei = sample / period;
      const float phasef = (float) (sample - phasei * period) / (float) period;
      float v = z[k];
      if (phasef < 0.03f && (phasei % skip_period) == 0) v *= ampl;
      f[i] = v;
    }
  }
}

/* init_dm_kernel, src/genbase.cu:525-552: chirp, FFT normalisation and band-pass taper; n = buflen/2 + 1 */
__global__ void vfg_k_init_kernel (cufftComplex *ker, double dm, long long n)
{
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
    double freq = (64. * (double) i) / (double) n;
    const double freq0 = 320.;
    const double arg = (2 * M_PI * dm / 2.41e-10) * freq * freq / (freq0 * freq0 * (freq0 + freq));
    double rs, rc;
    sincos (arg, &rs, &rc);
    double kx = rc / (2 * (n - 1)), ky = rs / (2 * (n - 1));
    freq *= 1. / 64;
    double scale = 1 - exp (-(freq * freq) / (0.05 * 0.05));
    scale -= exp (-((1 - freq) * (1 - freq)) / (0.10 * 0.10));
    scale *= (1 + 0.20 * freq);
    ker[i].x = (float) (kx * scale);
    ker[i].y = (float) (ky * scale);
  }
}

/* multiply_kernel, src/genbase.cu:587-598 */
__global__ void vfg_k_chirp (cufftComplex *dat, const cufftComplex *ker, long long n)
{
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
    const cufftComplex d = dat[i], k = ker[i];
    dat[i].x = __fsub_rn (__fmul_rn (d.x, k.x), __fmul_rn (d.y, k.y));
    dat[i].y = __fadd_rn (__fmul_rn (d.x, k.y), __fmul_rn (d.y, k.x));
  }
}

/* valid samples i in [n_lo, n_lo + new_samps) of the block: swap_sideband (buffer index parity),
 * add_rfi (phase of the output sample index out0 + j; a uniform deviate of +-2.5 during 10 % of every
 * 11.3 us), digitize.  out0 = index of the block's first output sample in the stream. */
__global__ void vfg_k_epilogue (const float *f, long long n_lo, long long new_samps, long long out0, int pol,
                                unsigned long long seed, int add_rfi, uint8_t *out, float *keep)
{
  const double tsamp_us = 1e6 / (double) VFG_RATE;
  for (long long j = (long long) blockIdx.x * blockDim.x + threadIdx.x; j < new_samps; j += (long long) gridDim.x * blockDim.x) {
    const long long i = n_lo + j;
    float v = f[i];
    if (i & 1) v = -v;
    if (add_rfi) {
      const long long o = out0 + j;
      const float phase = fmodf ((float) ((double) o * (tsamp_us / 11.3)), 1.0f);
      if (phase < 0.1f) {
        uint32_t r[4];
        vfg_philox ((uint32_t) (o >> 2), (uint32_t) ((unsigned long long) (o >> 2) >> 32), (uint32_t) pol, 1u,
                    (uint32_t) seed, (uint32_t) (seed >> 32), r);
        v += 5.0f * (vfg_u01 (r[o & 3]) - 0.5f);
      }
    }
    if (keep) keep[j] = v;
    const float tmp = v / 0.02957f / 2 + 128.5f;                  /* :700-706 */
    out[j] = tmp <= 0 ? 0 : (tmp >= 255 ? 255 : (uint8_t) tmp);
  }
}

/* one CTA per (frame, thread): 32-byte header + 5000 samples (src/genbase.cu:443-486; header bit layout
 * analysis/baseband.py:19-28, as host/vf_genbase.c writes it) */
__global__ void __launch_bounds__ (256) vfg_k_frames (const uint8_t *p0, const uint8_t *p1, uint8_t *out,
                                                       uint32_t second, int station)
{
  const int f = blockIdx.x >> 1, th = blockIdx.x & 1;
  uint8_t *dst = out + (size_t) blockIdx.x * VFG_VD_FRM;
  if (threadIdx.x < 8) {
    uint32_t w = 0;
    if (threadIdx.x == 0) w = second & 0x3FFFFFFFu;
    if (threadIdx.x == 1) w = ((uint32_t) f & 0xFFFFFFu) | (30u << 24);
    if (threadIdx.x == 2) w = (VFG_VD_FRM / 8) & 0xFFFFFFu;
    if (threadIdx.x == 3) w = ((uint32_t) (station & 0xFFFF)) | ((uint32_t) th << 16) | (7u << 26);
    reinterpret_cast<uint32_t *> (dst)[threadIdx.x] = w;
  }
  const uint2 *src = reinterpret_cast<const uint2 *> ((th ? p1 : p0) + (size_t) f * VFG_VD_DAT);   /* 5000 % 8 == 0 */
  uint2 *d = reinterpret_cast<uint2 *> (dst + 32);
  for (int i = threadIdx.x; i < VFG_VD_DAT / 8; i += blockDim.x) d[i] = src[i];
}

extern "C" {

int vfg_config_default (vfg_config *c)
{
  if (!c) return 1;
  memset (c, 0, sizeof (*c));
  c->dm = 30; c->pulse_period = 0.5; c->ampl[0] = c->ampl[1] = 0.05f;
  c->skip_period = 1; c->seed = 42; c->buflen = VFG_RATE / 4;
  return 0;
}

const char *vfg_last_error (const vfg_handle *h) { return h ? h->err : "no handle"; }
long long vfg_block_samples (const vfg_handle *h) { return h ? h->new_samps : 0; }
long long vfg_sweep_samples (const vfg_handle *h) { return h ? h->n_dm : 0; }

int vfg_destroy (vfg_handle *h)
{
  if (!h) return 0;
  cudaSetDevice (h->cfg.gpu_id);
  cudaDeviceSynchronize ();
  if (h->have_plans) { cufftDestroy (h->fwd); cufftDestroy (h->bwd); }
  cudaFree (h->d_c); cudaFree (h->d_ker);
  for (int p = 0; p < 2; ++p) { cudaFree (h->d_out[p]); cudaFree (h->d_keep[p]); cudaFree (h->d_sec[p]); }
  cudaFree (h->d_vdif);
  if (h->st) cudaStreamDestroy (h->st);
  free (h);
  return 0;
}

int vfg_create (const vfg_config *cfg, vfg_handle **out)
{
  if (!cfg || !out) return 1;
  *out = NULL;
  int ndev = 0;
  if (cudaGetDeviceCount (&ndev) != cudaSuccess || ndev <= 0 || cfg->gpu_id < 0 || cfg->gpu_id >= ndev) { cudaGetLastError (); return 25; }
  vfg_handle *h = (vfg_handle *) calloc (1, sizeof (*h));
  if (!h) return 21;
  h->cfg = *cfg;
  *out = h;
  /* sample counts of the sweep, src/genbase.cu:173-197 (the reference's integer truncations kept) */
  const double freq = 352, freq_hi = 384, freq_lo = 320, tsamp = 1.0 / VFG_RATE;
  const double t_dm_lo = cfg->dm / 2.41e-10 * (1. / (freq_lo * freq_lo) - 1. / (freq * freq));   /* us */
  const double t_dm_hi = cfg->dm / 2.41e-10 * (1. / (freq * freq) - 1. / (freq_hi * freq_hi));
  unsigned long n_lo = (unsigned long) t_dm_lo * 1e-6 / tsamp, n_hi = (unsigned long) t_dm_hi * 1e-6 / tsamp;
  n_lo += (n_lo & 1); n_hi += (n_hi & 1);
  { unsigned long tmp = n_lo; n_lo = n_hi; n_hi = tmp; }
  h->n_lo = (long long) n_lo; h->n_hi = (long long) n_hi; h->n_dm = h->n_lo + h->n_hi;
  h->period = (long long) (cfg->pulse_period / tsamp);
  h->buflen = cfg->buflen;
  if (h->buflen == 0)          /* automatic: the reference's block (:203), doubled until the sweep fits twice */
    for (h->buflen = VFG_RATE / 4; h->buflen < 2 * h->n_dm + 2 && h->buflen < (1ll << 30); h->buflen *= 2) ;
  if (h->buflen < 16 || (h->buflen & 1) || h->period <= 0 || cfg->skip_period < 1) return vfg_fail (h, "bad configuration");
  if (h->buflen < 2 * h->n_dm + 2)
    return vfg_fail (h, "Buffer not long enough to perform dedispersion! (%lld samples, sweep %lld)", h->buflen, h->n_dm);
  h->new_samps = h->buflen - h->n_dm;
  CKG (cudaSetDevice (cfg->gpu_id));
  CKG (cudaStreamCreateWithFlags (&h->st, cudaStreamNonBlocking));
  const long long nc = h->buflen / 2 + 1;
  CKG (cudaMalloc ((void **) &h->d_c, (size_t) nc * sizeof (cufftComplex)));
  h->d_f = reinterpret_cast<float *> (h->d_c);
  CKG (cudaMalloc ((void **) &h->d_ker, (size_t) nc * sizeof (cufftComplex)));
  for (int p = 0; p < 2; ++p) {
    CKG (cudaMalloc ((void **) &h->d_out[p], (size_t) h->new_samps));
    CKG (cudaMalloc ((void **) &h->d_keep[p], (size_t) h->new_samps * sizeof (float)));
  }
  CKF (cufftPlan1d (&h->fwd, (int) h->buflen, CUFFT_R2C, 1));
  CKF (cufftPlan1d (&h->bwd, (int) h->buflen, CUFFT_C2R, 1));
  h->have_plans = 1;
  CKF (cufftSetStream (h->fwd, h->st));
  CKF (cufftSetStream (h->bwd, h->st));
  vfg_k_init_kernel<<<1184, 256, 0, h->st>>> (h->d_ker, cfg->dm, nc);
  CKG (cudaGetLastError ());
  CKG (cudaStreamSynchronize (h->st));
  return 0;
}

/* next block of both pols into d_out / d_keep */
static int vfg_next_block (vfg_handle *h)
{
  const long long base = h->block * h->new_samps;          /* first input sample of the block */
  for (int pol = 0; pol < 2; ++pol) {
    vfg_k_noise<<<1184, 256, 0, h->st>>> (h->d_f, h->buflen, base, pol, h->cfg.seed, h->period, h->cfg.skip_period,
                                          1.0f + h->cfg.ampl[pol]);
    CKG (cudaGetLastError ());
    CKF (cufftExecR2C (h->fwd, h->d_f, h->d_c));                     /* in place */
    vfg_k_chirp<<<1184, 256, 0, h->st>>> (h->d_c, h->d_ker, h->buflen / 2 + 1);
    CKG (cudaGetLastError ());
    CKF (cufftExecC2R (h->bwd, h->d_c, h->d_f));
    vfg_k_epilogue<<<1184, 256, 0, h->st>>> (h->d_f, h->n_lo, h->new_samps, base, pol, h->cfg.seed, h->cfg.add_rfi,
                                             h->d_out[pol], h->d_keep[pol]);
    CKG (cudaGetLastError ());
  }
  h->block++;
  h->avail = h->new_samps; h->rd = 0;
  return 0;
}

/* n samples of both pols to dst0/dst1 (device or host, kind says which) */
static int vfg_take (vfg_handle *h, uint8_t *dst0, uint8_t *dst1, size_t n, cudaMemcpyKind kind)
{
  size_t done = 0;
  while (done < n) {
    if (h->avail == 0) { int rc = vfg_next_block (h); if (rc) return rc; }
    const size_t take = (n - done) < (size_t) h->avail ? (n - done) : (size_t) h->avail;
    CKG (cudaMemcpyAsync (dst0 + done, h->d_out[0] + h->rd, take, kind, h->st));
    CKG (cudaMemcpyAsync (dst1 + done, h->d_out[1] + h->rd, take, kind, h->st));
    h->rd += (long long) take; h->avail -= (long long) take; done += take;
    if (h->avail == 0 && done < n) CKG (cudaStreamSynchronize (h->st));   /* d_out is about to be rewritten */
  }
  CKG (cudaStreamSynchronize (h->st));
  return 0;
}

int vfg_generate (vfg_handle *h, uint8_t *pol0, uint8_t *pol1, size_t n)
{
  if (!h || !pol0 || !pol1) return 1;
  CKG (cudaSetDevice (h->cfg.gpu_id));
  return vfg_take (h, pol0, pol1, n, cudaMemcpyDeviceToHost);
}

int vfg_last_block_f32 (vfg_handle *h, int pol, float *out)
{
  if (!h || !out || pol < 0 || pol > 1 || h->block == 0) return 1;
  CKG (cudaSetDevice (h->cfg.gpu_id));
  CKG (cudaMemcpy (out, h->d_keep[pol], (size_t) h->new_samps * sizeof (float), cudaMemcpyDeviceToHost));
  return 0;
}

int vfg_generate_vdif_second (vfg_handle *h, int station, uint32_t second, uint8_t *out)
{
  if (!h || !out) return 1;
  CKG (cudaSetDevice (h->cfg.gpu_id));
  if (!h->d_vdif) {
    for (int p = 0; p < 2; ++p) CKG (cudaMalloc ((void **) &h->d_sec[p], (size_t) VFG_RATE));
    CKG (cudaMalloc ((void **) &h->d_vdif, (size_t) VFG_FRAMES * 2 * VFG_VD_FRM));
  }
  int rc = vfg_take (h, h->d_sec[0], h->d_sec[1], (size_t) VFG_RATE, cudaMemcpyDeviceToDevice);
  if (rc) return rc;
  vfg_k_frames<<<VFG_FRAMES * 2, 256, 0, h->st>>> (h->d_sec[0], h->d_sec[1], h->d_vdif, second, station);
  CKG (cudaGetLastError ());
  CKG (cudaMemcpyAsync (out, h->d_vdif, (size_t) VFG_FRAMES * 2 * VFG_VD_FRM, cudaMemcpyDeviceToHost, h->st));
  CKG (cudaStreamSynchronize (h->st));
  return 0;
}

}   /* extern "C" */
