/*
 * Deterministic, integer-only synthetic baseband in the spirit of the
 * reference's genbase (src/genbase.cu:79-708): Gaussian-like receiver noise
 * with a pulsed amplitude on top (set_profile, src/genbase.cu:554-585),
 * optional impulsive RFI (add_rfi, :671-687: a uniform deviate added during a
 * fraction of every RFI period), an optional narrow-band tone, and dropped
 * frames as the writer leaves them (payload bytes 0, src/writer.c:362,674-687).
 * genbase draws from cuRAND, which cannot be reproduced off the GPU, and
 * disperses with a float FFT whose rounding is platform dependent; golden
 * fixtures need an input that is bit-identical everywhere, so every sample
 * here is a pure function of (seed, pol, sample index) in integer arithmetic.
 *
 * noise: popcount of 128 hashed bits is Binomial(128, 1/2) (sigma 5.657,
 * excess kurtosis -1/64); 3*(pc-64) + d, d uniform in {-1,0,1}, has sigma
 * 16.99 counts (genbase digitises to sigma 1/(2*0.02957) = 16.91, :700-706)
 * about a mean of 128.
 */
#include <string.h>
#include "vf_genbase.h"

static inline uint64_t mix64 (uint64_t z)
{
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

/* 64-entry sine table, round(127*sin(2 pi i/64)) */
static const int8_t sin64[64] = {
  0, 12, 25, 37, 49, 60, 71, 81, 90, 98, 106, 112, 117, 122, 125, 126,
  127, 126, 125, 122, 117, 112, 106, 98, 90, 81, 71, 60, 49, 37, 25, 12,
  0, -12, -25, -37, -49, -60, -71, -81, -90, -98, -106, -112, -117, -122, -125, -126,
  -127, -126, -125, -122, -117, -112, -106, -98, -90, -81, -71, -60, -49, -37, -25, -12 };

void vf_gen_defaults (vf_gen_params *g)
{
  memset (g, 0, sizeof (*g));
  g->seed = 102;                 /* scripts/baseband_test:21 uses -r 102 */
  g->pulse_period = 64000000;    /* 0.5 s (-p 0.5) */
  g->pulse_width = 1920000;      /* 3 % duty, src/genbase.cu:576 */
  g->pulse_amp_q8[0] = 13;       /* 0.05 * 256 (-a 0.05) */
  g->pulse_amp_q8[1] = 1;        /* x 0.1 on pol 1 (-s 0.1) */
  g->rfi_period = 1446;          /* 11.3 us at 128 MS/s, src/genbase.cu:671-687 */
  g->rfi_width = 145;            /* 10 % of it */
  g->rfi_amp = 0;                /* off unless asked (-f) */
}

void vf_gen_samples (const vf_gen_params *g, int antenna, int pol,
                      uint64_t sample0, size_t n, uint8_t *out)
{
  const uint64_t key = mix64 (g->seed * 0x100000001B3ull + (uint64_t) antenna * 2 + (uint64_t) pol);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) {
    const uint64_t s = sample0 + i;
    if (g->drop_period > 0) {
      uint64_t frame = s / 5000;
      if ((frame + (uint64_t) pol * g->drop_pol_skew) % (uint64_t) g->drop_period < (uint64_t) g->drop_len) {
        out[i] = 0;
        continue;
      }
    }
    const uint64_t h1 = mix64 (key ^ (s * 2));
    const uint64_t h2 = mix64 (key ^ (s * 2 + 1));
    const uint64_t h3 = mix64 (h1 ^ (h2 << 1 | h2 >> 63));
    int dev = 3 * (__builtin_popcountll (h1) + __builtin_popcountll (h2) - 64)
              + (int) (h3 % 3) - 1;
    if (g->pulse_period > 0 && (s % (uint64_t) g->pulse_period) < (uint64_t) g->pulse_width)
      dev += (dev * g->pulse_amp_q8[pol]) / 256;
    if (g->rfi_amp > 0 && g->rfi_period > 0 &&
        (s % (uint64_t) g->rfi_period) < (uint64_t) g->rfi_width) {
      /* bursts only in every rfi_burst_every-th kurtosis-sized stretch, so
       * that the mask is a mixture of clean and dirty blocks */
      uint64_t stretch = s / 6250;
      if (g->rfi_burst_every <= 1 || mix64 (key ^ (stretch + 0x5555)) % (uint64_t) g->rfi_burst_every == 0)
        dev += (int) ((h3 >> 8) % (uint64_t) (2 * g->rfi_amp + 1)) - g->rfi_amp;
    }
    if (g->tone_amp > 0 && g->tone_step > 0)
      dev += (g->tone_amp * sin64[(s * (uint64_t) g->tone_step) & 63]) / 127;
    int v = 128 + dev;
    if (v < 1) v = 1;            /* 0 is reserved for dropped data (src/pb_kernels.cu:28-29) */
    if (v > 255) v = 255;
    out[i] = (uint8_t) v;
  }
}

/* One second of VDIF frames (src/genbase.cu:443-486): for each frame number,
 * thread 0 then thread 1; 32-byte header + 5000 payload bytes.  Header bit
 * layout per analysis/baseband.py:19-28.  Returns bytes written. */
size_t vf_gen_vdif_second (const vf_gen_params *g, int antenna, uint32_t second,
                            uint32_t first_frame, uint32_t nframes, uint8_t *out)
{
  uint8_t *p = out;
  for (uint32_t f = first_frame; f < first_frame + nframes; ++f)
    for (int th = 0; th < 2; ++th) {
      uint32_t w[8] = {0};
      w[0] = second & 0x3FFFFFFFu;
      w[1] = (f & 0xFFFFFFu) | (30u << 24);           /* ref epoch 30 */
      w[2] = (5032u / 8) & 0xFFFFFFu;                 /* frame length / 8 */
      w[3] = ((uint32_t) (antenna & 0xFFFF)) | ((uint32_t) th << 16) | (7u << 26); /* 8 bits/sample - 1 */
      memcpy (p, w, 32);
      vf_gen_samples (g, antenna, th, ((uint64_t) second * 25600 + f) * 5000, 5000, p + 32);
      p += 5032;
    }
  return (size_t) (p - out);
}

/* One full second (25600 frame pairs, 257 638 400 bytes) of antenna `antenna`:
 * the samples of each pol are generated in one parallel sweep and then cut
 * into frames (thread 0 then thread 1 per frame number, src/genbase.cu:443-486).
 * scratch: 2 * 128 000 000 bytes.  Returns bytes written. */
size_t vf_gen_vdif_block (const vf_gen_params *g, int antenna, uint32_t second, uint8_t *scratch, uint8_t *out)
{
  const size_t rate = 128000000;
  for (int pol = 0; pol < 2; ++pol)
    vf_gen_samples (g, antenna, pol, (uint64_t) second * rate, rate, scratch + (size_t) pol * rate);
#pragma omp parallel for schedule(static)
  for (long f = 0; f < 25600; ++f)
    for (int th = 0; th < 2; ++th) {
      uint8_t *p = out + ((size_t) f * 2 + th) * 5032;
      uint32_t w[8] = {0};
      w[0] = second & 0x3FFFFFFFu;
      w[1] = ((uint32_t) f & 0xFFFFFFu) | (30u << 24);
      w[2] = (5032u / 8) & 0xFFFFFFu;
      w[3] = ((uint32_t) (antenna & 0xFFFF)) | ((uint32_t) th << 16) | (7u << 26);
      memcpy (p, w, 32);
      memcpy (p + 32, scratch + (size_t) th * rate + (size_t) f * 5000, 5000);
    }
  return (size_t) 25600 * 2 * 5032;
}
