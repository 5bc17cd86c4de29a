/*
 * Deterministic genbase-style synthetic baseband (see vf_genbase.c).  The
 * reference's generator is src/genbase.cu; its cuRAND stream and float FFT
 * dispersion cannot be reproduced bit for bit off the GPU, so the test input
 * of this repo is a pure integer function of (seed, antenna, pol, sample).
 */
#ifndef VF_GENBASE_H
#define VF_GENBASE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  uint64_t seed;
  int pulse_period, pulse_width;   /* samples; period 0 = no pulse        */
  int pulse_amp_q8[2];             /* fractional amplitude x256, per pol  */
  int rfi_period, rfi_width;       /* impulsive RFI window, samples       */
  int rfi_amp;                     /* uniform in [-amp, amp]; 0 = off     */
  int rfi_burst_every;             /* <=1: every 6250-sample stretch; n: 1 in n */
  int tone_step, tone_amp;         /* tone at step/64 cycles per sample   */
  int drop_period, drop_len;       /* frames f with f % period < len are zero */
  int drop_pol_skew;               /* frame offset applied to pol 1       */
} vf_gen_params;

void vf_gen_defaults (vf_gen_params *g);
/* n samples of (antenna, pol) starting at absolute sample index sample0 */
void vf_gen_samples (const vf_gen_params *g, int antenna, int pol,
                     uint64_t sample0, size_t n, uint8_t *out);
/* nframes frame pairs (thread 0 then thread 1) of one second, starting at
 * frame first_frame; 5032 bytes per frame.  Returns bytes written. */
size_t vf_gen_vdif_second (const vf_gen_params *g, int antenna, uint32_t second,
                           uint32_t first_frame, uint32_t nframes, uint8_t *out);
/* one full second of frames (257 638 400 bytes); scratch: 256 000 000 bytes */
size_t vf_gen_vdif_block (const vf_gen_params *g, int antenna, uint32_t second, uint8_t *scratch, uint8_t *out);
#ifdef __cplusplus
}
#endif
#endif
