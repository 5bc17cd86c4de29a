/*
 * libvlitegen -- C ABI of the GPU baseband generator ("genbase" of the
 * reference, src/genbase.cu): Gaussian receiver noise with a pulsed amplitude,
 * dispersed coherently by overlap-save convolution with the chirp
 *   exp (i 2 pi DM / 2.41e-10 * f^2 / (f0^2 (f0 + f))),  f0 = 320 MHz
 * (src/genbase.cu:525-552, with its band-pass taper), side-band swapped to
 * VLITE's sense (:651-661), optional impulsive RFI (:671-687), digitised to
 * 8 bits (:690-708) and cut into VDIF frames (:443-486).
 *
 * a generator of test input, not part of the
 * baseband -> filterbank path.  The two large real FFTs are cuFFT (library
 * code, as in the reference, :271-274); the noise, profile, chirp-multiply,
 * epilogue and framing kernels are this library's.  It is a separate shared
 * object so that libvlitefast itself stays free of cuFFT.
 *
 * Differences from the reference, deliberate: the noise is a counter-based
 * generator (Philox-4x32-10 + Box-Muller), a pure function of (seed, pol,
 * absolute sample index) -- the reference draws from cuRAND's XORWOW stream,
 * which cannot be reproduced off the GPU -- so the overlap region of a block
 * is regenerated instead of kept (:372-389), and tests/ can restate the whole
 * generator in numpy.
 */
#ifndef VLITEGEN_H
#define VLITEGEN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vfg_config {
  double dm;             /* 30     -d  pc cm^-3                                        (src/genbase.cu:82)  */
  double pulse_period;   /* 0.5    -p  seconds                                          (:83)               */
  float ampl[2];         /* 0.05   -a  pulse amplitude above the noise, per pol; -s scales pol 1 (:84-85,171) */
  int skip_period;       /* 1      pulse every skip_period-th period                    (:91)               */
  int add_rfi;           /* 0      -f                                                    (:89)               */
  unsigned long long seed; /* 42   -r                                                    (:88)               */
  long long buflen;      /* 32000000 = VLITE_RATE/4 samples per convolution block (:203); even; 0 = that,
                            doubled until twice the dispersion sweep fits                                */
  int gpu_id;
  int reserved[7];
} vfg_config;

typedef struct vfg_handle vfg_handle;

int vfg_config_default (vfg_config *cfg);
/* kernel table, FFT plans, buffers.  Fails (non-zero) without a CUDA device or when the block is too
 * short for the dispersion sweep (:205-209). */
int vfg_create (const vfg_config *cfg, vfg_handle **out);
int vfg_destroy (vfg_handle *h);
const char *vfg_last_error (const vfg_handle *h);

/* samples each block delivers per pol (buflen minus the dispersion sweep) and the sweep itself */
long long vfg_block_samples (const vfg_handle *h);
long long vfg_sweep_samples (const vfg_handle *h);

/* The next n samples of both pols of the stream (which starts at sample 0 at vfg_create), to HOST
 * buffers.  Any n: blocks are generated as needed and the remainder is kept for the next call. */
int vfg_generate (vfg_handle *h, uint8_t *pol0, uint8_t *pol1, size_t n);
/* the pre-digitisation voltages of the block that produced the most recent samples (for tests):
 * vfg_block_samples floats of pol */
int vfg_last_block_f32 (vfg_handle *h, int pol, float *out);

/* One second of the stream as VDIF frames (25600 frame pairs, thread 0 then thread 1 per frame
 * number, 257 638 400 bytes; :443-486) to a HOST buffer. */
int vfg_generate_vdif_second (vfg_handle *h, int station, uint32_t second, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif
