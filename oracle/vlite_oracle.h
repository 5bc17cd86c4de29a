/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * CPU restatement ("oracle A") of VLITE-Fast's baseband -> filterbank chain
 * (src/pb_kernels.cu kernels launched in the order of
 * src/process_baseband.cu:1108-1375).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it, and only as
 * the checker or the reported CPU baseline -- never inside the product path.
 *
 * Parity pinning: the reference ships no golden vectors.
 * This restatement is pinned against outputs of the reference's own kernels
 * (oracle/_ref, built from /root/reference/src/pb_kernels.cu unmodified and
 * run on a B200) that are committed under tests/golden/, see
 * tests/test_oracle_golden.py and scripts/make_golden.py.
 */
#ifndef VLITE_ORACLE_H
#define VLITE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* geometry constants, src/process_baseband.h:16-55 */
#define ORC_NFFT      12500
#define ORC_NCHAN     6251          /* NFFT/2+1 */
#define ORC_NSCRUNCH  8
#define ORC_NKURTO    500
#define ORC_CHANMIN   2155
#define ORC_CHANMAX   6250
#define ORC_NCHANOUT  4096          /* CHANMAX-CHANMIN+1 */
#define ORC_SUB       25            /* NFFT/NKURTO */

typedef struct orc_fft_plan orc_fft_plan;
orc_fft_plan *orc_fft_plan_create (int n);
void orc_fft_plan_destroy (orc_fft_plan *pl);
size_t orc_fft_scratch_floats (const orc_fft_plan *pl);
void orc_rfft (const orc_fft_plan *pl, const float *in, float *out, float *scratch);

typedef struct orc_chain orc_chain;

/* ffts_per_seg: FFTs per polarisation per segment (reference: 1024; must be a
 * multiple of 8).  nbit in {2,4,8}; npol in {1,2}; rfi_mode in {0,1,2}. */
orc_chain *orc_create (int ffts_per_seg, int nbit, int npol, int rfi_mode, int nthreads);
void orc_destroy (orc_chain *c);
void orc_reset_bandpass (orc_chain *c);
void orc_set_bandpass (orc_chain *c, int which, const float *bp);   /* which: 0 main, 1 raw */
size_t orc_out_bytes (const orc_chain *c);

/* One segment.  pol0/pol1: ffts_per_seg*12500 unsigned samples each.
 * fb_main: excised stream (rfi_mode 1,2) or raw stream (rfi_mode 0).
 * fb_raw: raw stream, written only when rfi_mode == 2 (may be NULL).
 * Returns 0 on success. */
int orc_process_segment (orc_chain *c, const uint8_t *pol0, const uint8_t *pol1,
                         uint8_t *fb_main, uint8_t *fb_raw);

/* intermediates of the last segment (owned by the chain) */
const float *orc_get_pow (const orc_chain *c);       /* [2][T*25]              */
const float *orc_get_kur (const orc_chain *c);       /* [2][T*25]              */
const float *orc_get_dag (const orc_chain *c);       /* [2][T*25] (duplicated) */
const float *orc_get_pow_fb (const orc_chain *c);    /* [2][T]                 */
const float *orc_get_kur_fb (const orc_chain *c);    /* [2][T]                 */
const float *orc_get_dag_fb (const orc_chain *c);    /* [2][T]                 */
const float *orc_get_weights (const orc_chain *c);   /* [2][T] as left by apply_kurtosis */
const float *orc_get_ave_main (const orc_chain *c);  /* fft_ave of main stream [npol][T/8][6251] */
const float *orc_get_ave_raw (const orc_chain *c);   /* fft_ave of raw stream (mode 2)          */
const float *orc_get_power_main (const orc_chain *c);/* |X|^2 (before /w) [2][T][6251], main stream */
const float *orc_get_power_raw (const orc_chain *c); /* |X|^2 [2][T][6251], raw stream (mode 2)     */
const float *orc_get_bp_main (const orc_chain *c);   /* [2][6251] */
const float *orc_get_bp_raw (const orc_chain *c);    /* [2][6251] */
const uint32_t *orc_get_histo (const orc_chain *c);  /* [2][256] */

/* stand-alone stage entry points (for unit tests) */
void orc_stage_convert (const uint8_t *u, float *x, size_t n);
void orc_stage_kurtosis (const float *x, float *pw, float *kur, size_t nblock);
void orc_stage_dagostino (const float *kur, float *dag, size_t n, int nsamp);
void orc_digitise (const float *ave, uint8_t *out, int ntime, int npol, int nbit);
void orc_dagostino_constants (int nsamp, double out[5]);

#ifdef __cplusplus
}
#endif
#endif
