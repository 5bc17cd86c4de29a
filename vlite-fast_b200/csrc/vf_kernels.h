/*
 * Internal kernel interface of libvlitefast (not part of the C ABI).
 * Geometry follows src/process_baseband.h:16-55 of the reference.
 */
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <cuda_runtime.h>
#include "vf_fft12500.cuh"
#include "vf_fft6250.cuh"

#define VF_NCHAN_FFT   6251     /* NFFT/2+1, src/process_baseband.h:22 */
#define VF_NSCRUNCH    8        /* :24 */
#define VF_NKURTO      500      /* :35 */
#define VF_NSUB        25       /* NFFT/NKURTO */
#define VF_CHANMIN     2155     /* :53 */
#define VF_CHANMAX     6250     /* :54 */
#define VF_NCHANOUT    4096
#define VF_VD_FRM      5032     /* :16 */
#define VF_VD_DAT      5000     /* :17 */
#define VF_FRAME_RATE  25600    /* :19 */

/* Detected-power tile between the two kernels, per antenna: [4096 / VF_PBLK][T][VF_PBLK] float2.
 * VF_PBLK = 4096 is the plain [T][4096] tile.  A blocked tile (VF_PBLK = 16: the 16 channels of a
 * normaliser CTA contiguous over all time steps) was measured on B200 and rejected: the normaliser
 * was no faster (its staging is latency, not locality, bound) and the channeliser's scattered
 * 128-byte stores cost it 2 us per segment. */
#define VF_PBLK        VF_NCHANOUT
#define VF_PIDX(T, c)  (((size_t) ((c) / VF_PBLK) * (size_t) (T)) * VF_PBLK + ((c) % VF_PBLK))   /* relative to time step t's base (t * VF_PBLK) */

#define VF_TW6_LEN     (250 + 250 + 240 + VF6_M)   /* vf6_tables in one array: tw1 | tw5 | tw250 | tws */
#define VF_WIN         12512    /* 16-byte aligned window that covers any 12500-sample block */

/* Channeliser: one work item = one FFT time step of one antenna (both pols). */
struct vf_k1_params {
  const uint8_t *in;          /* [n_ant][2][T*12500] samples, every pol 16-byte aligned */
  size_t ant_stride, pol_stride;
  int T, n_ant, rfi_mode;
  float2 *P_raw, *P_kur;      /* [n_ant][T][4096] (pol0, pol1) detected power (see VF_PBLK) */
  float *w;                   /* [n_ant][T] excision weight (shared by both pols) */
  uint32_t *mask;             /* [n_ant][T] bit j = sub-block j zeroed */
  float *pw, *kur, *dag;      /* optional [n_ant][2][T*25] */
  float *pw_fb, *kur_fb, *dag_fb;  /* optional [n_ant][2][T] */
  unsigned int *histo;        /* optional [n_ant][2][256] */
  vf_fft_tables tb;           /* two-for-one 12500-point FFT (testing builds: the monolithic kernel) */
  vf6_tables tb6;             /* 6250-point real-input FFT of the product kernel */
  double dagc[5], dagc_fb[5]; /* mu1, A, Z1, Z2, Z3 for N = 500 and N = 12500 */
  double dag_thresh;          /* DAG_THRESH, src/process_baseband.h:42 */
  const float *wtab;          /* [26] device: k-fold float sum of float(500)/12500 */
  const float *frb_delays;    /* optional [6251]; FRB injection, src/pb_kernels.cu:348-391 */
  int nfft_since_frb;
  float frb_width, frb_amp;
  unsigned int *work_counter; /* pipelined kernel: items are drawn from this counter ... */
  unsigned int work_base;     /* ... whose value at launch was work_base (it advances by n_items + 2 * grid) */
};

/* Normaliser: bandpass IIR + pscrunch + tscrunch + select/digitise. */
struct vf_k2_params {
  const float2 *P_raw, *P_kur;
  const float *w;
  const uint32_t *mask;
  float2 *bp_raw, *bp_kur;    /* [n_ant][4096] running bandpass (pol0, pol1) */
  int T, n_ant, rfi_mode, npol, nbit;
  int n_seg;                  /* consecutive segments in this launch: data of (seg, ant) at index seg * n_ant + ant */
  float bp_scale;
  uint8_t *out_main, *out_raw;/* [n_ant][out_bytes] */
  size_t out_stride;
  float *ave_main, *ave_raw;  /* optional ring of ave_nseg tiles [n_ant][npol][T/8][4096]; segment seg of the launch goes */
  long ave_seg0; int ave_nseg; size_t ave_seg_elems;   /* to tile (ave_seg0 + seg) % ave_nseg, ave_seg_elems floats apart */
  float *rowok;               /* optional, beside the tile ring: [ave_nseg][n_ant_total][T/8], 1 = the scrunched row of the main
                                 stream was kept, 0 = zeroed by tscrunch_weights (:622-623); the co-add's count */
  size_t rowok_seg_elems;
  double min_weight;          /* MIN_WEIGHT, src/process_baseband.h:45 */
  long long *trace;           /* TESTING BUILDS: clock64 stamps of the first CTA, [stream][chunk][6], else NULL */
  float clip_floor;           /* set by vf_launch_k2: lower bound of 11 bp over a chunk / bp at its start (clip pre-test) */
  float min_weight_f;         /* smallest float >= min_weight: for a float w, (double) w >= min_weight <=> w >= min_weight_f */
};

struct vf_depack_params {
  const uint8_t *frames;      /* [nframes][5032] */
  size_t nframes;
  uint8_t *out;               /* [2][T*12500] */
  size_t pol_stride;
  size_t seg_stride;          /* out is [segment][2][T*12500]: bytes from one segment to the next */
  long long frames_per_seg;   /* frames per pol per segment */
  long long frame0;           /* frame number (within second) of sample 0 of out */
  long long nframes_per_pol;  /* frames that fit in out (all segments) */
  long expect_second;         /* VDIF seconds field every frame must carry, < 0 = not checked */
  unsigned int *bad;          /* [0] outside the window, [1] another second, [2] placed, [3] invalid bit */
};

/* both co-add kernels cover the n_seg segments of a batch in one launch (blockIdx.y = segment of the batch);
 * segment i of the batch lives in slot (seg0 + i) % nring of the ring of kept tiles */
struct vf_coadd_params {
  const float *sum;           /* [n_seg][npol][T/8][4096] summed tiles */
  const float *cnt;           /* [n_seg][T/8] antennas that contributed to the row (the same for every channel and pol) */
  int ntime, npol, nbit, n_seg;
  uint8_t *out;               /* [n_seg][out bytes] */
};

/* local part of the co-add: sum of the tiles of this handle's antennas and the count of kept rows */
struct vf_coadd_local_params {
  const float *tiles;         /* ring [nring][n_ant_total][npol][T/8][4096] */
  const float *rowok;         /* ring [nring][n_ant_total][T/8] */
  long seg0; int nring, n_ant_total;
  int n_ant, ntime, npol, n_seg;
  float *sum;                 /* [n_seg][npol][T/8][4096] */
  float *cnt;                 /* [n_seg][T/8] */
};

cudaError_t vf_launch_k1 (const vf_k1_params &p, int grid, int threads, cudaStream_t s);
cudaError_t vf_launch_k2 (const vf_k2_params &p, cudaStream_t s);
cudaError_t vf_launch_depack (const vf_depack_params &p, cudaStream_t s);
cudaError_t vf_launch_coadd (const vf_coadd_params &p, cudaStream_t s);
cudaError_t vf_launch_coadd_local (const vf_coadd_local_params &p, cudaStream_t s);
#ifdef VF_TESTING
cudaError_t vf_launch_debug_div (const float *p, const float *b, float *q_packed, float *q_ref, size_t n, cudaStream_t s);
#endif
cudaError_t vf_k1_configure (void);
