/*
 * libvlitefast -- C ABI of the B200-native baseband -> filterbank chain.
 *
 * The reference (kerrm/vlite-fast) has no plugin or FFI interface: its kernels
 * are launched inline from main() in src/process_baseband.cu.  The boundary
 * declared here is therefore the data that crosses the device edge there
 * (H2D at src/process_baseband.cu:1117-1122, D2H at :1370-1375, persistent
 * bandpass state :700-709) plus the compile-time / command-line knobs that
 * select behaviour (src/process_baseband.h:16-55, CLI :45-63).  Every entry
 * point cites the reference lines it replaces.  All paths are relative to the
 * reference tree.
 *
 * Conventions: plain C types only; every function returns an int status
 * (VF_OK == 0) and never exits or throws (the reference's cudacheck throws
 * int 20, src/cuda_util.cu:4-12, and dadacheck exits, src/util.c:7-15: those
 * behaviours belong in the executable, not the library).  The caller owns host
 * buffers (pinned memory recommended, the reference uses cudaMallocHost,
 * :578-579,696); the library owns all device memory, its streams (it never
 * touches the default stream) and one bandpass state per (handle, antenna).
 * One caller thread per handle.
 *
 * There is no CPU fallback: without a CUDA device vf_create fails.
 */
#ifndef VLITEFAST_H
#define VLITEFAST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VF_ABI_VERSION 2

enum {
  VF_OK = 0,
  VF_ERR_ARG = 1,        /* bad argument or unsupported configuration     */
  VF_ERR_CUDA = 20,      /* CUDA runtime error (cudacheck's code, src/cuda_util.cu:10) */
  VF_ERR_NOMEM = 21,
  VF_ERR_STATE = 22,     /* call out of sequence (e.g. wait without submit) */
  VF_ERR_VDIF = 23,      /* frames outside the segment / malformed headers  */
  VF_ERR_NCCL = 24,
  VF_ERR_NODEV = 25      /* no usable CUDA device                           */
};

/* Geometry and behaviour.  Defaults (vf_config_default) are the reference's
 * compile-time values, src/process_baseband.h:16-55, and CLI defaults,
 * src/process_baseband.cu:343-355. */
typedef struct vf_config {
  int abi_version;     /* VF_ABI_VERSION                                           */
  int nfft;            /* 12500  NFFT (only value supported)                        */
  int nscrunch;        /* 8      NSCRUNCH                                           */
  int ffts_per_seg;    /* 1024   FFTS_PER_SEG = 128e6/10/12500; multiple of 8, <= 8192 */
  int nkurto;          /* 500    NKURTO                                             */
  int chanmin;         /* 2155   CHANMIN                                            */
  int chanmax;         /* 6250   CHANMAX (chanmax-chanmin+1 must be 4096)           */
  int nbit;            /* 2      NBIT, -b {2,4,8}                                   */
  int npol;            /* 1      -P {1,2}                                           */
  int rfi_mode;        /* 2      -r: 0 raw only, 1 excised only, 2 both            */
  int do_histo;        /* 0      DOHISTO: per-pol 256-bin sample histogram          */
  int keep_stats;      /* 0      WRITE_KURTO: keep pow/kur/dag(+_fb) for vf_get_stats */
  int keep_power;      /* 0      keep the f32 pre-digitisation tile (vf_get_power_f32, co-add) */
  int inject_frb;      /* 0      -i: allow vf_set_frb_injection                     */
  int gpu_id;          /* 0      -g                                                 */
  int n_antennas;      /* 1      antennas batched on this handle                    */
  int k1_threads;      /* 0      must be 0 (the product channeliser).  Testing builds of the library
                                 (libvlitefast_testing.so) also take round 1's kernels for A/B comparison:
                                 1 = the pipelined channeliser with the two-for-one 12500-point FFT,
                                 320, 512 or 640 = the monolithic channeliser with that many threads  */
  int power_segments;  /* 0      f32 tiles kept for this many consecutive segments (0 = 1): lets
                                 vf_coadd_batch reduce a whole second in one collective       */
  int max_batch_segments;/* 0    vf_process_device: consecutive segments per launch pair (0 = up to 16, 1 = one
                                 launch pair per segment); statistics dumps, histogram and FRB injection imply 1 */
  int numa_pin;       /* 0      1 = vf_create binds the calling thread (and so the pinned buffers it allocates
                                 afterwards) to the CPUs of the GPU's NUMA node, read from sysfs           */
  double dag_thresh;   /* 3.0    DAG_THRESH: a 500-sample block is excised when its statistic exceeds this  */
  double min_weight;   /* 0.2    MIN_WEIGHT: FFT time steps below this excision weight leave the scrunches  */
} vf_config;

typedef struct vf_handle vf_handle;

/* reference defaults into *cfg */
int vf_config_default (vf_config *cfg);

/* Device buffers, FFT tables, streams; bandpass zeroed.
 * Replaces the allocation block src/process_baseband.cu:472-475,578-709. */
int vf_create (const vf_config *cfg, vf_handle **out);
int vf_destroy (vf_handle *h);                       /* :1572-1602 */
const char *vf_strerror (int code);
const char *vf_last_error (const vf_handle *h);      /* detail of the last failure */

/* bytes one segment produces per stream:
 * ffts_per_seg/nscrunch * npol * 4096 * nbit/8   (trim, :666-675) */
size_t vf_segment_out_bytes (const vf_handle *h);
size_t vf_segment_in_samples (const vf_handle *h);   /* per pol: ffts_per_seg * 12500 */

/* One segment of one antenna, pol-planar HOST input -- exactly the reference's
 * device-edge contract (loop body src/process_baseband.cu:1108-1375).  antenna
 * (0 .. n_antennas-1) selects the bandpass state, statistics and kept tile.
 * fb_main receives the excised stream (rfi_mode 1, 2) or the raw stream
 * (rfi_mode 0); fb_raw the raw stream of rfi_mode 2 (may be NULL).
 * Synchronous: returns with the outputs on the host. */
int vf_process_segment (vf_handle *h, int antenna,
                        const uint8_t *pol0, const uint8_t *pol1, size_t nsamp_per_pol,
                        uint8_t *fb_main, uint8_t *fb_raw, size_t *nbytes);

/* Same for n_ant antennas in one launch sequence (the reference runs one
 * process per antenna, config/hosts:4-19).  Arrays of n_ant pointers; antenna
 * i of the call uses bandpass state i. */
int vf_process_batch (vf_handle *h, int n_ant,
                      const uint8_t *const *pol0, const uint8_t *const *pol1, size_t nsamp_per_pol,
                      uint8_t *const *fb_main, uint8_t *const *fb_raw);

/* Raw VDIF frames (5032 B: 32-B header + 5000 samples) of one segment, in any
 * order, both threads interleaved; depacketised on the GPU by thread id and
 * frame number (host loop src/process_baseband.cu:1015-1067).  first_frame is
 * the frame number (within the second) of the segment's first sample. */
int vf_process_vdif (vf_handle *h, int antenna, const void *frames, size_t nframes,
                     uint32_t first_frame, uint8_t *fb_main, uint8_t *fb_raw, size_t *nbytes);

/* Double-buffered asynchronous form of vf_process_batch: submit returns once
 * the copies and kernels are enqueued; slot in {0,1}.  Host buffers must stay
 * valid (and should be pinned) until vf_wait (slot) returns.  The reference is
 * fully synchronous (:1117-1122, :1370-1375). */
int vf_submit_async (vf_handle *h, int slot, int n_ant,
                     const uint8_t *const *pol0, const uint8_t *const *pol1, size_t nsamp_per_pol,
                     uint8_t *const *fb_main, uint8_t *const *fb_raw);
/* A block of n_seg consecutive segments of n_ant antennas from ONE host buffer laid out
 * [n_seg][n_ant][2][ffts_per_seg*12500] (one antenna: the ten 100-ms segments of a second back to back, pol 0 then
 * pol 1 in each): one copy in, one launch pair over the block, outputs to fb_main / fb_raw [n_seg][n_ant][out_bytes].
 * n_seg <= max_batch_segments.  Replaces n_seg rounds of src/process_baseband.cu:1108-1375. */
int vf_submit_block_async (vf_handle *h, int slot, int n_ant, int n_seg, const uint8_t *in, uint8_t *fb_main, uint8_t *fb_raw);
int vf_wait (vf_handle *h, int slot);
/* Asynchronous form of vf_process_vdif: the frames are DMA'd straight from the
 * caller's (pinned) memory -- e.g. a block of the input ring -- and must stay
 * valid until vf_wait (slot), which also reports frames outside the window. */
int vf_submit_vdif_async (vf_handle *h, int slot, int antenna, const void *frames, size_t nframes,
                          uint32_t first_frame, uint8_t *fb_main, uint8_t *fb_raw);

/* The same for a block of n_seg (1..16) consecutive segments -- n_seg = 10: a one-second block of the input ring
 * (src/process_baseband.cu:1015-1067 places every frame of the second before the ten segments are processed):
 * one copy, one depacketiser launch over the whole block, ONE launch pair over the n_seg segments; fb_main /
 * fb_raw receive n_seg x out_bytes.  expect_second >= 0: frames whose VDIF seconds field differs are skipped.
 * Skipped frames are counted, never fatal: vf_wait returns VF_ERR_VDIF as a warning with the outputs complete
 * (the gaps are zeros = dropped samples) and vf_vdif_report gives the counts: [0] outside the block, [1] other
 * second, [2] placed, [3] invalid bit, [4] frames of a complete block.  Needs a handle without keep_stats /
 * do_histo / inject_frb when n_seg > 1. */
int vf_submit_vdif_block_async (vf_handle *h, int slot, int antenna, const void *frames, size_t nframes,
                                uint32_t first_frame, long expect_second, int n_seg, uint8_t *fb_main, uint8_t *fb_raw);
int vf_vdif_report (vf_handle *h, int slot, unsigned int counts[5]);
/* Allocates, for both slots, everything vf_submit_vdif_block_async (n_seg) would otherwise allocate on its first call:
 * start-up work, as the reference's allocations are (src/process_baseband.cu:572-690), so that the first second of an
 * observation costs what the others do.  Optional. */
int vf_reserve_vdif_blocks (vf_handle *h, int n_seg);

/* Device-resident form: d_in is [n_ant][2][ffts_per_seg*12500] bytes on the
 * device (256-byte aligned), d_fb_main / d_fb_raw [n_ant][out_bytes].  Enqueued
 * on the handle's stream; vf_sync waits for it.  n_seg consecutive segments
 * are laid out back to back in all three buffers. */
int vf_process_device (vf_handle *h, int n_ant, int n_seg, const uint8_t *d_in,
                       uint8_t *d_fb_main, uint8_t *d_fb_raw);
int vf_sync (vf_handle *h);
/* elapsed device time between the start and end of the last vf_process_device
 * / vf_process_* call, from CUDA events on the library's stream (ms) */
int vf_last_elapsed_ms (vf_handle *h, float *total_ms, float *k1_ms, float *k2_ms);
/* the same for the last asynchronous submission on `slot` (vf_submit_*_async), after its vf_wait: total from the
 * start of the copy in to the end of the copy out; the per-stage accumulators of the reference's RT_PROFILE table
 * (src/process_baseband.cu:1538-1563) */
int vf_slot_elapsed_ms (vf_handle *h, int slot, float *total_ms, float *k1_ms, float *k2_ms);
/* Device-side stopwatch (CUDA events) over everything enqueued on the handle's streams -- segments and co-adds --
 * between the two calls; vf_timer_end waits for that work and returns the elapsed milliseconds. */
int vf_timer_begin (vf_handle *h);
int vf_timer_end (vf_handle *h, float *ms);
/* profiling aid: serial != 0 stops consecutive segments from overlapping, which makes k1_ms / k2_ms
 * above pure kernel execution times (with overlap they include waiting for SMs) */
int vf_set_serial (vf_handle *h, int serial);

/* Bind the calling thread to the CPUs local to the GPU's PCI function (sysfs local_cpulist) before it allocates
 * pinned buffers; vf_config.numa_pin makes vf_create do this.  cpulist (optional, cap bytes) receives what was
 * applied, "" if the platform exposes no locality. */
int vf_bind_thread_to_gpu (int gpu_id, char *cpulist, size_t cap);

/* pinned host memory helpers (cudaMallocHost, :578-579) */
int vf_host_alloc (void **p, size_t bytes);
int vf_host_free (void *p);
/* page-lock memory the caller already owns -- e.g. the data blocks of a shared-memory ring
 * (psrdada's dada_db -l, dada_cuda_dbregister) -- so that it can be DMA'd like vf_host_alloc memory */
int vf_host_register (void *p, size_t bytes);
int vf_host_unregister (void *p);

/* Statistics of the last segment of `antenna` (needs keep_stats / do_histo);
 * any pointer may be NULL.  Layouts as in the reference (:622-643):
 * pow/kur/dag [2][T*25], pow_fb/kur_fb/dag_fb [2][T], weights [2][T] (as left by
 * apply_kurtosis), histo [2][256].  = the WRITE_KURTO / DOHISTO dumps
 * (:1378-1393, :1444-1450). */
int vf_get_stats (vf_handle *h, int antenna, float *pow, float *kur, float *dag,
                  float *pow_fb, float *kur_fb, float *dag_fb, float *weights, uint32_t *histo);

/* Excision mask of the last segment: mask[t] bit j set = 500-sample block j of
 * FFT time step t was zeroed in both pols (apply_kurtosis, src/pb_kernels.cu:
 * 256-276; one shared decision per pol pair, :132).  [T] words. */
int vf_get_mask (vf_handle *h, int antenna, uint32_t *mask);

/* Pre-digitisation tile fft_ave of the last segment (:657-664), trimmed to the
 * kept channels: [npol][T/8][4096] floats; which = 0 main stream, 1 raw stream
 * of rfi_mode 2.  Needs keep_power. */
int vf_get_power_f32 (vf_handle *h, int antenna, int which, float *out);

/* Detected power |X|^2 of the last segment (before normalisation) for the kept
 * channels: [T][4096][2] (pol0, pol1).  which = 0 main, 1 raw.  Time steps of
 * the excised stream that had nothing excised are identical to the raw stream. */
int vf_get_detected_power (vf_handle *h, int antenna, int which, float *out);

/* running bandpass, [2][4096] (pol, chan); which = 0 main, 1 raw (:700-709) */
int vf_get_bandpass (vf_handle *h, int antenna, int which, float *out);
int vf_set_bandpass (vf_handle *h, int antenna, int which, const float *in);
int vf_reset_bandpass (vf_handle *h, int antenna);   /* antenna < 0: all */

/* FRB injection (-i, :1098-1101, :1231-1251; kernels src/pb_kernels.cu:338-391):
 * nfft_since_frb < 0 disables.  dm in pc cm^-3 (reference: 80), width in FFT
 * steps (2e-3*10*1024), amp 1.05. */
int vf_set_frb_injection (vf_handle *h, int nfft_since_frb, float dm, float width, float amp);

/* ---- co-add (replaces scripts/start_coadd + external agdadacoadd, and the
 * per-segment ring write at src/process_baseband.cu:1416-1422) -------------
 * Sum of the main-stream f32 tiles of the antennas of the last launch (last
 * segment), reduced over `nranks` processes with NCCL together with the count
 * of antennas that kept each scrunched row (an antenna whose row was zeroed by
 * the excision weights, src/pb_kernels.cu:616-623, does not count), divided by
 * sqrt (count) and digitised on rank `root`.  total_antennas is informative
 * (the divisor is the reduced count).  Needs keep_power. */
int vf_coadd_init (vf_handle *h, int nranks, int rank, const void *nccl_unique_id /*128 B*/);
int vf_coadd_unique_id (void *out128);               /* ncclGetUniqueId */
int vf_coadd_segment (vf_handle *h, int root, int total_antennas, uint8_t *fb_coadd /*host, root only*/,
                      float *sum_f32 /*host, optional, root only*/);
/* The same for the last n_seg segments (n_seg <= power_segments) in ONE reduce:
 * the exchange is 2 MiB per antenna-segment, i.e. latency bound, so a second
 * of segments is reduced at once.  Outputs are [n_seg][...] in segment order.
 * wait == 0: returns once enqueued on the library's stream (vf_sync waits). */
int vf_coadd_batch (vf_handle *h, int root, int total_antennas, int n_seg, uint8_t *fb_coadd, float *sum_f32, int wait);

#ifdef __cplusplus
}
#endif
#endif
