/*
 * process_baseband -- host driver of the baseband -> filterbank chain.
 *
 * Plain-C counterpart of the reference's main() (src/process_baseband.cu:334-
 * 1615): command-line flags (:45-63, :358-470), wait for an observation header
 * on the input ring (:799-832), check the alignment of the first VDIF frame
 * (:837-849), open the SIGPROC files (:852-998), then per second of data ten
 * 100 ms segments through the GPU (:1108-1458), log the processing rate
 * (:1461-1477, :1534-1536).  All signal processing is in libvlitefast
 * (include/vlitefast.h); this file only moves frames and bytes.
 *
 * Differences from the reference, all deliberate:
 *  - psrdada is not available: the input ring is the shim of vf_ring.h, either
 *    a SysV shared-memory ring made by vf_dada_db and fed by another process
 *    (-k KEY, the reference's arrangement, scripts/start_dada:13), or an
 *    in-process one fed by a thread that replays a VDIF file (-f, the
 *    reference's readbase) or synthetic seconds (-S, the reference's genbase);
 *  - frames are not depacketised on the host (:1034-1035): one-second ring
 *    blocks live in pinned memory and each segment is DMA'd straight from the
 *    block and depacketised on the GPU (vf_submit_vdif_async), double buffered;
 *  - the final complete second of an observation is processed (the reference
 *    dispatches second N only when a frame of N+1 arrives, :1022-1067, and so
 *    drops it).
 */
#define _GNU_SOURCE
#include <errno.h>
#include <getopt.h>
#include <pthread.h>
#include <signal.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include "vlitefast.h"
#include "vf_control.h"
#include "vf_genbase.h"
#include "vf_ring.h"
#include "vf_sigproc.h"
#include "vf_vdif.h"

#define SEG_PER_SEC 10
#define FRAMES_PER_SEC_2POL (2 * VF_FRAME_RATE)                 /* 51200 */
#define SEC_BYTES ((size_t) FRAMES_PER_SEC_2POL * VF_VD_FRM)    /* 257 638 400, scripts/start_writer:12 */

static volatile sig_atomic_t g_quit = 0;
static FILE *g_log = NULL;
static int g_stdout = 0;

static void on_signal (int sig) { (void) sig; g_quit = 1; }

static void logmsg (const char *level, const char *fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start (ap, fmt);
  vsnprintf (buf, sizeof (buf), fmt, ap);
  va_end (ap);
  time_t now = time (NULL);
  struct tm tm;
  gmtime_r (&now, &tm);
  char ts[32];
  strftime (ts, sizeof (ts), "%Y-%m-%d-%H:%M:%S", &tm);
  if (g_log) { fprintf (g_log, "[%s] %s %s", ts, level, buf); fflush (g_log); }
  if (g_stdout || !g_log) { fprintf (stderr, "[%s] %s %s", ts, level, buf); }
}

static double now_s (void)
{
  struct timespec ts;
  clock_gettime (CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void usage (void)
{
  fprintf (stdout,
    "Usage: process_baseband [options]\n"
    "  -k KEY    read from the shared-memory ring with this (hexadecimal) key, made by vf_dada_db\n"
    "  -K KEY    write the main stream to this ring: the first 10 s at once, then second by second (heimdall)\n"
    "  -C KEY    write the main stream to this ring segment by segment (co-adder)\n"
    "  -f FILE   replay a VDIF file into the input ring (readbase)\n"
    "  -S N      synthesise N distinct seconds of baseband (genbase style) instead of -f\n"
    "  -L M      replay the synthetic seconds M times (stream of N*M seconds) [1]\n"
    "  -e SEED   generator seed [102];  -F  add impulsive RFI to the synthetic data\n"
    "  -a ID     station id of the synthetic stream [1]\n"
    "  -n NBUF   ring blocks of one second each [N for -S, else 4]\n"
    "  -D DIR    directory of the output filterbank files [.]\n"
    "  -w 0|1    write filterbank files [1]\n"
    "  -b NBIT   2, 4 or 8 bit output [2]\n"
    "  -P NPOL   1 (summed) or 2 polarisations [1]\n"
    "  -r MODE   RFI excision: 0 off, 1 excised stream only, 2 both streams [2]\n"
    "  -i        inject a DM 80 FRB every 60 s\n"
    "  -R        keep the running bandpass from one observation to the next, as the reference does\n"
    "            (src/process_baseband.cu:700-709); default: every observation starts from its own mean\n"
    "  -W        WRITE_KURTO: dump kurtosis, block kurtosis and weights of every segment beside the filterbank\n"
    "            (.kurto, .block_kurto, .weights; src/process_baseband.cu:1385-1393,1446-1450)\n"
    "  -H        DOHISTO: dump the per-polarisation sample histograms of every segment (.histo, :1378-1383,1444)\n"
    "  -T        print per-stage device times at the end of every observation (PROFILE, :1538-1558)\n"
    "  -s        process a single observation then quit\n"
    "  -g GPU    CUDA device [0]\n"
    "  -o        log to stderr as well\n"
    "  -l FILE   log file\n"
    "  -j        print a one-line JSON summary on stdout at exit\n"
    "  -m GROUP[:PORT]  listen for one-character commands (Q = quit) on this UDP multicast group\n"
    "            (the reference listens on 224.3.29.71:20000); a unicast address binds the port only\n");
}

typedef struct {
  vf_ring *ring;
  const char *file;
  int synth_n, loops, station;
  uint32_t second0;
  int rc;
} feeder_args;

/* file -> ring, one frame-aligned block at a time (src/readbase.c:52-105) */
static void *feeder_file (void *vp)
{
  feeder_args *a = (feeder_args *) vp;
  FILE *fp = fopen (a->file, "rb");
  if (!fp) { logmsg ("ERR", "cannot open %s: %s\n", a->file, strerror (errno)); a->rc = 1; vf_ring_shutdown (a->ring); return NULL; }
  char hdr[VF_RING_HEADER_SIZE] = "";
  /* the keys readbase sets, src/readbase.c:65-86 */
  vf_ascii_header_set (hdr, sizeof (hdr), "NAME", "%s", "REPLAY");
  vf_ascii_header_set (hdr, sizeof (hdr), "STATIONID", "%d", a->station);
  vf_ascii_header_set (hdr, sizeof (hdr), "NCHAN", "%d", 1);
  vf_ascii_header_set (hdr, sizeof (hdr), "BANDWIDTH", "%lf", -64.0);
  vf_ascii_header_set (hdr, sizeof (hdr), "CFREQ", "%lf", 352.0);
  vf_ascii_header_set (hdr, sizeof (hdr), "NPOL", "%d", 2);
  vf_ascii_header_set (hdr, sizeof (hdr), "NBIT", "%d", 8);
  vf_ascii_header_set (hdr, sizeof (hdr), "RA", "%lf", 0.87180);
  vf_ascii_header_set (hdr, sizeof (hdr), "DEC", "%lf", 0.72452);
  if (vf_ring_header_write (a->ring, hdr)) { fclose (fp); return NULL; }
  for (;;) {
    void *blk = vf_ring_block_write_open (a->ring);
    if (!blk) break;
    size_t n = fread (blk, 1, vf_ring_get_bufsz (a->ring), fp);
    n -= n % VF_VD_FRM;
    if (n == 0) { vf_ring_end_of_data (a->ring); break; }
    vf_ring_block_write_close (a->ring, n);
    if (g_quit) { vf_ring_end_of_data (a->ring); break; }
  }
  fclose (fp);
  return NULL;
}

/* synthetic seconds already sit in the ring blocks (block b holds template
 * second b % synth_n): publishing a block only needs its VDIF seconds patched */
static void *feeder_synth (void *vp)
{
  feeder_args *a = (feeder_args *) vp;
  char hdr[VF_RING_HEADER_SIZE] = "";
  /* the keys genbase sets, src/genbase.cu:331-353 */
  vf_ascii_header_set (hdr, sizeof (hdr), "NAME", "%s", "GENBASE");
  vf_ascii_header_set (hdr, sizeof (hdr), "STATIONID", "%d", a->station);
  vf_ascii_header_set (hdr, sizeof (hdr), "NCHAN", "%d", 1);
  vf_ascii_header_set (hdr, sizeof (hdr), "BANDWIDTH", "%lf", -64.0);
  vf_ascii_header_set (hdr, sizeof (hdr), "CFREQ", "%lf", 352.0);
  vf_ascii_header_set (hdr, sizeof (hdr), "NPOL", "%d", 2);
  vf_ascii_header_set (hdr, sizeof (hdr), "NBIT", "%d", 8);
  vf_ascii_header_set (hdr, sizeof (hdr), "RA", "%lf", 0.87180);
  vf_ascii_header_set (hdr, sizeof (hdr), "DEC", "%lf", 0.72452);
  if (vf_ring_header_write (a->ring, hdr)) return NULL;
  const long total = (long) a->synth_n * a->loops;
  for (long s = 0; s < total && !g_quit; ++s) {
    unsigned char *blk = (unsigned char *) vf_ring_block_write_open (a->ring);
    if (!blk) return NULL;
    const uint32_t sec = a->second0 + (uint32_t) s;
    /* 51 200 header words 5 KB apart: one cache miss each, so spread them over the cores or the
     * feeder, not the GPU, sets the pace of the synthetic stream */
#pragma omp parallel for schedule(static) num_threads(8)
    for (long f = 0; f < FRAMES_PER_SEC_2POL; ++f) {
      uint32_t *w0 = (uint32_t *) (blk + (size_t) f * VF_VD_FRM);
      *w0 = sec & 0x3FFFFFFFu;
    }
    vf_ring_block_write_close (a->ring, SEC_BYTES);
  }
  vf_ring_end_of_data (a->ring);
  return NULL;
}

/* Output rings of the reference (:306-332, :1416-1422, :1482-1494).  check_buffer: a ring with a
 * single free block left means that the consumer has stalled; the reference gives up (exit). */
typedef struct {
  vf_ring *ring_out, *ring_co;
  uint8_t *out10;
  size_t out_bytes;
  long done_segs;
  int out_buf_sec;
} out_sinks;

static int ring_sink_write (vf_ring *r, const void *buf, size_t n, const char *what);

/* one finished segment of the main stream: co-add ring at once (:1416-1422); heimdall ring through the
 * 10-s buffer (:1370-1376, :1482-1494: the whole buffer after the 10th second, then each new second,
 * which the reference puts back at the start of the buffer) */
static int sinks_segment (out_sinks *k, const uint8_t *seg)
{
  if (k->ring_co && ring_sink_write (k->ring_co, seg, k->out_bytes, "coadd")) return -1;
  if (k->ring_out) {
    const long per_sec = SEG_PER_SEC, warm = (long) k->out_buf_sec * per_sec;
    const long slot = k->done_segs < warm ? k->done_segs : k->done_segs % per_sec;
    memcpy (k->out10 + (size_t) slot * k->out_bytes, seg, k->out_bytes);
  }
  k->done_segs++;
  if (k->ring_out && k->done_segs % SEG_PER_SEC == 0) {
    const long sec = k->done_segs / SEG_PER_SEC;
    if (sec >= k->out_buf_sec) {
      const size_t n = (size_t) (sec == k->out_buf_sec ? k->out_buf_sec : 1) * SEG_PER_SEC * k->out_bytes;
      if (ring_sink_write (k->ring_out, k->out10, n, "outgoing")) return -1;
    }
  }
  return 0;
}

static int ring_sink_write (vf_ring *r, const void *buf, size_t n, const char *what)
{
  if (vf_ring_get_nfull (r) == vf_ring_get_nbufs (r) - 1) {
    fprintf (stderr, "failed buffer check\n");
    logmsg ("ERR", "Only one free buffer left!  Aborting output.\n");
    return -1;
  }
  const ssize_t w = vf_ring_write (r, buf, n);
  if (w != (ssize_t) n) {
    fprintf (stderr, "failed ipcio write\n");
    logmsg ("ERR", "Tried to write %zu bytes to %s psrdada buffer but only wrote %zd.", n, what, w);
    return -1;
  }
  return 0;
}

int main (int argc, char **argv)
{
  const char *file = NULL, *datadir = ".", *logfile = NULL;
  int synth_n = 0, loops = 1, station = 1, nbuf = 0, write_fb = 1, single = 0, json = 0, rfi_flag = 0;
  int keep_bandpass = 0, profile = 0;
  long shm_key = -1, key_out = 0, key_co = 0;                       /* :344-345 */
  unsigned long long seed = 102;
  vf_config cfg;
  vf_config_default (&cfg);
  int c;
  char mc_group[64] = "";
  int mc_port = VF_MC_READER_PORT, mc_sock = -1;
  while ((c = getopt (argc, argv, "hf:S:L:e:Fa:n:D:w:b:P:r:isg:ol:jk:K:C:p:m:RWHT")) != -1) {
    switch (c) {
      case 'h': usage (); return 0;
      case 'f': file = optarg; break;
      case 'S': synth_n = atoi (optarg); break;
      case 'L': loops = atoi (optarg); break;
      case 'e': seed = strtoull (optarg, NULL, 10); break;
      case 'F': rfi_flag = 1; break;
      case 'a': station = atoi (optarg); break;
      case 'n': nbuf = atoi (optarg); break;
      case 'D': datadir = optarg; break;
      case 'w': write_fb = atoi (optarg); break;
      case 'b': cfg.nbit = atoi (optarg);                         /* :415-426 */
        if (!(cfg.nbit == 2 || cfg.nbit == 4 || cfg.nbit == 8)) { fprintf (stderr, "Unsupported NBIT!\n"); return 1; }
        break;
      case 'P': cfg.npol = atoi (optarg);                         /* :409-413 */
        if (!(cfg.npol == 1 || cfg.npol == 2)) { fprintf (stderr, "Unsupported NPOL!\n"); return 1; }
        break;
      case 'r': cfg.rfi_mode = atoi (optarg);                     /* :427-437 */
        if (cfg.rfi_mode < 0 || cfg.rfi_mode > 2) { fprintf (stderr, "Unsupported RFI mode!\n"); return 1; }
        break;
      case 'i': cfg.inject_frb = 1; break;
      case 'R': keep_bandpass = 1; break;
      case 'W': cfg.keep_stats = 1; break;                          /* WRITE_KURTO, src/process_baseband.h:48 */
      case 'H': cfg.do_histo = 1; break;                            /* DOHISTO, :47 */
      case 'T': profile = 1; break;
      case 's': single = 1; break;
      case 'g': cfg.gpu_id = atoi (optarg); break;
      case 'o': g_stdout = 1; break;
      case 'l': logfile = optarg; break;
      case 'j': json = 1; break;
      case 'm': {
        strncpy (mc_group, optarg, sizeof (mc_group) - 1);
        char *colon = strchr (mc_group, ':');
        if (colon) { *colon = 0; mc_port = atoi (colon + 1); }
        break;
      }
      case 'k': shm_key = (long) strtoul (optarg, NULL, 16); break;  /* input ring key, :541-569 */
      case 'K': key_out = (long) strtoul (optarg, NULL, 16); break;  /* heimdall ring, :377-383 */
      case 'C': key_co = (long) strtoul (optarg, NULL, 16); break;   /* co-add ring, :370-376 */
      case 'p': break;                                            /* port of the reference: accepted, unused */
      default: usage (); return 1;
    }
  }
  if (!file && synth_n <= 0 && shm_key < 0) { usage (); return 1; }
  if (logfile) g_log = fopen (logfile, "a");
  signal (SIGINT, on_signal);
  signal (SIGTERM, on_signal);
  if (mc_group[0]) {                                              /* :764-768 */
    mc_sock = vf_mc_open (mc_group, mc_port);
    if (mc_sock < 0) logmsg ("ERR", "cannot open control socket %s:%d\n", mc_group, mc_port);
    else logmsg ("INFO", "listening for commands on %s:%d\n", mc_group, mc_port);
  }

  vf_handle *h = NULL;
  int rc = vf_create (&cfg, &h);
  if (rc) {
    logmsg ("ERR", "vf_create: %s (%s)\n", vf_strerror (rc), h ? vf_last_error (h) : "");
    return 20;                                                    /* cudacheck's code, src/cuda_util.cu:10 */
  }
  const size_t out_bytes = vf_segment_out_bytes (h);

  void *ring_mem = NULL;
  vf_ring *ring = NULL;
  int ring_registered = 0;
  if (shm_key >= 0) {
    /* dada_hdu_connect + lock_read (:541-569); the blocks are page-locked in place so that segments
     * are DMA'd straight out of the ring (psrdada: dada_cuda_dbregister) */
    ring = vf_ring_connect_shm ((int) shm_key);
    if (!ring) { logmsg ("ERR", "cannot connect to the ring with key %lx\n", shm_key); return 1; }
    if (vf_ring_get_bufsz (ring) != SEC_BYTES) { logmsg ("ERR", "ring blocks must be one second (%zu bytes)\n", SEC_BYTES); return 1; }
    nbuf = (int) vf_ring_get_nbufs (ring);
    if (vf_host_register (vf_ring_data_base (ring), (size_t) nbuf * SEC_BYTES) == VF_OK) ring_registered = 1;
    else logmsg ("INFO", "could not page-lock the ring; copies will be staged by the driver\n");
  } else {
    if (nbuf <= 0) nbuf = synth_n > 0 ? synth_n : 4;
    if (synth_n > 0 && nbuf != synth_n) { logmsg ("ERR", "-n must equal -S for the in-place synthetic replay\n"); return 1; }
    if (vf_host_alloc (&ring_mem, (size_t) nbuf * SEC_BYTES)) { logmsg ("ERR", "cannot pin %d ring blocks\n", nbuf); return 21; }
    ring = vf_ring_create ((uint64_t) nbuf, SEC_BYTES, ring_mem);
  }
  vf_ring *ring_out = NULL, *ring_co = NULL;
  if (key_out && !(ring_out = vf_ring_connect_shm ((int) key_out))) {                 /* :550-558 */
    logmsg ("ERR", "Unable to connect to outgoing PSRDADA buffer key=%lx!\n", key_out); return 1; }
  if (key_co && !(ring_co = vf_ring_connect_shm ((int) key_co))) {                    /* :560-568 */
    logmsg ("ERR", "Unable to connect to Coadding PSRDADA buffer key=%lx!\n", key_co); return 1; }
  /* 10 s of main-stream output, flushed to the heimdall ring once full and second by second after that (:692-697) */
  const int out_buf_sec = 10;
  uint8_t *out10 = ring_out ? (uint8_t *) malloc ((size_t) out_buf_sec * SEG_PER_SEC * out_bytes) : NULL;
  /* Statistics dumps, histogram and FRB injection are per segment (the library keeps one launch pair per segment
   * for them): the per-segment path.  Otherwise a whole one-second block goes to the library at once: one copy,
   * one depacketiser launch over the second, one launch pair over its ten segments. */
  const int per_segment = cfg.keep_stats || cfg.do_histo || cfg.inject_frb;
  const int seg_per_submit = per_segment ? 1 : SEG_PER_SEC;
  if (!per_segment && (rc = vf_reserve_vdif_blocks (h, SEG_PER_SEC))) {          /* start-up, like the reference's allocations (:572-690) */
    logmsg ("ERR", "vf_reserve_vdif_blocks: %s (%s)\n", vf_strerror (rc), vf_last_error (h));
    return 20;
  }
  uint8_t *obuf[2][2] = {{NULL, NULL}, {NULL, NULL}};             /* [slot][main, raw] pinned */
  for (int s = 0; s < 2; ++s)
    for (int k = 0; k < 2; ++k)
      if (vf_host_alloc ((void **) &obuf[s][k], (size_t) seg_per_submit * out_bytes)) return 21;
  const int T = cfg.ffts_per_seg;
  float *st_kur = NULL, *st_kur_fb = NULL, *st_w = NULL;
  uint32_t *st_histo = NULL;
  if (cfg.keep_stats) {
    st_kur = (float *) malloc (sizeof (float) * 2 * (size_t) T * 25);
    st_kur_fb = (float *) malloc (sizeof (float) * 2 * (size_t) T);
    st_w = (float *) malloc (sizeof (float) * 2 * (size_t) T);
  }
  if (cfg.do_histo) st_histo = (uint32_t *) malloc (sizeof (uint32_t) * 512);

  if (synth_n > 0) {
    /* templates straight into the ring blocks (not timed) */
    vf_gen_params g;
    vf_gen_defaults (&g);
    g.seed = seed;
    if (rfi_flag) { g.rfi_amp = 60; g.rfi_burst_every = 16; }
    uint8_t *scratch = (uint8_t *) malloc (2 * (size_t) VF_VLITE_RATE);
    if (!scratch) return 21;
    for (int s = 0; s < synth_n; ++s)
      vf_gen_vdif_block (&g, station, (uint32_t) s, scratch, (uint8_t *) ring_mem + (size_t) s * SEC_BYTES);
    free (scratch);
    logmsg ("INFO", "generated %d synthetic second(s), seed %llu\n", synth_n, seed);
  }

  feeder_args fa = { ring, file, synth_n, loops, station, 3600u * 5, 0 };
  pthread_t feeder;
  const int have_feeder = shm_key < 0;
  if (have_feeder) pthread_create (&feeder, NULL, file ? feeder_file : feeder_synth, &fa);

  double total_data_s = 0, total_wall_s = 0;
  long total_segments = 0;
  int exit_status = 0;

  while (!g_quit) {                                               /* observations, :784 */
    char inhdr[VF_RING_HEADER_SIZE];
    logmsg ("INFO", "Waiting for DADA header.\n");
    int hr;
    while ((hr = vf_ring_header_read (ring, inhdr, 200)) == 1 && !g_quit) {
      int cmds[5];
      vf_get_cmds (cmds, mc_sock);                                /* :793-795 */
      if (cmds[2]) { logmsg ("INFO", "Received CMD_QUIT.  Exiting.\n"); g_quit = 1; }
    }
    if (hr != 0) break;
    logmsg ("INFO", "psrdada header:\n%s", inhdr);
    logmsg ("INFO", "Beginning new observation.\n");
    vf_obs_info obs;
    vf_obs_info_from_header (inhdr, &obs);

    FILE *fb_main = NULL, *fb_raw = NULL;
    char fbfile[512] = "", fbfile_kur[512] = "";
    const double t_obs = now_s ();
    long seconds_done = 0, seg_counter = 0;
    int first = 1, aborted = 0;
    int pend_slot[2] = {0, 0};
    long pend_sec[2] = {0, 0};
    int blocks_open = 0;
    out_sinks sinks = { ring_out, ring_co, out10, out_bytes, 0, out_buf_sec };
    FILE *kurto_fp = NULL, *block_kurto_fp = NULL, *weight_fp = NULL, *histo_fp = NULL;
    double dev_ms = 0, k1_ms = 0, k2_ms = 0, wait_s = 0, write_s = 0;
    long skipped_frames = 0, missing_frames = 0;
    if (!keep_bandpass)
      vf_reset_bandpass (h, -1);  /* the reference keeps the bandpass across observations (:700-709, flag -R); a new
                                     stream here is by default a new antenna-time and starts from its own mean */

    /* one finished submission (a segment or a block of ten): wait, report skipped frames, write the outputs */
#define FINISH_SLOT(slot) do { \
      const double tw0_ = now_s (); \
      rc = vf_wait (h, (slot)); \
      wait_s += now_s () - tw0_; \
      pend_slot[(slot)] = 0; \
      if (rc == VF_ERR_VDIF) { \
        /* frames of another second or outside the block were skipped, their samples stay zero (dropped data): \
         * log and go on, as the reference goes on over the writer's fill frames (src/writer.c:674-687) */ \
        unsigned int cnt_[5]; \
        if (vf_vdif_report (h, (slot), cnt_) == VF_OK) { \
          skipped_frames += cnt_[0] + cnt_[1]; \
          logmsg ("WARN", "second %ld: %u frame(s) outside the block, %u of another second skipped; %u of %u placed\n", \
                  pend_sec[(slot)], cnt_[0], cnt_[1], cnt_[2], cnt_[4]); \
        } \
        rc = VF_OK; \
      } \
      if (rc) { logmsg ("ERR", "segment failed: %s\n", vf_last_error (h)); aborted = 1; exit_status = 1; break; } \
      { unsigned int cnt_[5]; if (vf_vdif_report (h, (slot), cnt_) == VF_OK && cnt_[2] < cnt_[4]) missing_frames += cnt_[4] - cnt_[2]; } \
      if (profile) { float a_ = 0, b_ = 0, c_ = 0; if (vf_slot_elapsed_ms (h, (slot), &a_, &b_, &c_) == VF_OK) { dev_ms += a_; k1_ms += b_; k2_ms += c_; } } \
      const double tf0_ = now_s (); \
      if (fb_main) fwrite (obuf[(slot)][0], 1, (size_t) seg_per_submit * out_bytes, fb_main);     /* :1438-1441 */ \
      if (fb_raw) fwrite (obuf[(slot)][1], 1, (size_t) seg_per_submit * out_bytes, fb_raw); \
      for (int q_ = 0; q_ < seg_per_submit && !aborted; ++q_) \
        if (sinks_segment (&sinks, obuf[(slot)][0] + (size_t) q_ * out_bytes)) { aborted = 1; exit_status = 1; } \
      if (per_segment && (kurto_fp || histo_fp)) {                                            /* :1378-1393, :1444-1450 */ \
        if (vf_get_stats (h, 0, NULL, st_kur, NULL, NULL, st_kur_fb, NULL, st_w, st_histo) == VF_OK) { \
          if (histo_fp) fwrite (st_histo, sizeof (uint32_t), 512, histo_fp); \
          if (kurto_fp) { \
            fwrite (st_kur, sizeof (float), 2 * (size_t) T * 25, kurto_fp); \
            fwrite (st_kur_fb, sizeof (float), 2 * (size_t) T, block_kurto_fp); \
            fwrite (st_w, sizeof (float), (size_t) T, weight_fp); \
          } \
        } \
      } \
      write_s += now_s () - tf0_; \
    } while (0)

    for (;;) {                                                    /* seconds */
      uint64_t nbytes = 0;
      if (!per_segment && pend_slot[seconds_done & 1]) {
        /* the block of two seconds ago: finish it and hand its ring block back BEFORE asking for the next one
         * (a ring of two blocks would otherwise have none left for the writer) */
        const int slot_ = (int) (seconds_done & 1);
        FINISH_SLOT (slot_);
        vf_ring_block_read_close (ring); blocks_open--;
        if (aborted) break;
      }
      const unsigned char *blk = (const unsigned char *) vf_ring_block_read_open (ring, &nbytes);
      if (!blk) break;                                            /* end of data: primary exit, :1044-1051 */
      if (nbytes < SEC_BYTES) {
        logmsg ("INFO", "Incomplete final second (%lu bytes), dropped.\n", (unsigned long) nbytes);
        blocks_open++;
        break;
      }
      const vf_vdif_header *vh = (const vf_vdif_header *) blk;
      if (first) {
        if (vf_vdif_frame_number (vh) != 0 || vf_vdif_thread_id (vh) != 0) {     /* :843-849 */
          logmsg ("ERR", "Incoming data were not aligned!\n");
          exit_status = 1; aborted = 1;
          blocks_open++;
          break;
        }
        if (write_fb) {                                           /* :852-998 */
          vf_fb_filename (fbfile, sizeof (fbfile), datadir, vh, obs.station_id, 0);
          vf_fb_filename (fbfile_kur, sizeof (fbfile_kur), datadir, vh, obs.station_id, 1);
          if (cfg.rfi_mode) {
            fb_main = fopen (fbfile_kur, "wb");
            if (cfg.rfi_mode == 2) fb_raw = fopen (fbfile, "wb");
          } else
            fb_main = fopen (fbfile, "wb");
          if (!fb_main || (cfg.rfi_mode == 2 && !fb_raw)) { logmsg ("ERR", "cannot open output file in %s\n", datadir); exit_status = 1; aborted = 1; blocks_open++; break; }
          vf_write_sigproc_header (fb_main, &obs, vh, cfg.nbit, cfg.npol);
          if (fb_raw) vf_write_sigproc_header (fb_raw, &obs, vh, cfg.nbit, cfg.npol);
        }
        {
          char dh[VF_RING_HEADER_SIZE];
          if (!write_fb) {
            vf_fb_filename (fbfile, sizeof (fbfile), datadir, vh, obs.station_id, 0);
            vf_fb_filename (fbfile_kur, sizeof (fbfile_kur), datadir, vh, obs.station_id, 1);
          }
          vf_write_psrdada_header (dh, &obs, vh, cfg.nbit, cfg.npol, cfg.rfi_mode ? fbfile_kur : fbfile);
          logmsg ("INFO", "output header:\n%s", dh);
          if (ring_out) vf_ring_header_write (ring_out, dh);                   /* :981-984 */
          if (ring_co) vf_ring_header_write (ring_co, dh);                     /* :986-989 */
        }
        if (cfg.keep_stats || cfg.do_histo) {
          /* the reference's dump files: the .fil name with another extension (:861-870) */
          char base[512], name[540];
          snprintf (base, sizeof (base), "%s", fbfile);
          char *dot = strrchr (base, '.');
          if (dot) *dot = 0;
          if (cfg.do_histo) { snprintf (name, sizeof (name), "%s.histo", base); histo_fp = fopen (name, "wb"); logmsg ("INFO", "Writing histograms to %s.\n", name); }
          if (cfg.keep_stats && cfg.rfi_mode) {
            snprintf (name, sizeof (name), "%s.kurto", base); kurto_fp = fopen (name, "wb"); logmsg ("INFO", "Writing kurtosis to %s.\n", name);
            snprintf (name, sizeof (name), "%s.block_kurto", base); block_kurto_fp = fopen (name, "wb"); logmsg ("INFO", "Writing block kurtosis to %s.\n", name);
            snprintf (name, sizeof (name), "%s.weights", base); weight_fp = fopen (name, "wb"); logmsg ("INFO", "Writing weights to %s.\n", name);
            if (!kurto_fp || !block_kurto_fp || !weight_fp) { logmsg ("ERR", "cannot open the statistics files\n"); exit_status = 1; aborted = 1; blocks_open++; break; }
          }
        }
        logmsg ("INFO", "Starting sec=%d, thread=%d\n", vf_vdif_frame_second (vh), vf_vdif_thread_id (vh));
        first = 0;
      }
      const int current_sec = vf_vdif_frame_second (vh);

      if (!per_segment) {
        /* ---- the whole second at once.  Frames may sit anywhere in the block: the depacketiser places them
         * by thread id and frame number across the second and skips frames of another second (:1015-1035) */
        const int slot = (int) (seconds_done & 1);
        rc = vf_submit_vdif_block_async (h, slot, 0, blk, FRAMES_PER_SEC_2POL, 0, (long) current_sec, SEG_PER_SEC,
                                         obuf[slot][0], obuf[slot][1]);
        if (rc) { logmsg ("ERR", "submit failed: %s\n", vf_last_error (h)); aborted = 1; exit_status = 1; blocks_open++; break; }
        pend_slot[slot] = 1; pend_sec[slot] = seconds_done;
        seg_counter += SEG_PER_SEC;
        blocks_open++;
      } else {
        /* ---- segment by segment (statistics dumps / histogram / FRB injection) */
        const int inject_now = cfg.inject_frb && (current_sec % 60 == 0);        /* :1098-1101 */
        if (inject_now) logmsg ("INFO", "Injecting an FRB!!!\n");
        for (int iseg = 0; iseg < SEG_PER_SEC && !aborted; ++iseg, ++seg_counter) {  /* :1108 */
          const int slot = (int) (seg_counter & 1);
          if (pend_slot[slot]) { FINISH_SLOT (slot); if (aborted) break; }
          if (cfg.inject_frb)
            vf_set_frb_injection (h, inject_now ? iseg * 1024 : -1, 80.f, (float) (2e-3 * 10 * 1024), 1.05f);   /* :1238-1240 */
          /* the window of this segment within the block; a frame the writer put elsewhere in the second is reported */
          rc = vf_submit_vdif_block_async (h, slot, 0, blk + (size_t) iseg * (FRAMES_PER_SEC_2POL / SEG_PER_SEC) * VF_VD_FRM,
                                           FRAMES_PER_SEC_2POL / SEG_PER_SEC, (uint32_t) (iseg * (VF_FRAME_RATE / SEG_PER_SEC)),
                                           (long) current_sec, 1, obuf[slot][0], obuf[slot][1]);
          if (rc) { logmsg ("ERR", "submit failed: %s\n", vf_last_error (h)); aborted = 1; exit_status = 1; break; }
          pend_slot[slot] = 1; pend_sec[slot] = seconds_done;
          if (kurto_fp || histo_fp) { FINISH_SLOT (slot); if (aborted) break; }   /* the dumps read the statistics of this segment */
        }
        /* two segments of this block may still be in flight: it stays open, and the block before it
         * (whose segments have all been waited for by now) goes back to the writer */
        if (blocks_open == 1) { vf_ring_block_read_close (ring); blocks_open = 0; }
        blocks_open++;
      }
      if (aborted) break;
      seconds_done++;
      if (seconds_done % 10 == 0) {                               /* RT_PROFILE watchdog, :1461-1477 */
        const double wall = now_s () - t_obs;
        if (wall - seconds_done > 0.5)
          logmsg ("ERR", "Falling behind real time: %.2f s wall for %ld s of data.\n", wall, seconds_done);
      }
      if (vf_test_for_cmd (VF_CMD_QUIT, mc_sock)) {                 /* once per second, :1081-1092 */
        logmsg ("INFO", "Received CMD_QUIT, indicating data taking is ceasing.  Exiting.\n");
        g_quit = 1;
      }
      if (g_quit) break;
    }
    /* end of the observation: drain what is in flight (oldest first), release the blocks */
    for (int k = 0; k < 2; ++k) {
      const int slot = per_segment ? (int) ((seg_counter + k) & 1) : (int) ((seconds_done + k) & 1);
      if (!pend_slot[slot]) continue;
      do { FINISH_SLOT (slot); } while (0);
    }
#undef FINISH_SLOT
    if (kurto_fp) fclose (kurto_fp);
    if (block_kurto_fp) fclose (block_kurto_fp);
    if (weight_fp) fclose (weight_fp);
    if (histo_fp) fclose (histo_fp);
    if (skipped_frames || missing_frames)
      logmsg ("WARN", "observation: %ld frame(s) skipped, %ld frame(s) absent (their samples are zeros)\n", skipped_frames, missing_frames);
    if (profile) {
      /* the reference's per-stage table (:1538-1558) for the stages that still exist: the chain is two kernels */
      logmsg ("INFO", "Device time.%.3f\n", dev_ms * 1e-3);
      logmsg ("INFO", "Channelise..%.3f   (unpack, statistics, excision, FFT, detection: Convert + Kurtosis + FFT of the reference)\n", k1_ms * 1e-3);
      logmsg ("INFO", "Normalise...%.3f   (Normalize + Pscrunch + Tscrunch + Digitize of the reference; includes the wait for bandpass order)\n", k2_ms * 1e-3);
      logmsg ("INFO", "Wait........%.3f   (host blocked on the device or the copies)\n", wait_s);
      logmsg ("INFO", "Write.......%.3f\n", write_s);
    }
    if (!first) {                                                 /* dada_hdu_unlock_write, :1498-1511 */
      if (ring_out) vf_ring_end_of_data (ring_out);
      if (ring_co) vf_ring_end_of_data (ring_co);
    }
    while (blocks_open > 0) { vf_ring_block_read_close (ring); blocks_open--; }
    if (!aborted) { uint64_t nb; while (vf_ring_block_read_open (ring, &nb)) vf_ring_block_read_close (ring); }   /* reach EOD */
    if (fb_main) fclose (fb_main);
    if (fb_raw) fclose (fb_raw);
    const double wall = now_s () - t_obs;
    logmsg ("INFO", "Proc Time...%.3f s for %ld s of data (%.1fx real time)\n", wall, seconds_done,
            wall > 0 ? seconds_done / wall : 0.0);                /* :1534-1536 */
    total_data_s += seconds_done; total_wall_s += wall; total_segments += seg_counter;
    if (aborted) { if (have_feeder) vf_ring_shutdown (ring); break; }
    if (single || synth_n > 0 || file) break;
  }
  if (have_feeder) {
    vf_ring_shutdown (ring);
    pthread_join (feeder, NULL);
  }
  if (json)
    printf ("{\"program\": \"process_baseband\", \"seconds\": %.0f, \"segments\": %ld, \"wall_s\": %.6f, \"x_realtime\": %.3f, "
            "\"nbit\": %d, \"npol\": %d, \"rfi_mode\": %d, \"bytes_in\": %.0f, \"exit\": %d}\n",
            total_data_s, total_segments, total_wall_s, total_wall_s > 0 ? total_data_s / total_wall_s : 0.0,
            cfg.nbit, cfg.npol, cfg.rfi_mode, total_data_s * (double) SEC_BYTES, exit_status | fa.rc);
  vf_mc_close (mc_sock);
  if (ring_out) vf_ring_destroy (ring_out);
  if (ring_co) vf_ring_destroy (ring_co);
  free (out10);
  if (ring_registered) vf_host_unregister (vf_ring_data_base (ring));
  vf_ring_destroy (ring);
  for (int s = 0; s < 2; ++s) for (int k = 0; k < 2; ++k) vf_host_free (obuf[s][k]);
  if (ring_mem) vf_host_free (ring_mem);
  vf_destroy (h);
  if (g_log) fclose (g_log);
  return exit_status | fa.rc;
}
