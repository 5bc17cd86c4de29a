#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "vf_sigproc.h"

#define NFFT 12500
#define NCHAN (NFFT / 2 + 1)
#define NSCRUNCH 8
#define CHANMIN 2155
#define CHANMAX 6250

void vf_send_string (const char *s, FILE *fp)
{
  int len = (int) strlen (s);
  fwrite (&len, sizeof (int), 1, fp);
  fwrite (s, 1, (size_t) len, fp);
}

void vf_send_int (const char *name, int v, FILE *fp)
{
  vf_send_string (name, fp);
  fwrite (&v, sizeof (int), 1, fp);
}

void vf_send_double (const char *name, double v, FILE *fp)
{
  vf_send_string (name, fp);
  fwrite (&v, sizeof (double), 1, fp);
}

/* field order and arithmetic of src/process_baseband.cu:239-268 (the float
 * intermediates of the coordinate conversion included) */
long vf_write_sigproc_header (FILE *fp, const vf_obs_info *obs, const vf_vdif_header *first, int nbit, int npol)
{
  long at = ftell (fp);
  double chbw = -64. / NCHAN;
  double tsamp = (double) NFFT / VF_VLITE_RATE * NSCRUNCH;
  double mjd = vf_vdif_frame_dmjd (first, VF_FRAME_RATE);
  vf_send_string ("HEADER_START", fp);
  vf_send_string ("source_name", fp);
  vf_send_string (obs->name, fp);
  vf_send_int ("barycentric", 0, fp);
  vf_send_int ("telescope_id", obs->station_id, fp);
  float hh = (float) ((180 / M_PI) * (24. / 360) * obs->ra);
  float mm = (hh - (int) hh) * 60;
  float ss = (mm - (int) mm) * 60;
  float sigproc_ra = (float) ((int) hh * 1e4 + (int) mm * 1e2 + ss);
  vf_send_double ("src_raj", sigproc_ra, fp);
  float dd = (float) ((180 / M_PI) * fabs (obs->dec));
  mm = (dd - (int) dd) * 60;
  ss = (mm - (int) mm) * 60;
  float sigproc_dec = (float) ((int) dd * 1e4 + (int) mm * 1e2 + ss);
  vf_send_double ("src_dej", sigproc_dec, fp);
  vf_send_int ("data_type", 1, fp);
  vf_send_double ("fch1", 384 + (CHANMIN - 0.5) * chbw, fp);
  vf_send_double ("foff", chbw, fp);
  vf_send_int ("nchans", CHANMAX - CHANMIN + 1, fp);
  vf_send_int ("nbits", nbit, fp);
  vf_send_double ("tstart", mjd, fp);
  vf_send_double ("tsamp", tsamp, fp);
  vf_send_int ("nifs", npol, fp);
  vf_send_string ("HEADER_END", fp);
  return ftell (fp) - at;
}

long vf_sigproc_header_to_buffer (char *buf, size_t cap, const vf_obs_info *obs, const vf_vdif_header *first, int nbit, int npol)
{
  FILE *fp = tmpfile ();
  if (!fp) return -1;
  long n = vf_write_sigproc_header (fp, obs, first, nbit, npol);
  if (n < 0 || (size_t) n > cap) { fclose (fp); return -1; }
  rewind (fp);
  size_t got = fread (buf, 1, (size_t) n, fp);
  fclose (fp);
  return got == (size_t) n ? n : -1;
}

int vf_ascii_header_set (char *hdr, size_t cap, const char *key, const char *fmt, ...)
{
  char val[256], line[512];
  va_list ap;
  va_start (ap, fmt);
  vsnprintf (val, sizeof (val), fmt, ap);
  va_end (ap);
  snprintf (line, sizeof (line), "%-12s %s\n", key, val);
  /* replace an existing key */
  size_t klen = strlen (key);
  char *p = hdr;
  while (*p) {
    char *eol = strchr (p, '\n');
    size_t l = eol ? (size_t) (eol - p) + 1 : strlen (p);
    if (!strncmp (p, key, klen) && (p[klen] == ' ' || p[klen] == '\t')) {
      size_t tail = strlen (p + l);
      if (strlen (hdr) - l + strlen (line) + 1 > cap) return -1;
      memmove (p + strlen (line), p + l, tail + 1);
      memcpy (p, line, strlen (line));
      return 0;
    }
    p += l;
  }
  if (strlen (hdr) + strlen (line) + 1 > cap) return -1;
  strcat (hdr, line);
  return 0;
}

int vf_ascii_header_get (const char *hdr, const char *key, const char *fmt, ...)
{
  size_t klen = strlen (key);
  const char *p = hdr;
  while (p && *p) {
    if (!strncmp (p, key, klen) && (p[klen] == ' ' || p[klen] == '\t')) {
      const char *v = p + klen;
      while (*v == ' ' || *v == '\t') ++v;
      va_list ap;
      va_start (ap, fmt);
      int n = vsscanf (v, fmt, ap);
      va_end (ap);
      return n;
    }
    p = strchr (p, '\n');
    if (p) ++p;
  }
  return -1;
}

/* src/process_baseband.cu:141-150: defaults when a key is absent */
void vf_obs_info_from_header (const char *hdr, vf_obs_info *obs)
{
  memset (obs, 0, sizeof (*obs));
  vf_ascii_header_get (hdr, "STATIONID", "%d", &obs->station_id);
  vf_ascii_header_get (hdr, "RA", "%lf", &obs->ra);
  vf_ascii_header_get (hdr, "DEC", "%lf", &obs->dec);
  if (vf_ascii_header_get (hdr, "NAME", "%127s", obs->name) < 1) strcpy (obs->name, "unknown");
  vf_ascii_header_get (hdr, "SCANSTART", "%lf", &obs->scanstart);
}

int vf_write_psrdada_header (char *hdr, const vf_obs_info *obs, const vf_vdif_header *first, int nbit, int npol, const char *fb_file)
{
  const size_t cap = 4096;
  memset (hdr, 0, cap);
  time_t epoch_seconds = vf_vdif_to_unixepoch (first);
  struct tm utc;
  gmtime_r (&epoch_seconds, &utc);
  char dada_utc[64];
  strftime (dada_utc, sizeof (dada_utc), "%Y-%m-%d-%H:%M:%S", &utc);      /* DADA_TIMESTR */
  double chbw = -64. / NCHAN;
  double tsamp = (double) NFFT / VF_VLITE_RATE * NSCRUNCH * 1e6;           /* microseconds */
  int nchan = CHANMAX - CHANMIN + 1;
  double bw = nchan * chbw;
  double freq = 384. + 0.5 * (CHANMIN + CHANMAX - 1) * chbw;
  int rc = 0;
  rc |= vf_ascii_header_set (hdr, cap, "STATIONID", "%d", obs->station_id);
  rc |= vf_ascii_header_set (hdr, cap, "BEAM", "%d", obs->station_id);
  rc |= vf_ascii_header_set (hdr, cap, "RA", "%lf", obs->ra);
  rc |= vf_ascii_header_set (hdr, cap, "DEC", "%lf", obs->dec);
  rc |= vf_ascii_header_set (hdr, cap, "NAME", "%s", obs->name);
  rc |= vf_ascii_header_set (hdr, cap, "SCANSTART", "%lf", obs->scanstart);
  rc |= vf_ascii_header_set (hdr, cap, "NCHAN", "%d", nchan);
  rc |= vf_ascii_header_set (hdr, cap, "BANDWIDTH", "%lf", bw);
  rc |= vf_ascii_header_set (hdr, cap, "CFREQ", "%lf", freq);
  rc |= vf_ascii_header_set (hdr, cap, "NPOL", "%d", npol);
  rc |= vf_ascii_header_set (hdr, cap, "NBIT", "%d", nbit);
  rc |= vf_ascii_header_set (hdr, cap, "TSAMP", "%lf", tsamp);
  rc |= vf_ascii_header_set (hdr, cap, "UTC_START", "%s", dada_utc);
  rc |= vf_ascii_header_set (hdr, cap, "UNIXEPOCH", "%lf", (double) epoch_seconds);
  rc |= vf_ascii_header_set (hdr, cap, "VDIF_MJD", "%d", vf_vdif_frame_mjd (first));
  rc |= vf_ascii_header_set (hdr, cap, "VDIF_SEC", "%lu", (unsigned long) vf_vdif_frame_mjd_sec (first));
  if (fb_file) rc |= vf_ascii_header_set (hdr, cap, "SIGPROC_FILE", "%s", fb_file);
  return rc;
}

void vf_fb_filename (char *out, size_t cap, const char *datadir, const vf_vdif_header *first, int station_id, int kur)
{
  char ts[64];
  time_t epoch_seconds = vf_vdif_to_unixepoch (first);
  struct tm utc;
  gmtime_r (&epoch_seconds, &utc);
  strftime (ts, sizeof (ts), "%Y%m%d_%H%M%S", &utc);
  /* CHANMIN < 2411 -> "_muos" (src/process_baseband.cu:299-302) */
  snprintf (out, cap, "%s/%s%s_ea%02d%s.fil", datadir, ts, CHANMIN < 2411 ? "_muos" : "", station_id, kur ? "_kur" : "");
}
