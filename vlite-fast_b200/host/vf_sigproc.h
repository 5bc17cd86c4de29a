/*
 * SIGPROC filterbank header and psrdada ASCII header of the output stream,
 * byte-compatible with the reference: send_* (src/util.c:51-81),
 * write_sigproc_header (src/process_baseband.cu:226-270), the key set of
 * write_psrdada_header (:136-201) and the file naming of get_fbfile (:288-304).
 */
#ifndef VF_SIGPROC_H
#define VF_SIGPROC_H
#include <stddef.h>
#include <stdio.h>
#include "vf_vdif.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  char name[128];       /* NAME       */
  int station_id;       /* STATIONID  */
  double ra, dec;       /* RA, DEC (radians, as the VLA executor gives them) */
  double scanstart;     /* SCANSTART  */
} vf_obs_info;

void vf_send_string (const char *s, FILE *fp);
void vf_send_int (const char *name, int v, FILE *fp);
void vf_send_double (const char *name, double v, FILE *fp);
/* returns the number of header bytes written */
long vf_write_sigproc_header (FILE *fp, const vf_obs_info *obs, const vf_vdif_header *first, int nbit, int npol);
/* header into buf (for tests); returns its length, or -1 if cap is too small */
long vf_sigproc_header_to_buffer (char *buf, size_t cap, const vf_obs_info *obs, const vf_vdif_header *first, int nbit, int npol);
/* 4096-byte ASCII header for the downstream ring (heimdall) */
int vf_write_psrdada_header (char *hdr4096, const vf_obs_info *obs, const vf_vdif_header *first, int nbit, int npol, const char *fb_file);
/* DATADIR/%Y%m%d_%H%M%S[_muos]_ea%02d[_kur].fil */
void vf_fb_filename (char *out, size_t cap, const char *datadir, const vf_vdif_header *first, int station_id, int kur);

/* psrdada ascii_header_get / ascii_header_set on a "KEY value\n" block */
int vf_ascii_header_set (char *hdr, size_t cap, const char *key, const char *fmt, ...);
int vf_ascii_header_get (const char *hdr, const char *key, const char *fmt, ...);
void vf_obs_info_from_header (const char *hdr, vf_obs_info *obs);

#ifdef __cplusplus
}
#endif
#endif
