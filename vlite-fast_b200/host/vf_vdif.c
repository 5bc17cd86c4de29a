#include <string.h>
#include "vf_vdif.h"

int vf_vdif_thread_id (const vf_vdif_header *h) { return (int) ((h->w[3] >> 16) & 0x3FFu); }
int vf_vdif_frame_number (const vf_vdif_header *h) { return (int) (h->w[1] & 0xFFFFFFu); }
uint32_t vf_vdif_epoch_sec_offset (const vf_vdif_header *h) { return h->w[0] & 0x3FFFFFFFu; }
int vf_vdif_frame_second (const vf_vdif_header *h) { return (int) (vf_vdif_epoch_sec_offset (h) % 86400u); }
int vf_vdif_epoch (const vf_vdif_header *h) { return (int) ((h->w[1] >> 24) & 0x3Fu); }
int vf_vdif_station_id (const vf_vdif_header *h) { return (int) (h->w[3] & 0xFFFFu); }
int vf_vdif_frame_bytes (const vf_vdif_header *h) { return (int) ((h->w[2] & 0xFFFFFFu) * 8u); }

/* days from civil date (proleptic Gregorian) to MJD */
static int ymd_to_mjd (int y, int m, int d)
{
  int a = (14 - m) / 12, yy = y + 4800 - a, mm = m + 12 * a - 3;
  int jdn = d + (153 * mm + 2) / 5 + 365 * yy + yy / 4 - yy / 100 + yy / 400 - 32045;
  return jdn - 2400001;        /* JDN at noon - 2400000.5 */
}

/* reference epoch e: 1 Jan (e even) or 1 Jul (e odd) of year 2000 + e/2 */
static int epoch_mjd (int epoch) { return ymd_to_mjd (2000 + epoch / 2, (epoch % 2) ? 7 : 1, 1); }

int vf_vdif_frame_mjd (const vf_vdif_header *h)
{
  return epoch_mjd (vf_vdif_epoch (h)) + (int) (vf_vdif_epoch_sec_offset (h) / 86400u);
}

int vf_vdif_frame_mjd_sec (const vf_vdif_header *h) { return (int) (vf_vdif_epoch_sec_offset (h) % 86400u); }

double vf_vdif_frame_dmjd (const vf_vdif_header *h, int frames_per_sec)
{
  return vf_vdif_frame_mjd (h) + (vf_vdif_frame_mjd_sec (h) + (double) vf_vdif_frame_number (h) / frames_per_sec) / 86400.0;
}

time_t vf_vdif_to_unixepoch (const vf_vdif_header *h)
{
  /* MJD 40587 = 1970-01-01 */
  return (time_t) (epoch_mjd (vf_vdif_epoch (h)) - 40587) * 86400 + (time_t) vf_vdif_epoch_sec_offset (h);
}

void vf_vdif_set (vf_vdif_header *h, uint32_t seconds, uint32_t frame, int epoch, int station, int thread)
{
  memset (h, 0, sizeof (*h));
  h->w[0] = seconds & 0x3FFFFFFFu;
  h->w[1] = (frame & 0xFFFFFFu) | (((uint32_t) epoch & 0x3Fu) << 24);
  h->w[2] = (VF_VD_FRM / 8) & 0xFFFFFFu;
  h->w[3] = ((uint32_t) station & 0xFFFFu) | (((uint32_t) thread & 0x3FFu) << 16) | (7u << 26);
}
