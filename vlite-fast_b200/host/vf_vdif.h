/*
 * VDIF frame header access for the VLITE stream (32-byte header, 5000 8-bit
 * samples, one channel; thread id != 0 is the second polarisation).  The
 * reference takes these from the external vdifio library (not in its tree,
 * src/INSTALL:6); the accessors below restate the ones it calls
 * (src/process_baseband.cu:191-194,241,843-844,1001-1002,1017-1019;
 * src/utils.c:498-514) from the published VDIF 1.1 header layout, which
 * analysis/baseband.py:19-28 of the reference restates too:
 *   word0 bits 0-29 seconds from the reference epoch, bit 31 invalid
 *   word1 bits 0-23 frame number within the second, bits 24-29 reference epoch
 *   word2 bits 0-23 frame length in units of 8 bytes
 *   word3 bits 0-15 station id, 16-25 thread id, 26-30 bits per sample - 1
 */
#ifndef VF_VDIF_H
#define VF_VDIF_H
#include <stdint.h>
#include <time.h>
#ifdef __cplusplus
extern "C" {
#endif

#define VF_VD_FRM 5032          /* src/process_baseband.h:16 */
#define VF_VD_DAT 5000          /* :17 */
#define VF_VLITE_RATE 128000000 /* :18 */
#define VF_FRAME_RATE 25600     /* :19 */

typedef struct { uint32_t w[8]; } vf_vdif_header;

int vf_vdif_thread_id (const vf_vdif_header *h);        /* getVDIFThreadID          */
int vf_vdif_frame_number (const vf_vdif_header *h);     /* getVDIFFrameNumber       */
int vf_vdif_frame_second (const vf_vdif_header *h);     /* getVDIFFrameSecond: seconds % 86400 */
int vf_vdif_epoch (const vf_vdif_header *h);            /* getVDIFEpoch             */
uint32_t vf_vdif_epoch_sec_offset (const vf_vdif_header *h);  /* getVDIFFrameEpochSecOffset */
int vf_vdif_station_id (const vf_vdif_header *h);
int vf_vdif_frame_bytes (const vf_vdif_header *h);
int vf_vdif_frame_mjd (const vf_vdif_header *h);        /* getVDIFFrameMJD          */
int vf_vdif_frame_mjd_sec (const vf_vdif_header *h);    /* getVDIFFrameMJDSec       */
double vf_vdif_frame_dmjd (const vf_vdif_header *h, int frames_per_sec);   /* getVDIFFrameDMJD */
/* src/utils.c:498-514 (without the mktime time-zone round trip) */
time_t vf_vdif_to_unixepoch (const vf_vdif_header *h);
void vf_vdif_set (vf_vdif_header *h, uint32_t seconds, uint32_t frame, int epoch, int station, int thread);

#ifdef __cplusplus
}
#endif
#endif
