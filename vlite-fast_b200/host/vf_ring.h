/*
 * psrdada-style ring buffer shim (in-process, or between processes through SysV shared memory).  The reference reads VDIF frames
 * from a psrdada shared-memory ring fed by writer / genbase / readbase and
 * writes filterbank data to two more (src/process_baseband.cu:541-569,
 * 799-849, 1038, 1416-1422, 1482-1494); psrdada is not in this image, so this
 * file provides the subset of its surface that process_baseband uses:
 *
 *   dada_hdu_create/connect/lock_*      -> vf_ring_create* / vf_ring_destroy
 *   ipcbuf_get_next_write/mark_filled   -> vf_ring_header_write (header block)
 *                                          vf_ring_block_write_open/close (data)
 *   ipcbuf_get_next_read/mark_cleared   -> vf_ring_header_read / vf_ring_block_read_open/close
 *   ipcio_read / ipcio_write            -> vf_ring_read / vf_ring_write (byte stream over the blocks)
 *   ipcbuf_get_nbufs / ipcbuf_get_nfull -> vf_ring_get_nbufs / vf_ring_get_nfull
 *   end of data (EOD)                   -> vf_ring_end_of_data
 *
 * One writer thread and one reader thread.  The data area may be caller
 * provided (pinned memory from vf_host_alloc) so that blocks can be DMA'd to
 * the GPU without a host copy: the reference's geometry is 1-second blocks of
 * 257 638 400 bytes (scripts/start_writer:12).
 */
#ifndef VF_RING_H
#define VF_RING_H
#include <stddef.h>
#include <stdint.h>
#include <sys/types.h>
#ifdef __cplusplus
extern "C" {
#endif

#define VF_RING_HEADER_SIZE 4096

typedef struct vf_ring vf_ring;

/* in-process ring; mem == NULL: malloc the data area (nbufs * bufsz bytes); nbufs <= 1024 */
vf_ring *vf_ring_create (uint64_t nbufs, uint64_t bufsz, void *mem);
/* ring in a SysV shared-memory segment, shared between processes the way psrdada's are:
 * create = "dada_db -k key -b bufsz -n nbufs" (fails if the key exists), connect =
 * dada_hdu_connect, remove = "dada_db -d".  The creator's vf_ring_destroy removes the segment. */
vf_ring *vf_ring_create_shm (int key, uint64_t nbufs, uint64_t bufsz);
vf_ring *vf_ring_connect_shm (int key);
int vf_ring_remove_shm (int key);
void vf_ring_disown (vf_ring *r);                  /* the creator's vf_ring_destroy only detaches (dada_db without -d) */
void vf_ring_destroy (vf_ring *r);
void *vf_ring_data_base (const vf_ring *r);        /* first data block (to page-lock the ring for DMA) */
uint64_t vf_ring_get_nbufs (const vf_ring *r);
uint64_t vf_ring_get_bufsz (const vf_ring *r);
uint64_t vf_ring_get_nfull (vf_ring *r);

/* header block: one ASCII header per observation */
int vf_ring_header_write (vf_ring *r, const char *hdr);            /* writer; blocks while the last one is unread */
int vf_ring_header_read (vf_ring *r, char *hdr4096, int timeout_ms); /* reader; 0 ok, 1 timeout, -1 shut down */

/* data blocks, writer */
void *vf_ring_block_write_open (vf_ring *r);                       /* blocks while the ring is full; NULL after shutdown */
int vf_ring_block_write_close (vf_ring *r, uint64_t nbytes);       /* nbytes < bufsz only for the last block */
ssize_t vf_ring_write (vf_ring *r, const void *src, size_t n);     /* ipcio_write */
int vf_ring_end_of_data (vf_ring *r);                              /* flushes a partial block and marks EOD */

/* data blocks, reader */
/* Several blocks may be open at once (each open returns the next one); close releases the OLDEST
 * open block to the writer.  NULL at EOD (then the next header may be read). */
const void *vf_ring_block_read_open (vf_ring *r, uint64_t *nbytes);
int vf_ring_block_read_close (vf_ring *r);
ssize_t vf_ring_read (vf_ring *r, void *dst, size_t n);            /* ipcio_read: short count only at EOD */

void vf_ring_shutdown (vf_ring *r);                                /* wake everybody, further calls fail */

#ifdef __cplusplus
}
#endif
#endif
