/*
 * The vf_ring surface process_baseband uses for -k / -K / -C (vf_ring.h), implemented on psrdada itself: with this
 * file linked in place of vf_ring.c (make PSRDADA=/path/to/psrdada) the executable attaches to rings made by
 * psrdada's own dada_db and fed by the reference's writer / genbase / readbase.  The three call families of the
 * reference map one to one:
 *
 *   connect      dada_hdu_create / set_key / connect, lock_read or lock_write      src/process_baseband.cu:541-569, 799
 *   input        ipcbuf_get_next_read + mark_cleared on the header block (:807-832); one-second data blocks in place
 *                (ipcio_open_block_read / close_block_read) instead of 51 200 ipcio_read calls of one frame (:837, :1038)
 *   outputs      ipcbuf_get_next_write + mark_filled for the header (:172, :199), ipcio_write for the data (:324),
 *                ipcbuf_get_nbufs / get_nfull for check_buffer (:306-320), dada_hdu_unlock_write at end of data (:1501-1513)
 *
 * psrdada is not in this image: `make -C vlite-fast_b200/host psrdada-check` type-checks this file against
 * psrdada/stub/ (headers that carry psrdada's prototypes and nothing else); it has never been linked or run.
 */
#ifdef HAVE_PSRDADA
#include <stdlib.h>
#include <string.h>
#include "dada_hdu.h"
#include "ipcio.h"
#include "ipcbuf.h"
#include "multilog.h"
#include "../vf_ring.h"

struct vf_ring {
  dada_hdu_t *hdu;
  multilog_t *log;
  int locked;              /* 0 none, 1 read, 2 write */
  int blocks_open;
};

static int lock_as (vf_ring *r, int how)
{
  if (r->locked == how) return 0;
  if (r->locked) return -1;
  if ((how == 1 ? dada_hdu_lock_read (r->hdu) : dada_hdu_lock_write (r->hdu)) < 0) return -1;
  r->locked = how;
  return 0;
}

vf_ring *vf_ring_connect_shm (int key)
{
  vf_ring *r = (vf_ring *) calloc (1, sizeof (*r));
  if (!r) return NULL;
  r->log = multilog_open ("process_baseband", 0);
  r->hdu = dada_hdu_create (r->log);
  dada_hdu_set_key (r->hdu, (key_t) key);
  if (dada_hdu_connect (r->hdu) < 0) { dada_hdu_destroy (r->hdu); free (r); return NULL; }
  return r;
}

void vf_ring_destroy (vf_ring *r)
{
  if (!r) return;
  if (r->locked == 1) dada_hdu_unlock_read (r->hdu);
  if (r->locked == 2) dada_hdu_unlock_write (r->hdu);
  dada_hdu_disconnect (r->hdu);
  dada_hdu_destroy (r->hdu);
  free (r);
}

uint64_t vf_ring_get_nbufs (const vf_ring *r) { return ipcbuf_get_nbufs ((ipcbuf_t *) r->hdu->data_block); }
uint64_t vf_ring_get_bufsz (const vf_ring *r) { return ipcbuf_get_bufsz ((ipcbuf_t *) r->hdu->data_block); }
uint64_t vf_ring_get_nfull (vf_ring *r) { return ipcbuf_get_nfull ((ipcbuf_t *) r->hdu->data_block); }
void *vf_ring_data_base (const vf_ring *r) { (void) r; return NULL; }   /* page-locking goes through dada_cuda_dbregister */

int vf_ring_header_read (vf_ring *r, char *hdr4096, int timeout_ms)
{
  (void) timeout_ms;                              /* psrdada blocks on the semaphore */
  if (lock_as (r, 1)) return -1;
  uint64_t n = 0;
  const char *h = ipcbuf_get_next_read (r->hdu->header_block, &n);
  if (!h) return -1;
  if (n > VF_RING_HEADER_SIZE) n = VF_RING_HEADER_SIZE;
  memset (hdr4096, 0, VF_RING_HEADER_SIZE);
  memcpy (hdr4096, h, n);
  return ipcbuf_mark_cleared (r->hdu->header_block) < 0 ? -1 : 0;
}

const void *vf_ring_block_read_open (vf_ring *r, uint64_t *nbytes)
{
  uint64_t id = 0;
  if (ipcbuf_eod ((ipcbuf_t *) r->hdu->data_block)) return NULL;
  const char *b = ipcio_open_block_read (r->hdu->data_block, nbytes, &id);
  if (b) r->blocks_open++;
  return b;
}

int vf_ring_block_read_close (vf_ring *r)
{
  if (r->blocks_open <= 0) return -1;
  r->blocks_open--;
  return ipcio_close_block_read (r->hdu->data_block, vf_ring_get_bufsz (r)) < 0 ? -1 : 0;
}

int vf_ring_header_write (vf_ring *r, const char *hdr)
{
  if (lock_as (r, 2)) return -1;
  char *h = ipcbuf_get_next_write (r->hdu->header_block);
  if (!h) return -1;
  memcpy (h, hdr, VF_RING_HEADER_SIZE);
  return ipcbuf_mark_filled (r->hdu->header_block, VF_RING_HEADER_SIZE) < 0 ? -1 : 0;
}

ssize_t vf_ring_write (vf_ring *r, const void *src, size_t n)
{
  return ipcio_write (r->hdu->data_block, (char *) src, n);
}

int vf_ring_end_of_data (vf_ring *r)
{
  if (r->locked != 2) return 0;
  r->locked = 0;
  return dada_hdu_unlock_write (r->hdu) < 0 ? -1 : 0;
}

void vf_ring_shutdown (vf_ring *r) { (void) r; }
#endif /* HAVE_PSRDADA */
