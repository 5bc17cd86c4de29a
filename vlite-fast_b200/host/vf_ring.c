#include <errno.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "vf_ring.h"

struct vf_ring {
  uint64_t nbufs, bufsz;
  unsigned char *mem;
  int own_mem;
  uint64_t *fill;               /* bytes valid in each block */
  unsigned char *eod;           /* block is the last one of its observation */
  uint64_t w_idx, r_idx;        /* blocks written / released so far */
  uint64_t p_idx;               /* blocks handed to the reader so far (>= r_idx: several may be open) */
  uint64_t w_off;               /* byte-stream writer: offset inside the open block */
  uint64_t r_off;               /* byte-stream reader: offset inside the open block */
  int w_open, r_open;
  int at_eod;                   /* the reader has consumed the EOD block */
  char header[VF_RING_HEADER_SIZE];
  int header_full;
  int shut;
  pthread_mutex_t mu;
  pthread_cond_t cv;
};

vf_ring *vf_ring_create (uint64_t nbufs, uint64_t bufsz, void *mem)
{
  if (!nbufs || !bufsz) return NULL;
  vf_ring *r = (vf_ring *) calloc (1, sizeof (*r));
  if (!r) return NULL;
  r->nbufs = nbufs; r->bufsz = bufsz;
  r->mem = (unsigned char *) mem;
  if (!mem) { r->mem = (unsigned char *) malloc (nbufs * bufsz); r->own_mem = 1; }
  r->fill = (uint64_t *) calloc (nbufs, sizeof (uint64_t));
  r->eod = (unsigned char *) calloc (nbufs, 1);
  if (!r->mem || !r->fill || !r->eod) { vf_ring_destroy (r); return NULL; }
  pthread_mutex_init (&r->mu, NULL);
  pthread_cond_init (&r->cv, NULL);
  return r;
}

void vf_ring_destroy (vf_ring *r)
{
  if (!r) return;
  if (r->own_mem) free (r->mem);
  free (r->fill); free (r->eod);
  pthread_mutex_destroy (&r->mu);
  pthread_cond_destroy (&r->cv);
  free (r);
}

uint64_t vf_ring_get_nbufs (const vf_ring *r) { return r->nbufs; }
uint64_t vf_ring_get_bufsz (const vf_ring *r) { return r->bufsz; }

uint64_t vf_ring_get_nfull (vf_ring *r)
{
  pthread_mutex_lock (&r->mu);
  uint64_t n = r->w_idx - r->r_idx;
  pthread_mutex_unlock (&r->mu);
  return n;
}

void vf_ring_shutdown (vf_ring *r)
{
  pthread_mutex_lock (&r->mu);
  r->shut = 1;
  pthread_cond_broadcast (&r->cv);
  pthread_mutex_unlock (&r->mu);
}

int vf_ring_header_write (vf_ring *r, const char *hdr)
{
  pthread_mutex_lock (&r->mu);
  while (r->header_full && !r->shut) pthread_cond_wait (&r->cv, &r->mu);
  if (r->shut) { pthread_mutex_unlock (&r->mu); return -1; }
  memset (r->header, 0, sizeof (r->header));
  strncpy (r->header, hdr, sizeof (r->header) - 1);
  r->header_full = 1;
  pthread_cond_broadcast (&r->cv);
  pthread_mutex_unlock (&r->mu);
  return 0;
}

int vf_ring_header_read (vf_ring *r, char *hdr, int timeout_ms)
{
  struct timespec ts;
  clock_gettime (CLOCK_REALTIME, &ts);
  ts.tv_sec += timeout_ms / 1000;
  ts.tv_nsec += (long) (timeout_ms % 1000) * 1000000L;
  if (ts.tv_nsec >= 1000000000L) { ts.tv_sec++; ts.tv_nsec -= 1000000000L; }
  pthread_mutex_lock (&r->mu);
  while (!r->header_full && !r->shut) {
    if (timeout_ms < 0) pthread_cond_wait (&r->cv, &r->mu);
    else if (pthread_cond_timedwait (&r->cv, &r->mu, &ts) == ETIMEDOUT) { pthread_mutex_unlock (&r->mu); return 1; }
  }
  if (!r->header_full) { pthread_mutex_unlock (&r->mu); return -1; }
  memcpy (hdr, r->header, sizeof (r->header));
  r->header_full = 0;
  r->at_eod = 0;
  pthread_cond_broadcast (&r->cv);
  pthread_mutex_unlock (&r->mu);
  return 0;
}

void *vf_ring_block_write_open (vf_ring *r)
{
  pthread_mutex_lock (&r->mu);
  while (r->w_idx - r->r_idx >= r->nbufs && !r->shut) pthread_cond_wait (&r->cv, &r->mu);
  void *p = r->shut ? NULL : r->mem + (r->w_idx % r->nbufs) * r->bufsz;
  if (p) { r->w_open = 1; r->w_off = 0; }
  pthread_mutex_unlock (&r->mu);
  return p;
}

static int close_write (vf_ring *r, uint64_t nbytes, int eod)
{
  pthread_mutex_lock (&r->mu);
  if (!r->w_open) { pthread_mutex_unlock (&r->mu); return -1; }
  const uint64_t b = r->w_idx % r->nbufs;
  r->fill[b] = nbytes; r->eod[b] = (unsigned char) eod;
  r->w_idx++; r->w_open = 0; r->w_off = 0;
  pthread_cond_broadcast (&r->cv);
  pthread_mutex_unlock (&r->mu);
  return 0;
}

int vf_ring_block_write_close (vf_ring *r, uint64_t nbytes)
{
  if (nbytes > r->bufsz) return -1;
  return close_write (r, nbytes, 0);
}

ssize_t vf_ring_write (vf_ring *r, const void *src, size_t n)
{
  const unsigned char *s = (const unsigned char *) src;
  size_t left = n;
  while (left) {
    if (!r->w_open && !vf_ring_block_write_open (r)) return -1;
    unsigned char *blk = r->mem + (r->w_idx % r->nbufs) * r->bufsz;
    size_t room = (size_t) (r->bufsz - r->w_off), take = left < room ? left : room;
    memcpy (blk + r->w_off, s, take);
    r->w_off += take; s += take; left -= take;
    if (r->w_off == r->bufsz && close_write (r, r->bufsz, 0)) return -1;
  }
  return (ssize_t) n;
}

int vf_ring_end_of_data (vf_ring *r)
{
  /* the EOD marker travels with a block: a partial one if bytes are pending, else an empty one */
  if (!r->w_open && !vf_ring_block_write_open (r)) return -1;
  return close_write (r, r->w_off, 1);
}

const void *vf_ring_block_read_open (vf_ring *r, uint64_t *nbytes)
{
  pthread_mutex_lock (&r->mu);
  for (;;) {
    if (r->at_eod || r->shut) { pthread_mutex_unlock (&r->mu); if (nbytes) *nbytes = 0; return NULL; }
    if (r->w_idx > r->p_idx) {
      const uint64_t b = r->p_idx % r->nbufs;
      if (r->fill[b] == 0 && r->eod[b] && r->p_idx == r->r_idx) {   /* empty EOD block, nothing else open */
        r->r_idx++; r->p_idx++; r->at_eod = 1;
        pthread_cond_broadcast (&r->cv);
        continue;
      }
      if (r->fill[b] == 0 && r->eod[b]) {          /* empty EOD block behind open blocks: end of data for now */
        pthread_mutex_unlock (&r->mu);
        if (nbytes) *nbytes = 0;
        return NULL;
      }
      r->p_idx++; r->r_open++; r->r_off = 0;
      if (nbytes) *nbytes = r->fill[b];
      pthread_mutex_unlock (&r->mu);
      return r->mem + b * r->bufsz;
    }
    pthread_cond_wait (&r->cv, &r->mu);
  }
}

int vf_ring_block_read_close (vf_ring *r)
{
  pthread_mutex_lock (&r->mu);
  if (!r->r_open) { pthread_mutex_unlock (&r->mu); return -1; }
  const uint64_t b = r->r_idx % r->nbufs;      /* blocks are released oldest first */
  if (r->eod[b]) r->at_eod = 1;
  r->r_idx++; r->r_open--; r->r_off = 0;
  pthread_cond_broadcast (&r->cv);
  pthread_mutex_unlock (&r->mu);
  return 0;
}

ssize_t vf_ring_read (vf_ring *r, void *dst, size_t n)
{
  unsigned char *d = (unsigned char *) dst;
  size_t got = 0;
  while (got < n) {
    if (!r->r_open) {
      uint64_t nb;
      if (!vf_ring_block_read_open (r, &nb)) break;       /* EOD */
    }
    const uint64_t b = r->r_idx % r->nbufs;
    size_t avail = (size_t) (r->fill[b] - r->r_off), take = (n - got) < avail ? (n - got) : avail;
    memcpy (d + got, r->mem + b * r->bufsz + r->r_off, take);
    r->r_off += take; got += take;
    if (r->r_off == r->fill[b]) vf_ring_block_read_close (r);
  }
  return (ssize_t) got;
}
