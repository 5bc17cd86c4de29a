#include <errno.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <sys/ipc.h>
#include <sys/shm.h>
#include <time.h>
#include "vf_ring.h"

#define VF_RING_MAGIC 0x56465247u   /* "VFRG" */
#define VF_RING_MAXBUF 1024

/* control block: lives in front of the data blocks, in malloc'd memory (in-process ring) or in a
 * SysV shared-memory segment (ring shared between processes, as psrdada's dada_db makes them) */
typedef struct {
  uint32_t magic;
  uint64_t nbufs, bufsz;
  uint64_t fill[VF_RING_MAXBUF];      /* bytes valid in each block */
  unsigned char eod[VF_RING_MAXBUF];  /* block is the last one of its observation */
  uint64_t w_idx, r_idx;              /* blocks written / released so far */
  uint64_t p_idx;                     /* blocks handed to the reader so far (>= r_idx: several may be open) */
  uint64_t w_off, r_off;              /* byte-stream offsets inside the open block */
  int w_open, r_open;
  int at_eod;                         /* the reader has consumed the EOD block */
  char header[VF_RING_HEADER_SIZE];
  int header_full;
  int shut;
  pthread_mutex_t mu;
  pthread_cond_t cv;
} vf_ring_ctl;

struct vf_ring {
  vf_ring_ctl *c;
  unsigned char *mem;          /* data blocks */
  int own_ctl, own_mem;        /* malloc'd here */
  int shmid;                   /* >= 0: attached SysV segment */
  int creator;
};

static size_t ctl_bytes (void) { return (sizeof (vf_ring_ctl) + 4095) & ~(size_t) 4095; }

static int ctl_init (vf_ring_ctl *c, uint64_t nbufs, uint64_t bufsz, int shared)
{
  memset (c, 0, sizeof (*c));
  c->nbufs = nbufs; c->bufsz = bufsz;
  pthread_mutexattr_t ma;
  pthread_condattr_t ca;
  pthread_mutexattr_init (&ma);
  pthread_condattr_init (&ca);
  if (shared) {
    pthread_mutexattr_setpshared (&ma, PTHREAD_PROCESS_SHARED);
    pthread_condattr_setpshared (&ca, PTHREAD_PROCESS_SHARED);
  }
  int rc = pthread_mutex_init (&c->mu, &ma) | pthread_cond_init (&c->cv, &ca);
  pthread_mutexattr_destroy (&ma);
  pthread_condattr_destroy (&ca);
  c->magic = VF_RING_MAGIC;
  return rc;
}

vf_ring *vf_ring_create (uint64_t nbufs, uint64_t bufsz, void *mem)
{
  if (!nbufs || !bufsz || nbufs > VF_RING_MAXBUF) return NULL;
  vf_ring *r = (vf_ring *) calloc (1, sizeof (*r));
  if (!r) return NULL;
  r->shmid = -1;
  r->c = (vf_ring_ctl *) malloc (sizeof (vf_ring_ctl));
  r->own_ctl = 1;
  r->mem = (unsigned char *) mem;
  if (!mem) { r->mem = (unsigned char *) malloc (nbufs * bufsz); r->own_mem = 1; }
  if (!r->c || !r->mem || ctl_init (r->c, nbufs, bufsz, 0)) { vf_ring_destroy (r); return NULL; }
  return r;
}

/* dada_db -k key -b bufsz -n nbufs: creates the segment; fails if it exists */
vf_ring *vf_ring_create_shm (int key, uint64_t nbufs, uint64_t bufsz)
{
  if (!nbufs || !bufsz || nbufs > VF_RING_MAXBUF) return NULL;
  const size_t total = ctl_bytes () + (size_t) (nbufs * bufsz);
  int id = shmget ((key_t) key, total, IPC_CREAT | IPC_EXCL | 0666);
  if (id < 0) return NULL;
  void *base = shmat (id, NULL, 0);
  if (base == (void *) -1) { shmctl (id, IPC_RMID, NULL); return NULL; }
  vf_ring *r = (vf_ring *) calloc (1, sizeof (*r));
  r->shmid = id; r->creator = 1;
  r->c = (vf_ring_ctl *) base;
  r->mem = (unsigned char *) base + ctl_bytes ();
  if (ctl_init (r->c, nbufs, bufsz, 1)) { vf_ring_destroy (r); return NULL; }
  return r;
}

/* dada_hdu_connect: attaches to a segment made by vf_ring_create_shm */
vf_ring *vf_ring_connect_shm (int key)
{
  int id = shmget ((key_t) key, 0, 0666);
  if (id < 0) return NULL;
  void *base = shmat (id, NULL, 0);
  if (base == (void *) -1) return NULL;
  vf_ring_ctl *c = (vf_ring_ctl *) base;
  if (c->magic != VF_RING_MAGIC) { shmdt (base); return NULL; }
  vf_ring *r = (vf_ring *) calloc (1, sizeof (*r));
  r->shmid = id;
  r->c = c;
  r->mem = (unsigned char *) base + ctl_bytes ();
  return r;
}

/* dada_db -d: marks the segment for removal (it disappears when the last process detaches) */
int vf_ring_remove_shm (int key)
{
  int id = shmget ((key_t) key, 0, 0666);
  if (id < 0) return -1;
  return shmctl (id, IPC_RMID, NULL);
}

void vf_ring_disown (vf_ring *r) { if (r) r->creator = 0; }

void vf_ring_destroy (vf_ring *r)
{
  if (!r) return;
  if (r->shmid >= 0) {
    if (r->c) shmdt (r->c);
    if (r->creator) shmctl (r->shmid, IPC_RMID, NULL);
  } else {
    if (r->c) { pthread_mutex_destroy (&r->c->mu); pthread_cond_destroy (&r->c->cv); }
    if (r->own_ctl) free (r->c);
    if (r->own_mem) free (r->mem);
  }
  free (r);
}

uint64_t vf_ring_get_nbufs (const vf_ring *r) { return r->c->nbufs; }
uint64_t vf_ring_get_bufsz (const vf_ring *r) { return r->c->bufsz; }
void *vf_ring_data_base (const vf_ring *r) { return r->mem; }

uint64_t vf_ring_get_nfull (vf_ring *r)
{
  vf_ring_ctl *c = r->c;
  pthread_mutex_lock (&c->mu);
  uint64_t n = c->w_idx - c->r_idx;
  pthread_mutex_unlock (&c->mu);
  return n;
}

void vf_ring_shutdown (vf_ring *r)
{
  vf_ring_ctl *c = r->c;
  pthread_mutex_lock (&c->mu);
  c->shut = 1;
  pthread_cond_broadcast (&c->cv);
  pthread_mutex_unlock (&c->mu);
}

static void deadline (struct timespec *ts, int ms)
{
  clock_gettime (CLOCK_REALTIME, ts);
  ts->tv_sec += ms / 1000;
  ts->tv_nsec += (long) (ms % 1000) * 1000000L;
  if (ts->tv_nsec >= 1000000000L) { ts->tv_sec++; ts->tv_nsec -= 1000000000L; }
}

int vf_ring_header_write (vf_ring *r, const char *hdr)
{
  vf_ring_ctl *c = r->c;
  pthread_mutex_lock (&c->mu);
  while (c->header_full && !c->shut) pthread_cond_wait (&c->cv, &c->mu);
  if (c->shut) { pthread_mutex_unlock (&c->mu); return -1; }
  memset (c->header, 0, sizeof (c->header));
  strncpy (c->header, hdr, sizeof (c->header) - 1);
  c->header_full = 1;
  pthread_cond_broadcast (&c->cv);
  pthread_mutex_unlock (&c->mu);
  return 0;
}

int vf_ring_header_read (vf_ring *r, char *hdr, int timeout_ms)
{
  vf_ring_ctl *c = r->c;
  struct timespec ts;
  deadline (&ts, timeout_ms < 0 ? 0 : timeout_ms);
  pthread_mutex_lock (&c->mu);
  while (!c->header_full && !c->shut) {
    if (timeout_ms < 0) pthread_cond_wait (&c->cv, &c->mu);
    else if (pthread_cond_timedwait (&c->cv, &c->mu, &ts) == ETIMEDOUT) { pthread_mutex_unlock (&c->mu); return 1; }
  }
  if (!c->header_full) { pthread_mutex_unlock (&c->mu); return -1; }
  memcpy (hdr, c->header, sizeof (c->header));
  c->header_full = 0;
  c->at_eod = 0;
  pthread_cond_broadcast (&c->cv);
  pthread_mutex_unlock (&c->mu);
  return 0;
}

void *vf_ring_block_write_open (vf_ring *r)
{
  vf_ring_ctl *c = r->c;
  pthread_mutex_lock (&c->mu);
  while (c->w_idx - c->r_idx >= c->nbufs && !c->shut) pthread_cond_wait (&c->cv, &c->mu);
  void *p = c->shut ? NULL : r->mem + (c->w_idx % c->nbufs) * c->bufsz;
  if (p) { c->w_open = 1; c->w_off = 0; }
  pthread_mutex_unlock (&c->mu);
  return p;
}

static int close_write (vf_ring *r, uint64_t nbytes, int eod)
{
  vf_ring_ctl *c = r->c;
  pthread_mutex_lock (&c->mu);
  if (!c->w_open) { pthread_mutex_unlock (&c->mu); return -1; }
  const uint64_t b = c->w_idx % c->nbufs;
  c->fill[b] = nbytes; c->eod[b] = (unsigned char) eod;
  c->w_idx++; c->w_open = 0; c->w_off = 0;
  pthread_cond_broadcast (&c->cv);
  pthread_mutex_unlock (&c->mu);
  return 0;
}

int vf_ring_block_write_close (vf_ring *r, uint64_t nbytes)
{
  if (nbytes > r->c->bufsz) return -1;
  return close_write (r, nbytes, 0);
}

ssize_t vf_ring_write (vf_ring *r, const void *src, size_t n)
{
  vf_ring_ctl *c = r->c;
  const unsigned char *s = (const unsigned char *) src;
  size_t left = n;
  while (left) {
    if (!c->w_open && !vf_ring_block_write_open (r)) return -1;
    unsigned char *blk = r->mem + (c->w_idx % c->nbufs) * c->bufsz;
    size_t room = (size_t) (c->bufsz - c->w_off), take = left < room ? left : room;
    memcpy (blk + c->w_off, s, take);
    c->w_off += take; s += take; left -= take;
    if (c->w_off == c->bufsz && close_write (r, c->bufsz, 0)) return -1;
  }
  return (ssize_t) n;
}

int vf_ring_end_of_data (vf_ring *r)
{
  /* the EOD marker travels with a block: a partial one if bytes are pending, else an empty one */
  if (!r->c->w_open && !vf_ring_block_write_open (r)) return -1;
  return close_write (r, r->c->w_off, 1);
}

const void *vf_ring_block_read_open (vf_ring *r, uint64_t *nbytes)
{
  vf_ring_ctl *c = r->c;
  pthread_mutex_lock (&c->mu);
  for (;;) {
    if (c->at_eod || c->shut) { pthread_mutex_unlock (&c->mu); if (nbytes) *nbytes = 0; return NULL; }
    if (c->w_idx > c->p_idx) {
      const uint64_t b = c->p_idx % c->nbufs;
      if (c->fill[b] == 0 && c->eod[b] && c->p_idx == c->r_idx) {   /* empty EOD block, nothing else open */
        c->r_idx++; c->p_idx++; c->at_eod = 1;
        pthread_cond_broadcast (&c->cv);
        continue;
      }
      if (c->fill[b] == 0 && c->eod[b]) {          /* empty EOD block behind open blocks: end of data for now */
        pthread_mutex_unlock (&c->mu);
        if (nbytes) *nbytes = 0;
        return NULL;
      }
      c->p_idx++; c->r_open++; c->r_off = 0;
      if (nbytes) *nbytes = c->fill[b];
      pthread_mutex_unlock (&c->mu);
      return r->mem + b * c->bufsz;
    }
    pthread_cond_wait (&c->cv, &c->mu);
  }
}

int vf_ring_block_read_close (vf_ring *r)
{
  vf_ring_ctl *c = r->c;
  pthread_mutex_lock (&c->mu);
  if (!c->r_open) { pthread_mutex_unlock (&c->mu); return -1; }
  const uint64_t b = c->r_idx % c->nbufs;      /* blocks are released oldest first */
  if (c->eod[b]) c->at_eod = 1;
  c->r_idx++; c->r_open--; c->r_off = 0;
  pthread_cond_broadcast (&c->cv);
  pthread_mutex_unlock (&c->mu);
  return 0;
}

ssize_t vf_ring_read (vf_ring *r, void *dst, size_t n)
{
  vf_ring_ctl *c = r->c;
  unsigned char *d = (unsigned char *) dst;
  size_t got = 0;
  while (got < n) {
    if (!c->r_open) {
      uint64_t nb;
      if (!vf_ring_block_read_open (r, &nb)) break;       /* EOD */
    }
    const uint64_t b = c->r_idx % c->nbufs;
    size_t avail = (size_t) (c->fill[b] - c->r_off), take = (n - got) < avail ? (n - got) : avail;
    memcpy (d + got, r->mem + b * c->bufsz + c->r_off, take);
    c->r_off += take; got += take;
    if (c->r_off == c->fill[b]) vf_ring_block_read_close (r);
  }
  return (ssize_t) got;
}
