/*
 * vf_dada_db -- creates or destroys a shared-memory ring, the job psrdada's
 * dada_db does for the reference (scripts/start_dada:13:
 * "dada_db -k 40 -b 257638400 -n 8 -l"): writers (genbase -k, the reference's
 * writer / genbase / readbase) and the reader (process_baseband -k) attach to
 * it by key.
 */
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include "vf_ring.h"

int main (int argc, char **argv)
{
  unsigned long key = 0x40;                       /* scripts/start_dada:13 */
  unsigned long long bufsz = 257638400ull, nbufs = 8;
  int destroy = 0, c;
  while ((c = getopt (argc, argv, "hk:b:n:dlp")) != -1) {
    switch (c) {
      case 'k': key = strtoul (optarg, NULL, 16); break;     /* psrdada keys are hexadecimal */
      case 'b': bufsz = strtoull (optarg, NULL, 10); break;
      case 'n': nbufs = strtoull (optarg, NULL, 10); break;
      case 'd': destroy = 1; break;
      case 'l': case 'p': break;                             /* lock / page-in: the reader page-locks the blocks itself */
      default:
        fprintf (stdout, "Usage: vf_dada_db [-k hexkey] [-b block_bytes] [-n nblocks] [-d destroy]\n");
        return c == 'h' ? 0 : 1;
    }
  }
  if (destroy) {
    if (vf_ring_remove_shm ((int) key)) { fprintf (stderr, "vf_dada_db: no ring with key %lx\n", key); return 1; }
    printf ("Destroyed DADA data block with key = %lx\n", key);
    return 0;
  }
  vf_ring *r = vf_ring_create_shm ((int) key, nbufs, bufsz);
  if (!r) { fprintf (stderr, "vf_dada_db: cannot create a ring with key %lx (does it exist already?)\n", key); return 1; }
  printf ("Created DADA data block with nbufs=%llu bufsz=%llu key=%lx\n", nbufs, bufsz, key);
  vf_ring_disown (r);
  vf_ring_destroy (r);
  return 0;
}
