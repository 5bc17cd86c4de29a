/*
 * genbase -- writes seconds of deterministic synthetic VLITE baseband as a VDIF
 * file (25600 frame pairs of 5032 bytes per second).  Stand-in for the
 * reference's src/genbase.cu (cuRAND noise + FFT dispersion, not reproducible
 * off the GPU): same frame layout and ordering (:443-486), same flag letters
 * where they mean the same thing (-t seconds, -r seed, -f RFI, -a amplitude,
 * -p period).
 */
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "vf_genbase.h"

int main (int argc, char **argv)
{
  vf_gen_params g;
  vf_gen_defaults (&g);
  int nsec = 1, antenna = 1, c;
  const char *out = NULL;
  while ((c = getopt (argc, argv, "ht:r:fa:p:n:o:")) != -1) {
    switch (c) {
      case 't': nsec = atoi (optarg); break;
      case 'r': g.seed = strtoull (optarg, NULL, 10); break;
      case 'f': g.rfi_amp = 60; g.rfi_burst_every = 16; break;
      case 'a': g.pulse_amp_q8[0] = (int) (atof (optarg) * 256 + 0.5); g.pulse_amp_q8[1] = g.pulse_amp_q8[0] / 10; break;
      case 'p': g.pulse_period = (int) (atof (optarg) * 128000000); g.pulse_width = (int) (0.03 * g.pulse_period); break;
      case 'n': antenna = atoi (optarg); break;
      case 'o': out = optarg; break;
      default:
        fprintf (stdout, "Usage: genbase -o FILE [-t seconds] [-r seed] [-f] [-a amp] [-p period_s] [-n station]\n");
        return c == 'h' ? 0 : 1;
    }
  }
  if (!out) { fprintf (stderr, "genbase: -o FILE is required\n"); return 1; }
  FILE *fp = fopen (out, "wb");
  if (!fp) { perror (out); return 1; }
  const size_t sec_bytes = (size_t) 25600 * 2 * 5032;
  unsigned char *scratch = malloc (256000000), *blk = malloc (sec_bytes);
  if (!scratch || !blk) return 1;
  for (int s = 0; s < nsec; ++s) {
    vf_gen_vdif_block (&g, antenna, (unsigned) (18000 + s), scratch, blk);
    if (fwrite (blk, 1, sec_bytes, fp) != sec_bytes) { perror ("fwrite"); return 1; }
  }
  fclose (fp);
  free (scratch); free (blk);
  return 0;
}
