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
#include "vf_ring.h"
#include "vf_sigproc.h"
#include "vlitegen.h"

int main (int argc, char **argv)
{
  vf_gen_params g;
  vf_gen_defaults (&g);
  int nsec = 1, antenna = 1, c;
  long key = -1;
  const char *out = NULL;
  float pol_ratio = 1.0f;           /* -s, src/genbase.cu:85 */
  int gpu = 0;                      /* -G: noise + coherent dispersion on the GPU (libvlitegen), as src/genbase.cu does */
  vfg_config gc;
  vfg_config_default (&gc);
  while ((c = getopt (argc, argv, "ht:r:fa:p:n:o:k:Gd:s:")) != -1) {
    switch (c) {
      case 't': nsec = atoi (optarg); break;
      case 'r': g.seed = strtoull (optarg, NULL, 10); gc.seed = g.seed; break;
      case 'f': g.rfi_amp = 60; g.rfi_burst_every = 16; gc.add_rfi = 1; break;
      case 'a': g.pulse_amp_q8[0] = (int) (atof (optarg) * 256 + 0.5); g.pulse_amp_q8[1] = g.pulse_amp_q8[0] / 10; gc.ampl[0] = (float) atof (optarg); break;
      case 'p': g.pulse_period = (int) (atof (optarg) * 128000000); g.pulse_width = (int) (0.03 * g.pulse_period); gc.pulse_period = atof (optarg); break;
      case 'n': antenna = atoi (optarg); break;
      case 'o': out = optarg; break;
      case 'G': gpu = 1; break;
      case 'd': gc.dm = atof (optarg); break;                      /* -G only, src/genbase.cu:131-137 */
      case 's': pol_ratio = (float) atof (optarg); break;          /* pol 1 / pol 0 amplitude ratio (:171) */
      case 'k': key = (long) strtoul (optarg, NULL, 16); break;    /* write to this ring (src/genbase.cu:300-353) */
      default:
        fprintf (stdout, "Usage: genbase (-o FILE | -k hexkey) [-t seconds] [-r seed] [-f] [-a amp] [-p period_s] [-n station]\n"
                         "       -G  generate on the GPU with coherent dispersion (-d DM [30], -s pol-1 amplitude ratio [1])\n");
        return c == 'h' ? 0 : 1;
    }
  }
  const size_t sec_bytes = (size_t) 25600 * 2 * 5032;
  unsigned char *scratch = malloc (256000000);
  if (!scratch) return 1;
  vfg_handle *gh = NULL;
  if (gpu) {
    gc.ampl[1] = gc.ampl[0] * pol_ratio;
    gc.buflen = 0;                  /* as long as the sweep needs */
    if (vfg_create (&gc, &gh)) { fprintf (stderr, "genbase: %s\n", gh ? vfg_last_error (gh) : "no CUDA device"); return 20; }
  }
  if (key >= 0) {
    /* one observation into the ring: header block, nsec one-second blocks, end of data */
    vf_ring *ring = vf_ring_connect_shm ((int) key);
    if (!ring) { fprintf (stderr, "genbase: no ring with key %lx (vf_dada_db -k %lx)\n", key, key); return 1; }
    if (vf_ring_get_bufsz (ring) < sec_bytes) { fprintf (stderr, "genbase: ring blocks are smaller than one second of VDIF\n"); return 1; }
    char hdr[VF_RING_HEADER_SIZE] = "";
    vf_ascii_header_set (hdr, sizeof (hdr), "NAME", "%s", "GENBASE");              /* src/genbase.cu:331-353 */
    vf_ascii_header_set (hdr, sizeof (hdr), "STATIONID", "%d", antenna);
    vf_ascii_header_set (hdr, sizeof (hdr), "NCHAN", "%d", 1);
    vf_ascii_header_set (hdr, sizeof (hdr), "BANDWIDTH", "%lf", -64.0);
    vf_ascii_header_set (hdr, sizeof (hdr), "CFREQ", "%lf", 352.0);
    vf_ascii_header_set (hdr, sizeof (hdr), "NPOL", "%d", 2);
    vf_ascii_header_set (hdr, sizeof (hdr), "NBIT", "%d", 8);
    vf_ascii_header_set (hdr, sizeof (hdr), "RA", "%lf", 0.87180);
    vf_ascii_header_set (hdr, sizeof (hdr), "DEC", "%lf", 0.72452);
    if (vf_ring_header_write (ring, hdr)) return 1;
    for (int s = 0; s < nsec; ++s) {
      unsigned char *b = (unsigned char *) vf_ring_block_write_open (ring);
      if (!b) return 1;
      if (gh) { if (vfg_generate_vdif_second (gh, antenna, (unsigned) (18000 + s), b)) { fprintf (stderr, "genbase: %s\n", vfg_last_error (gh)); return 20; } }
      else vf_gen_vdif_block (&g, antenna, (unsigned) (18000 + s), scratch, b);
      vf_ring_block_write_close (ring, sec_bytes);
    }
    vf_ring_end_of_data (ring);
    vf_ring_destroy (ring);
    free (scratch);
    return 0;
  }
  if (!out) { fprintf (stderr, "genbase: -o FILE or -k KEY is required\n"); return 1; }
  FILE *fp = fopen (out, "wb");
  if (!fp) { perror (out); return 1; }
  unsigned char *blk = malloc (sec_bytes);
  if (!blk) return 1;
  for (int s = 0; s < nsec; ++s) {
    if (gh) { if (vfg_generate_vdif_second (gh, antenna, (unsigned) (18000 + s), blk)) { fprintf (stderr, "genbase: %s\n", vfg_last_error (gh)); return 20; } }
    else vf_gen_vdif_block (&g, antenna, (unsigned) (18000 + s), scratch, blk);
    if (fwrite (blk, 1, sec_bytes, fp) != sec_bytes) { perror ("fwrite"); return 1; }
  }
  fclose (fp);
  free (scratch); free (blk);
  return 0;
}
