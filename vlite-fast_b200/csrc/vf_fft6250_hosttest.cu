/* CPU emulation of the passes of the 6250-point real-input FFT (vf_fft6250.cuh), one loop per barrier interval:
 * reads 12500 sample bytes of one polarisation from stdin, optional excision mask in argv[1] (hex), writes
 * float[4096] detected powers of bins 2155..6250 followed by float2[6250] Z in natural order.
 * Used by tests/test_fft_host.py. */
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include "vf_fft6250.cuh"

int main (int argc, char **argv)
{
  std::vector<uint8_t> b (12500);
  if (fread (b.data (), 1, 12500, stdin) != 12500) return 2;
  uint32_t mask = argc > 1 ? (uint32_t) strtoul (argv[1], 0, 16) : 0;
  for (size_t i = 0; i < 12500; i += 4) {
    uint32_t w;
    memcpy (&w, &b[i], 4);
    w = vf_sanitise_word (w);
    memcpy (&b[i], &w, 4);
  }
  std::vector<float2> tw1 (250), tw5 (250), tw250 (240), tws (6250), W (VF6_WLEN, make_float2 (0.f, 0.f));
  for (int p = 0; p < 250; ++p) {
    double a1 = -2.0 * M_PI * p / 6250.0, a5 = -2.0 * M_PI * 5 * p / 6250.0;
    tw1[p] = make_float2 ((float) cos (a1), (float) sin (a1));
    tw5[p] = make_float2 ((float) cos (a5), (float) sin (a5));
  }
  for (int k = 1; k < 25; ++k)
    for (int p = 0; p < 10; ++p) {
      double a = -2.0 * M_PI * (p * k) / 250.0;
      tw250[(k - 1) * 10 + p] = make_float2 ((float) cos (a), (float) sin (a));
    }
  for (int k = 0; k < 6250; ++k) {
    double a = -2.0 * M_PI * k / 12500.0;
    tws[k] = make_float2 ((float) cos (a), (float) sin (a));
  }
  vf6_tables tb = { tw1.data (), tw5.data (), tw250.data (), tws.data () };
  for (int p = 0; p < VF6_NA; ++p) {
    if (mask) vf6_pass1<true> (p, b.data (), mask, tb, W.data ());
    else vf6_pass1<false> (p, b.data (), 0, tb, W.data ());
  }
  for (int i = 0; i < VF6_NA; ++i) vf6_pass2 (i, tb, W.data ());
  /* argv[2] = "fused": pass 3 fused with the split pass (what the kernel runs), powers straight from registers */
  const bool fused = argc > 2 && !strcmp (argv[2], "fused");
  std::vector<float> P (4096), P2 (2 * 4096, -1.f);
  if (fused) for (int u = 0; u < VF6_NU; ++u) vf6_pass3_split (u, W.data (), tws.data (), P2.data ());
  /* argv[2] = "count": how many units write each channel (must be exactly one: two units writing the same channel
   * would race in the kernel).  Each unit runs into a buffer of its own, pre-filled with -1. */
  if (argc > 2 && !strcmp (argv[2], "count")) {
    std::vector<int> cnt (4096, 0);
    for (int u = 0; u < VF6_NU; ++u) {
      std::vector<float> Q (2 * 4096, -1.f);
      vf6_pass3_split (u, W.data (), tws.data (), Q.data ());
      for (int c = 0; c < 4096; ++c) { if (Q[2 * c] >= 0.f) cnt[c]++; if (Q[2 * c + 1] >= 0.f) cnt[c] += 1000; }   /* odd slots: the other polarisation's */
    }
    fwrite (cnt.data (), sizeof (int), 4096, stdout);
    return 0;
  }
  for (int m = 0; m < VF6_NC; ++m) vf6_pass3 (m, W.data ());
  std::vector<float2> Z (6250);
  for (int c = 0; c < 4096; ++c) P[c] = fused ? P2[2 * c] : vf6_detect (c, W.data (), tws.data ());
  for (int k = 0; k < 6250; ++k) Z[k] = W[vf6_zpos (k)];
  fwrite (P.data (), sizeof (float), 4096, stdout);
  fwrite (Z.data (), sizeof (float2), 6250, stdout);
  return 0;
}
