/* CPU emulation of the kernel's FFT passes (one loop per barrier interval):
 * reads 2 x 12500 sample bytes from stdin, optional excision mask in argv[1]
 * (hex), writes float2[4096] detected powers (pol0, pol1) for bins 2155..6250
 * followed by float2[12500] Z in natural order (bins outside the stored range
 * [2155, 10345] are zero) to stdout.  Used by tests/test_fft_host.py. */
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include "vf_fft12500.cuh"

int main (int argc, char **argv)
{
  std::vector<uint8_t> b (25000);
  if (fread (b.data (), 1, 25000, stdin) != 25000) return 2;
  uint32_t mask = argc > 1 ? (uint32_t) strtoul (argv[1], 0, 16) : 0;
  const int lo = argc > 2 ? atoi (argv[2]) : 2155, hi = argc > 3 ? atoi (argv[3]) : 10345;
  /* sanitise (0 -> 128) word by word, as the kernel does */
  for (size_t i = 0; i < 25000; i += 4) {
    uint32_t w;
    memcpy (&w, &b[i], 4);
    w = vf_sanitise_word (w);
    memcpy (&b[i], &w, 4);
  }
  std::vector<float2> tw1 (500), tw5 (500), tw500 (500), W (VF_WLEN, make_float2 (0.f, 0.f));
  for (int p = 0; p < 500; ++p) {
    double a1 = -2.0 * M_PI * p / 12500.0, a5 = -2.0 * M_PI * 5 * p / 12500.0;
    tw1[p] = make_float2 ((float) cos (a1), (float) sin (a1));
    tw5[p] = make_float2 ((float) cos (a5), (float) sin (a5));
  }
  for (int k = 1; k < 25; ++k)
    for (int p = 0; p < 20; ++p) {
      double a = -2.0 * M_PI * (p * k) / 500.0;
      tw500[(k - 1) * 20 + p] = make_float2 ((float) cos (a), (float) sin (a));
    }
  vf_fft_tables tb = { tw1.data (), tw5.data (), tw500.data () };
  for (int p = 0; p < VF_NA; ++p) {
    if (mask) vf_pass1<true> (p, b.data (), b.data () + 12500, mask, tb, W.data ());
    else vf_pass1<false> (p, b.data (), b.data () + 12500, 0, tb, W.data ());
  }
  for (int i = 0; i < VF_NA; ++i) vf_pass2 (i, tb, W.data ());
  /* the store range is a compile-time parameter of pass 3: everything (tests of the whole spectrum) or the
   * kernel's range (the channels the filterbank keeps and their mirror images) */
  if (lo == 0 && hi == 12499) for (int m = 0; m < VF_NC; ++m) vf_pass3<0, 12499> (m, W.data ());
  else if (lo == 2155 && hi == 10345) for (int m = 0; m < VF_NC; ++m) vf_pass3<2155, 10345> (m, W.data ());
  else return 3;
  std::vector<float2> P (4096), Z (12500, make_float2 (0.f, 0.f));
  for (int c = 0; c < 4096; ++c) P[c] = vf_detect (c + 2155, W.data ());
  for (int k = lo; k <= hi; ++k) Z[k] = W[vf_zpos (k)];
  fwrite (P.data (), sizeof (float2), 4096, stdout);
  fwrite (Z.data (), sizeof (float2), 12500, stdout);
  return 0;
}
