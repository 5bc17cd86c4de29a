/* CPU emulation of the kernel's FFT passes (one loop per barrier interval):
 * reads 2 x 12500 sample bytes from stdin, optional excision mask in argv[1]
 * (hex), writes float2[4096] detected powers (pol0, pol1) for bins 2155..6250
 * followed by float2[12500] raw Z to stdout.  Used by tests/test_fft_host.py. */
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "vf_fft12500.cuh"

int main (int argc, char **argv)
{
  std::vector<uint8_t> b (25000);
  if (fread (b.data (), 1, 25000, stdin) != 25000) return 2;
  uint32_t mask = argc > 1 ? (uint32_t) strtoul (argv[1], 0, 16) : 0;
  std::vector<float2> tw1 (500), tw5 (500), tw500 (500), W (12500);
  for (int p = 0; p < 500; ++p) {
    double a1 = -2.0 * M_PI * p / 12500.0, a5 = -2.0 * M_PI * 5 * p / 12500.0, a500 = -2.0 * M_PI * p / 500.0;
    tw1[p] = make_float2 ((float) cos (a1), (float) sin (a1));
    tw5[p] = make_float2 ((float) cos (a5), (float) sin (a5));
    tw500[p] = make_float2 ((float) cos (a500), (float) sin (a500));
  }
  vf_fft_tables tb = { tw1.data (), tw5.data (), tw500.data () };
  for (int p = 0; p < VF_NA; ++p) vf_pass_a (p, b.data (), b.data () + 12500, mask, tb, W.data ());
  {
    std::vector<float2> regs (500 * 25);
    for (int i = 0; i < VF_NA; ++i) { float2 v[25]; vf_pass_b_load (i, W.data (), v); for (int j = 0; j < 25; ++j) regs[i * 25 + j] = v[j]; }
    for (int i = 0; i < VF_NA; ++i) { float2 v[25]; for (int j = 0; j < 25; ++j) v[j] = regs[i * 25 + j]; vf_pass_b_store (i, v, tb, W.data ()); }
  }
  {
    std::vector<float2> regs (625 * 20);
    for (int q = 0; q < VF_NC; ++q) { float2 v[20]; vf_pass_c_load (q, W.data (), v); for (int j = 0; j < 20; ++j) regs[q * 20 + j] = v[j]; }
    for (int q = 0; q < VF_NC; ++q) { float2 v[20]; for (int j = 0; j < 20; ++j) v[j] = regs[q * 20 + j]; vf_pass_c_store (q, v, W.data (), 0, 12499); }
  }
  std::vector<float2> P (4096);
  for (int c = 0; c < 4096; ++c) P[c] = vf_detect (c + 2155, W.data ());
  fwrite (P.data (), sizeof (float2), 4096, stdout);
  fwrite (W.data (), sizeof (float2), 12500, stdout);
  return 0;
}
