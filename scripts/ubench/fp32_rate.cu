// Issue-rate microbenchmark for the packed fp32 instructions of sm_100a (FFMA2 / FADD2 / FMUL2)
// against scalar FFMA / FADD: warp-instructions per cycle per SM at 4..20 resident warps.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp32_rate fp32_rate.cu && ./fp32_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 fma2 (float2 a, float2 b, float2 c)
{
  float2 r;
  asm volatile ("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
       : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
__device__ __forceinline__ float2 add2 (float2 a, float2 b)
{
  float2 r;
  asm volatile ("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
       : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 mul2 (float2 a, float2 b)
{
  float2 r;
  asm volatile ("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
       : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float sfma (float a, float b, float c)
{
  float r; asm volatile ("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
}
__device__ __forceinline__ float sadd (float a, float b)
{
  float r; asm volatile ("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r;
}

#define NACC 8
#define ITER 512
template <int MODE>
__global__ void k (float *out, long long *cyc, float bx, float cx)
{
  float2 acc[NACC];
  for (int i = 0; i < NACC; ++i) acc[i] = make_float2 (threadIdx.x * 0.001f + i, i * 0.5f);
  float2 b = make_float2 (bx, bx * 1.0001f), c = make_float2 (cx, cx * 0.999f);
  __syncthreads ();
  long long t0 = clock64 ();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        if (MODE == 0) acc[i] = fma2 (acc[i], b, c);                                    // FFMA2 r,r,r
        if (MODE == 1) acc[i] = add2 (acc[i], b);                                       // FADD2
        if (MODE == 2) acc[i] = mul2 (acc[i], b);                                       // FMUL2
        if (MODE == 3) { acc[i].x = sfma (acc[i].x, b.x, c.x); acc[i].y = sfma (acc[i].y, b.y, c.y); }   // 2 x FFMA r,r,r
        if (MODE == 4) { acc[i].x = sadd (acc[i].x, b.x); acc[i].y = sadd (acc[i].y, b.y); }             // 2 x FADD
        if (MODE == 5) acc[i] = fma2 (acc[i], make_float2 (0.999f, 0.999f), c);         // FFMA2 with an immediate
        if (MODE == 6) { acc[i].x = sfma (acc[i].x, 0.999f, c.x); acc[i].y = sfma (acc[i].y, 0.999f, c.y); } // FFMA imm
        if (MODE == 7) acc[i] = add2 (acc[i], make_float2 (b.y, -b.x));                 // FADD2 with swap + negate
      }
  }
  long long t1 = clock64 ();
  float s = 0;
  for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// Dependent-issue latency: ONE warp, one accumulator, every instruction waits for the one before.
template <int MODE>
__global__ void klat (float *out, long long *cyc, float bx, float cx)
{
  float2 a = make_float2 (threadIdx.x * 0.001f, 0.5f);
  const float2 b = make_float2 (bx, bx * 1.0001f), c = make_float2 (cx, cx * 0.999f);
  long long t0 = clock64 ();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int u = 0; u < 32; ++u) {
      if (MODE == 0) a = fma2 (a, b, c);
      if (MODE == 1) a = add2 (a, b);
      if (MODE == 2) a = mul2 (a, b);
      if (MODE == 3) { a.x = sfma (a.x, b.x, c.x); a.y = sfma (a.y, b.y, c.y); }      // two independent scalar chains
      if (MODE == 4) a.x = sfma (a.x, b.x, c.x);                                       // one scalar chain
    }
  }
  long long t1 = clock64 ();
  out[threadIdx.x] = a.x + a.y;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE> void runlat (const char *name, float *out, long long *cyc)
{
  klat<MODE><<<1, 32>>> (out, cyc, 1.0001f, 0.5f);
  klat<MODE><<<1, 32>>> (out, cyc, 1.0001f, 0.5f);
  cudaDeviceSynchronize ();
  long long h; cudaMemcpy (&h, cyc, sizeof (long long), cudaMemcpyDeviceToHost);
  printf ("%-44s %.2f cycles per dependent step\n", name, (double) h / (ITER * 32.0));
}

template <int MODE> void run (const char *name, int nsm, float *out, long long *cyc)
{
  for (int threads = 128; threads <= 1024; threads *= 2) {
    k<MODE><<<nsm, threads>>> (out, cyc, 1.0001f, 0.5f);
    k<MODE><<<nsm, threads>>> (out, cyc, 1.0001f, 0.5f);
    cudaDeviceSynchronize ();
    long long h[1024]; cudaMemcpy (h, cyc, nsm * sizeof (long long), cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < nsm; ++i) mean += h[i]; mean /= nsm;
    const double inst = (double) ITER * 4 * NACC * (threads / 32) * ((MODE == 3 || MODE == 4 || MODE == 6) ? 2 : 1);
    printf ("%-34s warps/SM %2d  warp-instr/cycle/SM %.3f  (fp32 lane-ops/cycle/SM %.1f)\n", name, threads / 32, inst / mean,
            inst / mean * 32 * ((MODE == 3 || MODE == 4 || MODE == 6) ? 1 : 2));
  }
}

int main ()
{
  cudaDeviceProp p; cudaGetDeviceProperties (&p, 0);
  const int nsm = p.multiProcessorCount;
  float *out; long long *cyc;
  cudaMalloc (&out, (size_t) nsm * 1024 * sizeof (float)); cudaMalloc (&cyc, nsm * sizeof (long long));
  printf ("%s, %d SMs\n", p.name, nsm);
  run<0> ("FFMA2 r,r,r", nsm, out, cyc);
  run<1> ("FADD2 r,r", nsm, out, cyc);
  run<2> ("FMUL2 r,r", nsm, out, cyc);
  run<3> ("FFMA r,r,r (scalar)", nsm, out, cyc);
  run<4> ("FADD r,r (scalar)", nsm, out, cyc);
  run<5> ("FFMA2 r,imm,r", nsm, out, cyc);
  run<6> ("FFMA r,imm,r (scalar)", nsm, out, cyc);
  run<7> ("FADD2 r, swap/neg r", nsm, out, cyc);
  printf ("dependent chains, one warp:\n");
  runlat<0> ("FFMA2 chain", out, cyc);
  runlat<1> ("FADD2 chain", out, cyc);
  runlat<2> ("FMUL2 chain", out, cyc);
  runlat<3> ("2 x scalar FFMA chains (x and y of a float2)", out, cyc);
  runlat<4> ("scalar FFMA chain", out, cyc);
  return 0;
}
