// Micro-benchmark (developer tool): issue rate of FFMA vs the packed FFMA2 (fma.rn.f32x2, sm_100+) and of a
// FFMA2 + FMNMX mix shaped like the BVH slab test.  Prints warp-instructions per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_ffma2 scripts/ubench_ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 r, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

template <int MODE>
__global__ void k(float *out, int iters, float s) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  u64 p[8];
  for (int i = 0; i < 8; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
  const u64 sp = pk(s, s * 1.0001f), cp = pk(0.5f, 0.25f);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fma1(a[i], s, 0.5f);          // 16 FFMA
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], sp, cp);             // 8 FFMA2 = 16 FMAs
    } else if (MODE == 2) {                                              // 8 FFMA2 + 8 FMNMX
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], sp, cp);
#pragma unroll
      for (int i = 0; i < 8; ++i) { float x, y; upk(p[i], x, y); x = fminf(x, y + 1.0f); p[i] = pk(x, y); }
    } else {                                                             // 16 FFMA + 8 FMNMX
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fma1(a[i], s, 0.5f);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[2 * i] = fminf(a[2 * i], a[2 * i + 1] + 1.0f);
    }
  }
  float r = 0;
  for (int i = 0; i < 16; ++i) r += a[i];
  for (int i = 0; i < 8; ++i) { float x, y; upk(p[i], x, y); r += x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  int sms = pr.multiProcessorCount, blocks = sms * 8, threads = 256, iters = 20000;
  float *out; cudaMalloc(&out, blocks * threads * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const char *names[4] = {"16 FFMA", "8 FFMA2", "8 FFMA2 + 8 FMNMX(+FADD)", "16 FFMA + 8 FMNMX(+FADD)"};
  const double per_iter[4] = {16, 8, 8 + 16, 16 + 16};   // warp instructions per iteration (FMNMX + its FADD)
  for (int m = 0; m < 4; ++m) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (m == 0) k<0><<<blocks, threads>>>(out, iters, 0.999f);
      if (m == 1) k<1><<<blocks, threads>>>(out, iters, 0.999f);
      if (m == 2) k<2><<<blocks, threads>>>(out, iters, 0.999f);
      if (m == 3) k<3><<<blocks, threads>>>(out, iters, 0.999f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warps = (double)blocks * threads / 32, winst = warps * iters * per_iter[m];
    double clocks = ms * 1e-3 * clk_khz * 1e3;
    printf("%-28s %8.3f ms  %.2f warp-inst/clk/SM (nominal clock %d MHz), %.2f FMA-lane-ops/clk/SM\n", names[m], ms, winst / clocks / sms, clk_khz / 1000,
           warps * iters * 16 * 32 / clocks / sms);
  }
  return 0;
}
