// Throughput probes for the pipes the pose kernels lean on (fp64 FMA, fp32 FMA, f32->f64 convert).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/microbench.cu -o tools/microbench
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(double* out, int iters, float seed) {
  double a[8];
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3 + i; f[i] = seed + threadIdx.x + i; }
  const double m = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = fma(a[i], m, c);                       // DFMA
      if (MODE == 1) f[i] = fmaf(f[i], 1.0000001f, 1e-9f);         // FFMA
      if (MODE == 2) { a[i] += (double)f[i]; f[i] += 1.0f; }       // F2F.F64.F32 + DADD + FADD
      if (MODE == 3) { a[i] += 1.0; f[i] += 1.0f; }                // DADD + FADD (baseline for mode 2)
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int sms, double clk_ghz) {
  const int threads = 512, blocks = sms * 2, iters = 4096;
  double* out;
  cudaMalloc(&out, sizeof(double) * threads * blocks);
  probe<MODE><<<blocks, threads>>>(out, 16, 1.0f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<MODE><<<blocks, threads>>>(out, iters, 1.0f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)threads * blocks * iters * 8;
  printf("%-28s %8.3f ms  %8.2f Gop/s  %6.2f lane-ops/clk/SM (at %.3f GHz)\n", name, ms, ops / ms / 1e6,
         ops / (ms * 1e-3) / sms / (clk_ghz * 1e9), clk_ghz);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double ghz = clk_khz / 1e6;
  printf("%s, %d SMs, %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
  run<0>("DFMA", p.multiProcessorCount, ghz);
  run<1>("FFMA", p.multiProcessorCount, ghz);
  run<2>("F2F.f64.f32 + DADD + FADD", p.multiProcessorCount, ghz);
  run<3>("DADD + FADD", p.multiProcessorCount, ghz);
  return 0;
}
