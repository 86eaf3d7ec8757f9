// Micro-benchmark: FP64 vector (DFMA) vs FP64 tensor (DMMA m8n8k4) issue rates on sm_100a, alone and interleaved.
// Decides whether the sum-factorised contractions can be moved onto DMMA without competing with the
// per-point physics for the FP64 pipe.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NF, int NM>  // NF independent DFMA chains, NM independent DMMA chains per thread
__global__ void __launch_bounds__(256) k(double *out, int iters, double seed) {
  double f[NF > 0 ? NF : 1], c0[NM > 0 ? NM : 1], c1[NM > 0 ? NM : 1];
  const double a = seed + threadIdx.x * 1e-9, b = 0.999999;
#pragma unroll
  for (int i = 0; i < NF; i++) f[i] = a + i;
#pragma unroll
  for (int i = 0; i < NM; i++) c0[i] = a + i, c1[i] = a - i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < NM; i++) dmma(c0[i], c1[i], a, b);
#pragma unroll
      for (int i = 0; i < NF; i++) f[i] = fma(f[i], b, a);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NF; i++) s += f[i];
#pragma unroll
  for (int i = 0; i < NM; i++) s += c0[i] + c1[i];
  if (s == 1.2345) out[0] = s;
}

template <int NF, int NM>
void run(const char *name, int warps_per_sm) {
  int dev; cudaGetDevice(&dev);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  double *out; cudaMalloc(&out, 8);
  const int iters = 20000;
  const int blocks = p.multiProcessorCount * warps_per_sm / 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NF, NM><<<blocks, 256>>>(out, 100, 1.0);
  cudaEventRecord(e0);
  k<NF, NM><<<blocks, 256>>>(out, iters, 1.0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warps = double(blocks) * 8;
  const double nfma = warps * iters * 4.0 * NF, nmma = warps * iters * 4.0 * NM;
  const double tf = (nfma * 64 + nmma * 512) / (ms * 1e-3) / 1e12;
  printf("%-28s warps/SM %2d  %8.3f ms  DFMA %6.2f TF  DMMA %6.2f TF  total %6.2f TF  (warp-inst/clk/SMSP: dfma %.3f dmma %.3f @1.965GHz)\n",
         name, warps_per_sm, ms, nfma * 64 / (ms * 1e-3) / 1e12, nmma * 512 / (ms * 1e-3) / 1e12, tf,
         nfma / (ms * 1e-3) / (p.multiProcessorCount * 4 * 1.965e9), nmma / (ms * 1e-3) / (p.multiProcessorCount * 4 * 1.965e9));
  cudaFree(out);
}

int main() {
  for (int w : {8, 16, 32}) {
    run<8, 0>("DFMA x8 chains", w);
    run<0, 4>("DMMA x4 chains", w);
    run<0, 8>("DMMA x8 chains", w);
    run<8, 1>("DFMA x8 + DMMA x1", w);
    run<8, 2>("DFMA x8 + DMMA x2", w);
    run<4, 4>("DFMA x4 + DMMA x4", w);
  }
  run<1, 0>("DFMA latency (1 chain, 1 warp/SMSP)", 4);
  run<0, 1>("DMMA latency (1 chain, 1 warp/SMSP)", 4);
  return 0;
}
