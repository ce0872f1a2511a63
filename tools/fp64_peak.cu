// fp64_peak.cu -- measures the fp64 pipe of the device the roofline notes quote against:
//   dfma   : independent DFMA chains (vector fp64 pipe)
//   dmma   : mma.sync.m8n8k4.f64 chains (fp64 tensor path), counted in FMA-equivalents (256 per warp instruction)
//   mixed  : both in one instruction stream -- tells whether DMMA has issue/pipe capacity of its own
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_peak tools/fp64_peak.cu ; prints one JSON line
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NF, int NM>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b) {
  double f[NF > 0 ? NF : 1];
  double c[NM > 0 ? 2 * NM : 1];
#pragma unroll
  for (int i = 0; i < NF; i++) f[i] = threadIdx.x * 1e-9 + i;
#pragma unroll
  for (int i = 0; i < 2 * NM; i++) c[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < NF; i++) f[i] = fma(f[i], a, b);
#pragma unroll
      for (int i = 0; i < NM; i++) dmma(c[2 * i], c[2 * i + 1], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NF; i++) s += f[i];
#pragma unroll
  for (int i = 0; i < 2 * NM; i++) s += c[i];
  if (s == 12345.678) out[0] = s;
}

template <int NF, int NM>
double run(int nSM, int warpsPerSM, double *fmaRate, double *mmaRate) {
  double *d;
  cudaMalloc(&d, 8);
  const int iters = 20000;
  const int ctas = nSM * (warpsPerSM * 32 / 256);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    k<NF, NM><<<ctas, 256>>>(d, iters, 0.999999, 1e-7);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double threads = (double)ctas * 256;
  *fmaRate = threads * iters * 4.0 * NF / (best * 1e-3);                    // DFMA thread-ops / s
  *mmaRate = (threads / 32) * iters * 4.0 * NM * 256.0 / (best * 1e-3);     // FMA-equivalents / s
  cudaFree(d);
  return best;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double f, m;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, p.multiProcessorCount, clk);
  run<16, 0>(p.multiProcessorCount, 32, &f, &m);
  printf(", \"dfma_tflops\": %.3f", 2 * f / 1e12);
  const double fpeak = f;
  run<16, 0>(p.multiProcessorCount, 8, &f, &m);
  printf(", \"dfma_tflops_8warps\": %.3f", 2 * f / 1e12);
  run<0, 8>(p.multiProcessorCount, 32, &f, &m);
  printf(", \"dmma_tflops\": %.3f", 2 * m / 1e12);
  run<0, 8>(p.multiProcessorCount, 8, &f, &m);
  printf(", \"dmma_tflops_8warps\": %.3f", 2 * m / 1e12);
  run<8, 4>(p.multiProcessorCount, 32, &f, &m);
  printf(", \"mixed_8dfma_4dmma\": {\"dfma_tflops\": %.3f, \"dmma_tflops\": %.3f}", 2 * f / 1e12, 2 * m / 1e12);
  run<16, 2>(p.multiProcessorCount, 32, &f, &m);
  printf(", \"mixed_16dfma_2dmma\": {\"dfma_tflops\": %.3f, \"dmma_tflops\": %.3f}", 2 * f / 1e12, 2 * m / 1e12);
  printf(", \"dfma_per_clk_per_sm\": %.2f}\n", fpeak / p.multiProcessorCount / (clk * 1e3));
  return 0;
}
