// Measured FP64 (DFMA) issue rate of one SM: ILP independent chains per thread, WARPS warps per SM.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/fp64_peak.cu -o tools/micro/fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

template <int ILP>
void run(int warps_per_sm, int sms, double ghz) {
  double* out;
  cudaMalloc(&out, 8);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  dfma_kernel<ILP><<<sms, warps_per_sm * 32>>>(out, 100, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  dfma_kernel<ILP><<<sms, warps_per_sm * 32>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fma_per_sm = (double)iters * ILP * warps_per_sm * 32;
  const double cycles = ms * 1e-3 * ghz * 1e9;
  printf("ILP %d warps/SM %2d: %.3f ms  %.1f DFMA/clk/SM  (%.1f TFLOP/s on %d SMs at %.3f GHz)\n", ILP, warps_per_sm, ms, fma_per_sm / cycles,
         2.0 * fma_per_sm * sms / (ms * 1e-3) / 1e12, sms, ghz);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  for (int w : {4, 8, 12, 16, 24, 32}) run<1>(w, p.multiProcessorCount, ghz);
  for (int w : {4, 8, 12, 16, 32}) run<4>(w, p.multiProcessorCount, ghz);
  for (int w : {4, 8, 16}) run<8>(w, p.multiProcessorCount, ghz);
  return 0;
}
