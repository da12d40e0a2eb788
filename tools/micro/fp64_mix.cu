// Does an FP64 instruction block the warp scheduler's issue port for both of its pipe cycles?  4 independent DFMA chains
// per thread, plus K independent 32-bit IMADs (or FFMAs) per DFMA.  If the extra instructions are free up to K = 1,
// they issue in the shadow of the half-rate DFMA; if time grows by 1 cycle per extra instruction, they do not.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/fp64_mix.cu -o tools/micro/fp64_mix
#include <cstdio>
#include <cuda_runtime.h>

template <int K, bool FLOAT>
__global__ void mix_kernel(double* out, int iters, double a, double b, int m, float fa) {
  double x[4];
  int y[4 * (K > 0 ? K : 1)];
  float z[4 * (K > 0 ? K : 1)];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < 4 * (K > 0 ? K : 1); ++i) { y[i] = threadIdx.x + i; z[i] = threadIdx.x * 0.5f + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x[i] = fma(x[i], a, b);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (FLOAT) z[i * K + k] = fmaf(z[i * K + k], fa, fa);
        else y[i * K + k] = y[i * K + k] * m + it;
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < 4 * (K > 0 ? K : 1); ++i) s += y[i] + z[i];
  if (s == 12345.678) out[0] = s;
}

template <int K, bool FLOAT>
void run(int warps_per_sm, int sms, double ghz) {
  double* out;
  cudaMalloc(&out, 8);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  mix_kernel<K, FLOAT><<<sms, warps_per_sm * 32>>>(out, 100, 1.0000001, 1e-9, 3, 1.0001f);
  cudaEventRecord(e0);
  mix_kernel<K, FLOAT><<<sms, warps_per_sm * 32>>>(out, iters, 1.0000001, 1e-9, 3, 1.0001f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double cycles = ms * 1e-3 * ghz * 1e9;
  const double dfma_warp_instr_per_smsp = (double)iters * 4 * warps_per_sm / 4;
  printf("%s x%d per DFMA, warps/SM %2d: %.3f ms  %.2f cycles per DFMA per scheduler\n", FLOAT ? "FFMA" : "IMAD", K, warps_per_sm, ms,
         cycles / dfma_warp_instr_per_smsp);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  const int sms = p.multiProcessorCount;
  for (int w : {8, 16}) {
    run<0, false>(w, sms, ghz);
    run<1, false>(w, sms, ghz);
    run<2, false>(w, sms, ghz);
    run<3, false>(w, sms, ghz);
    run<1, true>(w, sms, ghz);
    run<2, true>(w, sms, ghz);
  }
  return 0;
}
