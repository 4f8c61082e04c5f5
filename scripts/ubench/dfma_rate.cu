// DFMA / FFMA issue rate per SM (is fp64 a full-rate pipe on this part?). nvcc -arch=sm_100a -O3 dfma_rate.cu -o dfma_rate
#include <cstdio>
#include <cuda_runtime.h>
template <typename T>
__global__ void chains(T* out, int iters, T a, T b) {
  T x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = T(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = x[i] * a + b;
  }
  T s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename T>
static void run(const char* name, int sms, double clock_ghz) {
  T* out; cudaMalloc(&out, sizeof(T) * sms * 4 * 512);
  const int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  chains<T><<<sms * 4, 512>>>(out, iters, T(1.0000001), T(1e-9));
  cudaEventRecord(e0);
  chains<T><<<sms * 4, 512>>>(out, iters, T(1.0000001), T(1e-9));
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fma = double(sms) * 4 * 512 * iters * 8;
  printf("%s: %.3f ms, %.2f T FMA/s, %.1f FMA lanes/clk/SM at %.2f GHz\n", name, ms, fma / ms / 1e9, fma / (ms * 1e-3) / sms / (clock_ghz * 1e9), clock_ghz);
  cudaFree(out);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const double ghz = p.clockRate * 1e-6;
  printf("%s, %d SMs, %.2f GHz\n", p.name, p.multiProcessorCount, ghz);
  run<float>("fp32", p.multiProcessorCount, ghz);
  run<double>("fp64", p.multiProcessorCount, ghz);
  return 0;
}
