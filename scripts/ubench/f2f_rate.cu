// fp32 -> fp64 conversion rate: F2F.F64.F32 against an integer-ALU widening of the bit pattern (exact for normal numbers and
// zero). nvcc -gencode arch=compute_100a,code=sm_100a -O3 f2f_rate.cu -o f2f_rate
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double widen_int(float f) {
  const unsigned b = __float_as_uint(f);
  const unsigned mag = b & 0x7fffffffu;
  unsigned hi = (mag >> 3) + 0x38000000u;
  if (mag == 0u) hi = 0u;
  return __hiloint2double(int(hi | (b & 0x80000000u)), int(b << 29));
}
template <int MODE>   // 0: FFMA only, 1: + F2F, 2: + integer widening, 3: half and half
__global__ void k(double* out, int iters, float a, float c) {
  float x[8];
  double s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = float(threadIdx.x + i) * 1e-3f + 1.f; s[i] = 0.0; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      x[i] = x[i] * a + c;
      if (MODE == 1 || (MODE == 3 && (i & 1))) s[i] += double(x[i]);
      else if (MODE == 2 || MODE == 3) s[i] += widen_int(x[i]);
      else s[i] += __hiloint2double(__float_as_int(x[i]), 0);   // (keeps the DADD, no conversion)
    }
  }
  double t = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int MODE>
static float run(int sms) {
  double* out; cudaMalloc(&out, 8ull * sms * 4 * 512);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms * 4, 512>>>(out, 2048, 0.99999f, 1e-5f);
  cudaEventRecord(e0);
  k<MODE><<<sms * 4, 512>>>(out, 2048, 0.99999f, 1e-5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaFree(out);
  return ms;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  const double n = double(sms) * 4 * 512 * 2048 * 8, ghz = p.clockRate * 1e-6;
  const float base = run<0>(sms), f2f = run<1>(sms), wi = run<2>(sms), mix = run<3>(sms);
  auto rate = [&](float ms) { return n / (ms * 1e-3) / sms / (ghz * 1e9); };
  printf("%s: FFMA+DADD only %.3f ms (%.1f el/clk/SM); + F2F.F64.F32 %.3f ms (%.1f); + integer widening %.3f ms (%.1f); half/half %.3f ms (%.1f)\n",
         p.name, base, rate(base), f2f, rate(f2f), wi, rate(wi), mix, rate(mix));
  return 0;
}
