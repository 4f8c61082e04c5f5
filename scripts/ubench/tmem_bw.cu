// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM on sm_100a, as a function of the
// number of reading warps and the load width. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw.bin tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

template <int X>
__device__ __forceinline__ uint32_t ld(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ld<16>(uint32_t taddr) {
  uint32_t v[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s ^= v[i];
  return s;
}
template <>
__device__ __forceinline__ uint32_t ld<32>(uint32_t taddr) {
  uint32_t v[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s ^= v[i];
  return s;
}

// NOWAIT variant: issue two x16 loads back to back, one wait
__device__ __forceinline__ uint32_t ld2x16(uint32_t taddr) {
  uint32_t v[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr + 16) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s ^= v[i];
  return s;
}

template <int MODE>   // 16: x16+wait, 32: x32+wait, 2: 2 x16 then wait
__global__ void __launch_bounds__(1024, 1) tmem_read_kernel(int iters, unsigned long long* cycles, uint32_t* sink) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_ptr + (uint32_t((warp & 3) * 32) << 16);
  // warps beyond the first four share lane quadrants and read different column ranges
  const int grp = warp >> 2, n_grp = blockDim.x >> 7;
  const int cols_per_grp = 512 / n_grp;
  uint32_t acc = 0;
  __syncthreads();
  unsigned long long t0 = clock64();
  constexpr int W = MODE == 16 ? 16 : 32;
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int c = 0; c < cols_per_grp; c += W) {
      const uint32_t a = base + uint32_t(grp * cols_per_grp + c);
      if (MODE == 16) acc ^= ld<16>(a);
      else if (MODE == 32) acc ^= ld<32>(a);
      else acc ^= ld2x16(a);
    }
  }
  __syncthreads();
  unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_ptr), "r"(512) : "memory");
}

int main() {
  unsigned long long* d_cyc; uint32_t* d_sink;
  cudaMalloc(&d_cyc, 148 * 8); cudaMalloc(&d_sink, 4096);
  const int iters = 2000;
  int warps_list[] = {4, 8, 16, 32};
  for (int mode : {16, 32, 2}) {
    for (int w : warps_list) {
      for (int grid : {1, 148}) {
        auto launch = [&]() {
          if (mode == 16) tmem_read_kernel<16><<<grid, w * 32>>>(iters, d_cyc, d_sink);
          else if (mode == 32) tmem_read_kernel<32><<<grid, w * 32>>>(iters, d_cyc, d_sink);
          else tmem_read_kernel<2><<<grid, w * 32>>>(iters, d_cyc, d_sink);
        };
        launch(); launch();
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        unsigned long long cyc[148];
        cudaMemcpy(cyc, d_cyc, grid * 8, cudaMemcpyDeviceToHost);
        double mx = 0; for (int i = 0; i < grid; ++i) mx = cyc[i] > mx ? double(cyc[i]) : mx;
        // every iteration reads the whole 128 lanes x 512 columns x 4 B = 256 KB
        double bytes = double(iters) * 128.0 * 512.0 * 4.0;
        printf("mode=%s warps=%2d grid=%3d  cycles/256KB=%8.1f  bytes/cycle/SM=%7.1f\n",
               mode == 16 ? "x16" : mode == 32 ? "x32" : "2x16", w, grid, mx / iters, bytes / mx);
      }
    }
  }
  return 0;
}
