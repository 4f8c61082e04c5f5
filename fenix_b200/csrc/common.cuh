// common.cuh — shared device helpers for libfenix_knn (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fx {

constexpr uint64_t KEY_PAD = 0xffffffffffffffffull;  // sorts after every real (distance,row) key

// Monotone map float -> uint32: a < b  <=>  f2ord(a) < f2ord(b)  (NaN sorts above +inf).
__host__ __device__ __forceinline__ uint32_t f2ord(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  union { float f; uint32_t u; } cvt; cvt.f = f; uint32_t u = cvt.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } cvt; cvt.u = u; return cvt.f;
#endif
}

// (distance, local row) packed so that u64 order == (distance asc, row asc).
__device__ __forceinline__ uint64_t make_key(float dist, uint32_t row) {
  return (uint64_t(f2ord(dist)) << 32) | uint64_t(row);
}

// Reference-convention distance from fp64 partial sums (coder.py:38-50 of the reference).
//   qq = sum q_i^2, xx = sum x_i^2, qx = sum q_i x_i   (all accumulated in fp64)
// Rounded once to fp32 at the end; -0.0 is canonicalised to +0.0.
__device__ __forceinline__ float finish_distance(int metric, double qq, double xx, double qx) {
  double d;
  if (metric == 0) {                       // l2 / euclidean: torch.cdist -> sqrt(clamp_min(.,0))
    double d2 = (qq - 2.0 * qx) + xx;
    d = sqrt(d2 > 0.0 ? d2 : 0.0);
  } else if (metric == 1) {                // cosine: 0.5 - 0.5 * <q/max(|q|,eps), x/max(|x|,eps)>
    double nq = sqrt(qq), nx = sqrt(xx);
    nq = nq > 1e-12 ? nq : 1e-12;
    nx = nx > 1e-12 ? nx : 1e-12;
    d = 0.5 - 0.5 * (qx / (nq * nx));
  } else {                                 // dot / inner_product: -<q, x>
    d = -qx;
  }
  return float(d) + 0.0f;
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

}  // namespace fx
