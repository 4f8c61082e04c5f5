// direct_scan.cuh — the latency path: ONE launch answers a handful of queries over a small shard, exactly.
//
// A single query against a 100k x 128 shard (BASELINE C1) is 51 MB of fp32 rows that live in L2: the tensor-core
// pipeline (prep, sample prepass, threshold pick, filter, finish: six launches, ~55 us of device time for ~12 us of
// filter) is all fixed cost there. This kernel reads the fp32 rows themselves once (coalesced, eight lanes per row),
// accumulates q.x and |x|^2 in fp64 from exact fp32 products - the arithmetic of the finish kernel's rerank, in the
// same summation order, so the distances are bit-identical to the tensor-core path's - collects (distance, row) keys
// per CTA in shared memory, selects each CTA's k smallest with warp-level sorted lists held in registers (shuffle
// bitonic networks, a handful of barriers; block-wide shared-memory bitonic sorts measured 17 us here), and the last
// CTA to finish joins the per-CTA lists the same way and writes the result. No shadow, no certificate, no second launch. A small query block (<= 896 floats) travels in the kernel
// parameters and the results may be written straight to mapped pinned host memory, so fx_search issues no copies.
// (Reading the queries from mapped host memory instead was measured: every CTA fetches them over PCIe, ~30 us per query.) Replaces index.py:162-168 (distance column + select_k + take indices) like the rest
// of the library; HBM/L2-bound byte work, deliberately kept off the tensor cores.
#pragma once
#include "common.cuh"
#include "exact_scan.cuh"

namespace fx {

constexpr int DS_THREADS = 512;          // 16 warps: 64 rows in flight per CTA and load round
constexpr int DS_MAX_Q = 8;              // queries per launch
constexpr int DS_MAX_K = 128;
constexpr int DS_BUF = 1024;             // candidate keys per (CTA, query) between trims
constexpr int DS_ROWS_PER_STEP = (DS_THREADS / 8) * 2;   // 8 lanes per row, two rows per lane group and step
constexpr int DS_CHECK_STEPS = 4;        // steps between overflow checks: 4 * 128 = 512 appends at most, BUF - K >= 896
constexpr int DS_INLINE_FLOATS = 896;    // query floats carried in the kernel parameters (3.5 KB of the 4 KB parameter space)

struct DirectParams {
  const float* X; int64_t n_rows; int pitch; int dim; int64_t row_base;
  const float* Q;            // [n_q][dim] in device memory, or null: the queries are q_inline
  int n_q; int metric; int k;
  const uint8_t* mask;       // [n_rows] or null
  uint64_t* partial;         // [gridDim.x][n_q][k] sorted keys of every CTA
  unsigned int* ticket;      // zero on entry; the last CTA leaves it zero again
  int64_t* out_rows; float* out_dist;   // [n_q][k]; device memory or mapped pinned host memory
  float q_inline[DS_INLINE_FLOATS];
};

// ---- warp-level sorted lists: 32 * R keys per warp, key idx = r * 32 + lane held in v[r] of that lane ----
// One compare-exchange step of the bitonic network (partner idx ^ STRIDE, ascending where (idx & SIZE) == 0). Strides
// >= 32 pair registers of the same lane, smaller ones are shuffles; every loop bound is a compile-time constant, so the
// lists stay in registers.
template <int R, int SIZE, int STRIDE>
__device__ __forceinline__ void warp_cmpx(uint64_t (&v)[R], int lane) {
  if constexpr (STRIDE >= 32) {
    constexpr int RS = STRIDE / 32;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if ((r & RS) == 0) {
        const bool up = ((r * 32) & SIZE) == 0;
        const uint64_t a = v[r], b = v[r | RS];
        const uint64_t lo = a < b ? a : b, hi = a < b ? b : a;
        v[r] = up ? lo : hi; v[r | RS] = up ? hi : lo;
      }
    }
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const uint64_t o = __shfl_xor_sync(0xffffffffu, v[r], STRIDE);
      const bool up = ((r * 32 + lane) & SIZE) == 0;
      const bool keep_min = ((lane & STRIDE) == 0) == up;
      v[r] = (keep_min == (v[r] < o)) ? v[r] : o;
    }
  }
}
template <int R, int SIZE, int STRIDE>
__device__ __forceinline__ void warp_stage(uint64_t (&v)[R], int lane) {   // strides STRIDE, STRIDE / 2, ..., 1 of one stage
  warp_cmpx<R, SIZE, STRIDE>(v, lane);
  if constexpr (STRIDE > 1) warp_stage<R, SIZE, STRIDE / 2>(v, lane);
}
template <int R, int SIZE>
__device__ __forceinline__ void warp_sort_from(uint64_t (&v)[R], int lane) {   // stages SIZE, 2 SIZE, ..., 32 R
  warp_stage<R, SIZE, SIZE / 2>(v, lane);
  if constexpr (SIZE < 32 * R) warp_sort_from<R, SIZE * 2>(v, lane);
}
template <int R>
__device__ __forceinline__ void warp_sort(uint64_t (&v)[R], int lane) { warp_sort_from<R, 2>(v, lane); }
// best (sorted) <- the 32 R smallest of best U other (both sorted ascending): elementwise minimum against the reversed
// other list is bitonic, one merge stage sorts it
template <int R>
__device__ __forceinline__ void warp_merge_keep_low(uint64_t (&best)[R], const uint64_t (&other)[R], int lane) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const uint64_t o = __shfl_sync(0xffffffffu, other[R - 1 - r], 31 - lane);
    best[r] = best[r] < o ? best[r] : o;
  }
  warp_stage<R, 32 * R, 16 * R>(best, lane);
}

// The CTA's 32 R smallest keys of `cnt` keys, sorted, left in lists[0, 32 R). get(i) returns key i. Every warp sorts
// chunks of 32 R keys in registers and folds them into its running best list; a tree over the warps' lists (shared
// memory, log2(warps) barriers) joins them. All threads of the CTA call it.
template <int R, typename Get>
__device__ __forceinline__ void cta_select(Get get, int cnt, uint64_t* lists) {
  constexpr int L = 32 * R;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  uint64_t best[R];
#pragma unroll
  for (int r = 0; r < R; ++r) best[r] = KEY_PAD;
  for (int c0 = warp * L; c0 < cnt; c0 += n_warps * L) {
    uint64_t v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { const int i = c0 + r * 32 + lane; v[r] = i < cnt ? get(i) : KEY_PAD; }
    warp_sort<R>(v, lane);
    warp_merge_keep_low<R>(best, v, lane);
  }
#pragma unroll
  for (int r = 0; r < R; ++r) lists[warp * L + r * 32 + lane] = best[r];
  __syncthreads();
  for (int span = 1; span < n_warps; span <<= 1) {
    if ((warp & (2 * span - 1)) == 0 && warp + span < n_warps) {
      uint64_t v[R];
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = lists[(warp + span) * L + r * 32 + lane];
      warp_merge_keep_low<R>(best, v, lane);
#pragma unroll
      for (int r = 0; r < R; ++r) lists[warp * L + r * 32 + lane] = best[r];
    }
    __syncthreads();
  }
}

template <int NQ, int R>
__global__ void __launch_bounds__(DS_THREADS, 1)
knn_direct_kernel(DirectParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // double qs[NQ][pitch] | u64 buf[NQ][DS_BUF] | u64 lists[warps][32 R]
  double* qs = reinterpret_cast<double*>(smem_raw);
  uint64_t* buf = reinterpret_cast<uint64_t*>(qs + size_t(NQ) * p.pitch);
  uint64_t* lists = buf + size_t(NQ) * DS_BUF;
  __shared__ double s_qq[NQ];
  __shared__ uint64_t s_tau[NQ];
  __shared__ int s_cnt[NQ];
  __shared__ unsigned int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane & 7, grp = lane >> 3;

  for (int q = 0; q < NQ; ++q) {
    for (int d = tid; d < p.pitch; d += DS_THREADS) {
      float v = 0.f;
      if (q < p.n_q && d < p.dim) v = p.Q != nullptr ? p.Q[size_t(q) * p.dim + d] : p.q_inline[q * p.dim + d];
      qs[size_t(q) * p.pitch + d] = double(v);
    }
  }
  if (tid < NQ) { s_cnt[tid] = 0; s_tau[tid] = KEY_PAD; }
  __syncthreads();
  if (warp < NQ) {   // |q|^2, summed as the finish kernel sums it (lane-strided, xor tree)
    double s = 0.0;
    const double* qv = qs + size_t(warp) * p.pitch;
    for (int d = lane; d < p.pitch; d += 32) s = fma(qv[d], qv[d], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_qq[warp] = s;
  }
  __syncthreads();

  const int n4 = p.pitch >> 2;
  const int64_t n_steps = (p.n_rows + DS_ROWS_PER_STEP - 1) / DS_ROWS_PER_STEP;
  int since_check = 0;
  for (int64_t step = blockIdx.x; step < n_steps; step += gridDim.x) {
    // rows of this lane group: two of the warp's eight
    const int64_t r0 = step * DS_ROWS_PER_STEP + warp * 8 + grp, r1 = r0 + 4;
    bool live0 = r0 < p.n_rows, live1 = r1 < p.n_rows;
    if (p.mask != nullptr) { live0 = live0 && p.mask[r0] != 0; live1 = live1 && p.mask[r1] != 0; }
    const float4* x0p = reinterpret_cast<const float4*>(p.X + size_t(live0 ? r0 : 0) * p.pitch);
    const float4* x1p = reinterpret_cast<const float4*>(p.X + size_t(live1 ? r1 : 0) * p.pitch);
    double xx0 = 0.0, xx1 = 0.0, qx0[NQ], qx1[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) { qx0[q] = 0.0; qx1[q] = 0.0; }
    if (__any_sync(0xffffffffu, live0 || live1)) {
#pragma unroll 4
      for (int j = sub; j < n4; j += 8) {
        const float4 a = live0 ? __ldg(x0p + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 b = live1 ? __ldg(x1p + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        const double a0 = a.x, a1 = a.y, a2 = a.z, a3 = a.w, b0 = b.x, b1 = b.y, b2 = b.z, b3 = b.w;
        xx0 = fma(a0, a0, xx0); xx1 = fma(b0, b0, xx1);
        xx0 = fma(a1, a1, xx0); xx1 = fma(b1, b1, xx1);
        xx0 = fma(a2, a2, xx0); xx1 = fma(b2, b2, xx1);
        xx0 = fma(a3, a3, xx0); xx1 = fma(b3, b3, xx1);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const double2* qp = reinterpret_cast<const double2*>(qs + size_t(q) * p.pitch + 4 * j);
          const double2 u = qp[0], v = qp[1];
          qx0[q] = fma(a0, u.x, qx0[q]); qx1[q] = fma(b0, u.x, qx1[q]);
          qx0[q] = fma(a1, u.y, qx0[q]); qx1[q] = fma(b1, u.y, qx1[q]);
          qx0[q] = fma(a2, v.x, qx0[q]); qx1[q] = fma(b2, v.x, qx1[q]);
          qx0[q] = fma(a3, v.y, qx0[q]); qx1[q] = fma(b3, v.y, qx1[q]);
        }
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        xx0 += __shfl_xor_sync(0xffffffffu, xx0, o); xx1 += __shfl_xor_sync(0xffffffffu, xx1, o);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          qx0[q] += __shfl_xor_sync(0xffffffffu, qx0[q], o); qx1[q] += __shfl_xor_sync(0xffffffffu, qx1[q], o);
        }
      }
      // every lane of a group holds the group's sums: lane `sub` finishes query `sub` (all lanes busy at NQ = 8)
      if (sub < p.n_q) {
        double d0 = qx0[0], d1 = qx1[0];
#pragma unroll
        for (int q = 1; q < NQ; ++q) { if (sub == q) { d0 = qx0[q]; d1 = qx1[q]; } }
        const uint64_t tau = s_tau[sub];
        const double qq = s_qq[sub];
        if (live0) {
          const uint64_t key = make_key(finish_distance(p.metric, qq, xx0, d0), uint32_t(r0));
          if (key < tau) buf[size_t(sub) * DS_BUF + atomicAdd(&s_cnt[sub], 1)] = key;
        }
        if (live1) {
          const uint64_t key = make_key(finish_distance(p.metric, qq, xx1, d1), uint32_t(r1));
          if (key < tau) buf[size_t(sub) * DS_BUF + atomicAdd(&s_cnt[sub], 1)] = key;
        }
      }
    }
    if (++since_check < DS_CHECK_STEPS) continue;
    since_check = 0;
    // a buffer that the next DS_CHECK_STEPS steps could overflow is cut back to its 32 R (>= k) best, whose k-th key
    // becomes the admission threshold (a CTA sees n / gridDim rows: C1 never gets here)
    __syncthreads();
    unsigned need = 0;
    for (int q = 0; q < p.n_q; ++q) need |= (s_cnt[q] > DS_BUF - DS_CHECK_STEPS * DS_ROWS_PER_STEP) ? (1u << q) : 0u;
    __syncthreads();
    for (int q = 0; q < p.n_q; ++q) {
      if (!(need >> q & 1u)) continue;
      const int c = s_cnt[q];
      uint64_t* b = buf + size_t(q) * DS_BUF;
      cta_select<R>([b](int i) { return b[i]; }, c, lists);
      for (int i = tid; i < p.k; i += DS_THREADS) b[i] = lists[i];
      if (tid == 0) { s_cnt[q] = min(c, p.k); s_tau[q] = (c >= p.k) ? lists[p.k - 1] : KEY_PAD; }
      __syncthreads();
    }
  }
  __syncthreads();

  // ---- this CTA's k best per query, published ----
  for (int q = 0; q < p.n_q; ++q) {
    const uint64_t* b = buf + size_t(q) * DS_BUF;
    cta_select<R>([b](int i) { return b[i]; }, s_cnt[q], lists);
    for (int i = tid; i < p.k; i += DS_THREADS) p.partial[(size_t(blockIdx.x) * p.n_q + q) * p.k + i] = lists[i];
    __syncthreads();
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last == 0u) return;
  __threadfence();

  // ---- the last CTA joins gridDim.x lists of k keys per query ----
  const int total = int(gridDim.x) * p.k;
  for (int q = 0; q < p.n_q; ++q) {
    const uint64_t* src = p.partial;
    const int n_q = p.n_q, k = p.k;
    cta_select<R>([src, n_q, k, q](int i) { const int cta = i / k, e = i - cta * k; return __ldcg(src + (size_t(cta) * n_q + q) * k + e); },
                  total, lists);
    for (int i = tid; i < p.k; i += DS_THREADS) {
      const uint64_t key = lists[i];
      const bool pad = key == KEY_PAD;
      p.out_rows[size_t(q) * p.k + i] = pad ? int64_t(-1) : p.row_base + int64_t(key & 0xffffffffull);
      p.out_dist[size_t(q) * p.k + i] = pad ? __int_as_float(0x7f800000) : ord2f(uint32_t(key >> 32));
    }
    __syncthreads();
  }
  if (tid == 0) *p.ticket = 0u;
  __threadfence_system();   // results may live in mapped host memory
}

// ---- host side ----
struct DirectPlan { bool ok = false; int nq_t = 0; int r = 0; int grid = 0; size_t smem = 0; size_t partial_bytes = 0; };

// Shapes the kernel takes: a handful of queries, k <= 128, queries and candidate buffers within shared memory.
inline DirectPlan direct_plan(int64_t n_rows, int pitch, int64_t n_q, int k, int sm_count) {
  DirectPlan pl;
  if (n_rows < 1 || n_q < 1 || n_q > DS_MAX_Q || k < 1 || k > DS_MAX_K) return pl;
  pl.nq_t = n_q <= 1 ? 1 : n_q <= 2 ? 2 : n_q <= 4 ? 4 : 8;
  pl.r = k <= 32 ? 1 : k <= 64 ? 2 : 4;
  const int64_t n_steps = (n_rows + DS_ROWS_PER_STEP - 1) / DS_ROWS_PER_STEP;
  pl.grid = int(std::max<int64_t>(1, std::min<int64_t>(sm_count, n_steps)));
  pl.smem = size_t(pl.nq_t) * pitch * 8 + size_t(pl.nq_t) * DS_BUF * 8 + size_t(DS_THREADS / 32) * 32 * pl.r * 8;
  if (pl.smem > size_t(200) * 1024) return pl;
  pl.partial_bytes = size_t(pl.grid) * n_q * k * 8;
  pl.ok = true;
  return pl;
}

template <int NQ, int R>
inline cudaError_t direct_attr() {
  return cudaFuncSetAttribute(knn_direct_kernel<NQ, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}
template <int NQ>
inline cudaError_t direct_attr_q() {
  cudaError_t e = direct_attr<NQ, 1>();
  if (e == cudaSuccess) e = direct_attr<NQ, 2>();
  if (e == cudaSuccess) e = direct_attr<NQ, 4>();
  return e;
}
inline cudaError_t direct_set_attributes() {
  cudaError_t e = direct_attr_q<1>();
  if (e == cudaSuccess) e = direct_attr_q<2>();
  if (e == cudaSuccess) e = direct_attr_q<4>();
  if (e == cudaSuccess) e = direct_attr_q<8>();
  return e;
}

template <int NQ>
inline void direct_launch_q(const DirectPlan& pl, const DirectParams& p, cudaStream_t stream) {
  switch (pl.r) {
    case 1: knn_direct_kernel<NQ, 1><<<pl.grid, DS_THREADS, pl.smem, stream>>>(p); break;
    case 2: knn_direct_kernel<NQ, 2><<<pl.grid, DS_THREADS, pl.smem, stream>>>(p); break;
    default: knn_direct_kernel<NQ, 4><<<pl.grid, DS_THREADS, pl.smem, stream>>>(p); break;
  }
}
inline cudaError_t direct_launch(const DirectPlan& pl, const DirectParams& p, cudaStream_t stream) {
  switch (pl.nq_t) {
    case 1: direct_launch_q<1>(pl, p, stream); break;
    case 2: direct_launch_q<2>(pl, p, stream); break;
    case 4: direct_launch_q<4>(pl, p, stream); break;
    default: direct_launch_q<8>(pl, p, stream); break;
  }
  return cudaGetLastError();
}

}  // namespace fx
