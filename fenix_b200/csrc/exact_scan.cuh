// exact_scan.cuh — fp64-accumulating CUDA-core kernels: the certificate-free exact path.
//
// Used (a) as the checker / fallback of the tcgen05 filter path, (b) for masked (filtered)
// searches, (c) for the full distance column (reference index.py:162 with maxval=None).
// Every distance is computed as the reference defines it (coder.py:38-50) from fp64 sums of
// exact fp32 products and rounded once to fp32, then ordered by (distance, row).
#pragma once
#include "common.cuh"

namespace fx {

constexpr int SCAN_THREADS = 256;      // one corpus row per thread per tile
constexpr int SCAN_MAX_QB = 8;         // queries sharing one pass over the rows

// Block-wide bitonic sort (ascending) of n (power of two) u64 keys in shared memory.
__device__ __forceinline__ void block_bitonic_sort(uint64_t* keys, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
        int pos = 2 * i - (i & (stride - 1));
        int j = pos + stride;
        bool up = (pos & size) == 0;
        uint64_t a = keys[pos], b = keys[j];
        if ((a > b) == up) { keys[pos] = b; keys[j] = a; }
      }
      __syncthreads();
    }
  }
}

// Dynamic shared memory layout (host computes the same):
//   double  qs[QB][dimp]      queries widened to fp64 (dimp = pitch)
//   double  qq[QB]            squared query norms
//   u64     tau[QB]           current admission key per query
//   u64     floor[QB]         keys at or below it are not admitted (large-k searches run in passes of <= 2048 neighbours:
//                             pass p takes the smallest keys ABOVE the last key of pass p - 1); 0 = no floor
//   u64     buf[QB][BUF]      candidate keys
//   int     cnt[QB]
struct ScanParams {
  const float* X;          // [n_rows][pitch]
  int64_t n_rows;
  int pitch;               // floats per row, multiple of 4, pad columns are zero
  int dim;
  const float* Q;          // device [n_q][dim]
  int n_q;
  int metric;
  const uint8_t* mask;     // device [n_rows] or null
  const int* q_list;       // optional: indices of the queries to process (fallback subset)
  int n_list;              // number of entries of q_list (or n_q when q_list == null)
  int k;                   // 0 => distance-column mode
  int buf;                 // BUF: power of two >= k + SCAN_THREADS
  int qb;                  // queries per block (<= SCAN_MAX_QB)
  const uint64_t* floor;   // optional [n_list]: admission floor per list slot (multi-pass large-k search)
  uint64_t* partial;       // [n_list][gridDim.x][k] sorted keys per (query, row block)
  float* dist_out;         // [n_list][n_rows] (distance-column mode)
};

__global__ void __launch_bounds__(SCAN_THREADS)
exact_scan_kernel(ScanParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int QB = p.qb;
  double* qs = reinterpret_cast<double*>(smem_raw);
  double* qq = qs + size_t(QB) * p.pitch;
  uint64_t* tau = reinterpret_cast<uint64_t*>(qq + QB);
  uint64_t* flo = tau + QB;
  uint64_t* buf = flo + QB;
  int* cnt = reinterpret_cast<int*>(buf + size_t(QB) * p.buf);

  const int g0 = blockIdx.y * QB;                       // first list slot of this group
  const int nq_here = min(QB, p.n_list - g0);

  // stage queries (fp64) and their squared norms
  for (int i = threadIdx.x; i < QB * p.pitch; i += blockDim.x) {
    int q = i / p.pitch, d = i - q * p.pitch;
    double v = 0.0;
    if (q < nq_here && d < p.dim) {
      int qi = p.q_list ? p.q_list[g0 + q] : (g0 + q);
      v = double(p.Q[size_t(qi) * p.dim + d]);
    }
    qs[i] = v;
  }
  if (threadIdx.x < QB) {
    cnt[threadIdx.x] = 0; tau[threadIdx.x] = KEY_PAD;
    flo[threadIdx.x] = (p.floor != nullptr && int(threadIdx.x) < nq_here) ? p.floor[g0 + threadIdx.x] : 0ull;
  }
  __syncthreads();
  if (threadIdx.x < QB) {
    double s = 0.0;
    const double* qv = qs + size_t(threadIdx.x) * p.pitch;
    for (int d = 0; d < p.pitch; ++d) s = fma(qv[d], qv[d], s);
    qq[threadIdx.x] = s;
  }
  __syncthreads();

  const int64_t n_tiles = (p.n_rows + SCAN_THREADS - 1) / SCAN_THREADS;
  const int vec = p.pitch >> 2;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r = tile * SCAN_THREADS + threadIdx.x;
    bool live = r < p.n_rows;
    if (live && p.mask) live = p.mask[r] != 0;
    if (live || (p.k == 0 && r < p.n_rows)) {
      double acc[SCAN_MAX_QB];
#pragma unroll
      for (int q = 0; q < SCAN_MAX_QB; ++q) acc[q] = 0.0;
      double xx = 0.0;
      const float4* xp = reinterpret_cast<const float4*>(p.X + size_t(r) * p.pitch);
      for (int j = 0; j < vec; ++j) {
        float4 xv = __ldg(xp + j);
        double x0 = xv.x, x1 = xv.y, x2 = xv.z, x3 = xv.w;
        xx = fma(x0, x0, xx); xx = fma(x1, x1, xx); xx = fma(x2, x2, xx); xx = fma(x3, x3, xx);
#pragma unroll
        for (int q = 0; q < SCAN_MAX_QB; ++q) {
          if (q < QB) {
            const double2* qp = reinterpret_cast<const double2*>(qs + size_t(q) * p.pitch + 4 * j);
            double2 a = qp[0], b = qp[1];
            acc[q] = fma(x0, a.x, acc[q]); acc[q] = fma(x1, a.y, acc[q]);
            acc[q] = fma(x2, b.x, acc[q]); acc[q] = fma(x3, b.y, acc[q]);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < SCAN_MAX_QB; ++q) {
        if (q < nq_here) {
          float d = finish_distance(p.metric, qq[q], xx, acc[q]);
          if (p.k == 0) {
            p.dist_out[size_t(g0 + q) * p.n_rows + r] = d;
          } else {
            uint64_t key = make_key(d, uint32_t(r));
            if (key < tau[q] && key > flo[q]) {
              int pos = atomicAdd(&cnt[q], 1);
              buf[size_t(q) * p.buf + pos] = key;
            }
          }
        }
      }
    }
    if (p.k == 0) continue;
    __syncthreads();
    // every thread snapshots which buffers could overflow during the next tile, then a second
    // barrier keeps early finishers from appending while others still read the counters
    unsigned need = 0;
    for (int q = 0; q < nq_here; ++q) need |= (cnt[q] > p.buf - SCAN_THREADS) ? (1u << q) : 0u;
    __syncthreads();
    for (int q = 0; q < nq_here; ++q) {
      if (!(need >> q & 1u)) continue;
      int c = cnt[q];
      uint64_t* b = buf + size_t(q) * p.buf;
      for (int i = c + threadIdx.x; i < p.buf; i += blockDim.x) b[i] = KEY_PAD;
      __syncthreads();
      block_bitonic_sort(b, p.buf);
      if (threadIdx.x == 0) {
        cnt[q] = min(c, p.k);
        tau[q] = (c >= p.k) ? b[p.k - 1] : KEY_PAD;
      }
      __syncthreads();
    }
  }
  if (p.k == 0) return;

  __syncthreads();
  for (int q = 0; q < nq_here; ++q) {
    int c = cnt[q];
    uint64_t* b = buf + size_t(q) * p.buf;
    for (int i = c + threadIdx.x; i < p.buf; i += blockDim.x) b[i] = KEY_PAD;
    __syncthreads();
    block_bitonic_sort(b, p.buf);
    uint64_t* out = p.partial + (size_t(g0 + q) * gridDim.x + blockIdx.x) * p.k;
    for (int i = threadIdx.x; i < p.k; i += blockDim.x) out[i] = b[i];
    __syncthreads();
  }
}

// One block per query: merge W sorted key lists of length k into rows/dist outputs.
// q_list (optional) maps list slot -> output query index.
__global__ void __launch_bounds__(256)
merge_keys_kernel(const uint64_t* __restrict__ lists, int W, int k, int n_sort,
                  const int* __restrict__ q_list, int64_t row_base,
                  int64_t* __restrict__ out_rows, float* __restrict__ out_dist,
                  int out_stride, int out_off, uint64_t* __restrict__ floor_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
  const int slot = blockIdx.x;
  const int total = W * k;
  const uint64_t* src = lists + size_t(slot) * total;
  for (int i = threadIdx.x; i < n_sort; i += blockDim.x) keys[i] = (i < total) ? src[i] : KEY_PAD;
  __syncthreads();
  block_bitonic_sort(keys, n_sort);
  const int qi = q_list ? q_list[slot] : slot;
  // out_stride / out_off: the k results land in columns [out_off, out_off + k) of rows of out_stride entries
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    uint64_t key = keys[i];
    bool pad = key == KEY_PAD;
    out_rows[size_t(qi) * out_stride + out_off + i] = pad ? int64_t(-1) : row_base + int64_t(key & 0xffffffffull);
    out_dist[size_t(qi) * out_stride + out_off + i] = pad ? __int_as_float(0x7f800000) : ord2f(uint32_t(key >> 32));
  }
  if (floor_out != nullptr && threadIdx.x == 0) floor_out[slot] = keys[k - 1];   // next pass starts above this key
}

// Merge n_lists per-shard (row:int64, dist:f32) result lists into the global top-k.
// One block per query; bitonic sort on (ord(dist), row) pairs held in shared memory.
__global__ void __launch_bounds__(256)
merge_pairs_kernel(const int64_t* __restrict__ rows, const float* __restrict__ dist,
                   int n_lists, int64_t n_q, int k, int n_sort, int64_t rows_stride_bytes, int64_t dist_stride_bytes,
                   int64_t* __restrict__ out_rows, float* __restrict__ out_dist) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int64_t* srow = reinterpret_cast<int64_t*>(smem_raw);
  uint32_t* sord = reinterpret_cast<uint32_t*>(srow + n_sort);
  const int64_t q = blockIdx.x;
  const int total = n_lists * k;
  for (int i = threadIdx.x; i < n_sort; i += blockDim.x) {
    if (i < total) {
      int l = i / k, j = i - l * k;
      // list l: rows at `rows` + l * rows_stride bytes, distances at `dist` + l * dist_stride bytes, each [n_q][k]
      // (list-major arrays: strides n_q*k*8 and n_q*k*4; exchange slots of a sharded search: both = the slot size)
      const size_t off = size_t(q) * k + j;
      const int64_t r = reinterpret_cast<const int64_t*>(reinterpret_cast<const char*>(rows) + size_t(l) * rows_stride_bytes)[off];
      const float d = reinterpret_cast<const float*>(reinterpret_cast<const char*>(dist) + size_t(l) * dist_stride_bytes)[off];
      srow[i] = r < 0 ? INT64_MAX : r;
      sord[i] = r < 0 ? 0xffffffffu : f2ord(d);
    } else {
      srow[i] = INT64_MAX; sord[i] = 0xffffffffu;
    }
  }
  __syncthreads();
  for (int size = 2; size <= n_sort; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < (n_sort >> 1); i += blockDim.x) {
        int pos = 2 * i - (i & (stride - 1));
        int j = pos + stride;
        bool up = (pos & size) == 0;
        uint32_t ao = sord[pos], bo = sord[j];
        int64_t ar = srow[pos], br = srow[j];
        bool gt = (ao > bo) || (ao == bo && ar > br);
        if (gt == up) { sord[pos] = bo; sord[j] = ao; srow[pos] = br; srow[j] = ar; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    int64_t r = srow[i];
    bool pad = r == INT64_MAX;
    out_rows[q * k + i] = pad ? int64_t(-1) : r;
    out_dist[q * k + i] = pad ? __int_as_float(0x7f800000) : ord2f(sord[i]);
  }
}

// The same merge without a shared-memory limit (lists * k > 8192): every list is sorted by (distance, row) and keys are
// unique, so an entry's global rank is its own position plus, for every other list, the number of entries below it
// (binary search). One thread per entry; entries ranked < k are written straight to their output slot.
__global__ void __launch_bounds__(256)
merge_rank_kernel(const int64_t* __restrict__ rows, const float* __restrict__ dist,
                  int n_lists, int64_t n_q, int k, int64_t rows_stride_bytes, int64_t dist_stride_bytes,
                  int64_t* __restrict__ out_rows, float* __restrict__ out_dist) {
  const int64_t q = blockIdx.x;
  auto list_rows = [&](int l) { return reinterpret_cast<const int64_t*>(reinterpret_cast<const char*>(rows) + size_t(l) * rows_stride_bytes) + size_t(q) * k; };
  auto list_dist = [&](int l) { return reinterpret_cast<const float*>(reinterpret_cast<const char*>(dist) + size_t(l) * dist_stride_bytes) + size_t(q) * k; };
  auto key_of = [&](int l, int j, uint32_t& o, int64_t& r) {
    r = list_rows(l)[j];
    if (r < 0) { r = INT64_MAX; o = 0xffffffffu; } else o = f2ord(list_dist(l)[j]);
  };
  const int total = n_lists * k;
  for (int i = threadIdx.x; i < k; i += blockDim.x) { out_rows[q * k + i] = -1; out_dist[q * k + i] = __int_as_float(0x7f800000); }
  __syncthreads();
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int l = i / k, j = i - l * k;
    uint32_t o; int64_t r; key_of(l, j, o, r);
    if (r == INT64_MAX) continue;   // pad
    int rank = j;
    for (int m = 0; m < n_lists && rank < k; ++m) {
      if (m == l) continue;
      int lo = 0, hi = k;            // entries of list m strictly below (o, r)
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        uint32_t om; int64_t rm; key_of(m, mid, om, rm);
        if (om < o || (om == o && rm < r)) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    if (rank < k) { out_rows[q * k + rank] = r; out_dist[q * k + rank] = ord2f(o); }
  }
}

// Per-row cached terms for the tensor-core filter epilogue (one warp per row):
//   hx[r] = -0.5 * |x_r|^2   (L2 score  s = <q,x> - 0.5|x|^2,  d^2 = |q|^2 - 2 s)
//   rx[r] = 1 / max(|x_r|, 1e-12)            (cosine score s = <q,x> * rx)
// and the largest |x_r|^2 of the shard (error-bound constant), via atomicMax on its bits.
__global__ void __launch_bounds__(256)
row_norms_kernel(const float* __restrict__ X, int64_t n_rows, int pitch,
                 float* __restrict__ hx, float* __restrict__ rx, unsigned int* __restrict__ max_n2_bits,
                 double* __restrict__ norm_sum) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  float local_max = 0.f;
  double local_sum = 0.0;   // sum of |x| over this warp's rows (mean norm of the shard: how far the norms spread)
  for (int64_t r = int64_t(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); r < n_rows;
       r += int64_t(gridDim.x) * warps_per_block) {
    const float4* xp = reinterpret_cast<const float4*>(X + size_t(r) * pitch);
    double s = 0.0;
    for (int j = lane; j < (pitch >> 2); j += 32) {
      float4 v = __ldg(xp + j);
      s = fma(double(v.x), double(v.x), s); s = fma(double(v.y), double(v.y), s);
      s = fma(double(v.z), double(v.z), s); s = fma(double(v.w), double(v.w), s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      hx[r] = float(-0.5 * s);
      double n = sqrt(s);
      rx[r] = float(1.0 / (n > 1e-12 ? n : 1e-12));
      // round |x|^2 UP to fp32 so the bound constant never under-estimates
      float n2 = __double2float_ru(s);
      local_max = fmaxf(local_max, n2);
      local_sum += n;
    }
  }
  if (lane == 0 && local_max > 0.f) atomicMax(max_n2_bits, __float_as_uint(local_max));
  if (lane == 0 && local_sum > 0.0) atomicAdd(norm_sum, local_sum);
}

// Result slots of an empty shard: (row = -1, distance = +inf).
__global__ void fill_pad_kernel(int64_t* __restrict__ rows, float* __restrict__ dist, int64_t n) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    rows[i] = -1;
    dist[i] = __int_as_float(0x7f800000);
  }
}

}  // namespace fx
