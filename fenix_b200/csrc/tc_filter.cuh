// tc_filter.cuh — tcgen05/TMEM TF32 filter + fp64 rerank path (host glue + kernels).
// STUB for the first milestone: the tensor-core path reports "unsupported" so every search
// runs the exact scan kernel. Replaced by the real kernels next.
#pragma once
#include <string>
#include "common.cuh"

namespace fx {

struct TcState { int sm_count = 0; };
struct TcCorpus { int dummy = 0; };
struct TcSearch {
  const float* X; const float* hx; const float* rx; int64_t n_rows; int dim; int pitch; int64_t row_base;
  float max_norm; const float* Q; int n_q; int metric; int k; bool certify;
  int64_t* out_rows; float* out_dist; cudaStream_t stream; cudaEvent_t ev_k0, ev_k1;
};

inline bool tc_init(TcState* st, int sm_count, std::string*) { st->sm_count = sm_count; return true; }
inline bool tc_bind_corpus(TcState*, TcCorpus*, const float*, int64_t, int, int, std::string*) { return true; }
inline bool tc_supported(const TcState*, const TcCorpus*, int64_t, int, int, int) { return false; }
inline size_t tc_scratch_bytes(const TcState*, const TcSearch&) { return 0; }
inline bool tc_search(TcState*, TcCorpus*, const TcSearch&, void*, int*, std::string* err) { *err = "tc path not built"; return false; }
inline const int* tc_flags(const TcState*, const TcSearch&, void*) { return nullptr; }

}  // namespace fx
