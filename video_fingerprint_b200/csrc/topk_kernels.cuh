// Flat inner-product top-k (placeholder until the screen + exact re-score pipeline lands).
#pragma once
#include <string>

#include "gemm_launch.cuh"

namespace vfp {

inline size_t topk_workspace_bytes(int64_t n_q, int64_t n_db, int k) {
  (void)n_q; (void)n_db; (void)k;
  return 1024;
}

inline int topk_run(const float*, const float*, int64_t, int64_t, int, float, float*, int64_t*, unsigned long long*,
                    uint8_t*, cudaStream_t, std::string* err) {
  *err = "not implemented yet";
  return 1;
}

}  // namespace vfp
