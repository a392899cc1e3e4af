// Similarity-join support kernels around the tcgen05 screen GEMM (EpiJoinThreshold):
//   1. the screen runs on bf16 copies of the embeddings at (thr - margin) and emits candidate pairs;
//   2. rescore_pairs_kernel recomputes every candidate in fp32 from the original embeddings and keeps
//      s >= thr, which restores the reference's fp32 `np.dot(E, E.T) >= threshold` semantics
//      (fingerprint.py:493-499) - bf16 rounding can only add candidates, never lose a pair, as long as
//      margin >= 2^-8 * max|q| * max|d|.
#pragma once
#include "sm100_primitives.cuh"

namespace vfp {

// One warp per candidate: 256-d fp32 dot product, 8 elements per lane.
__global__ void __launch_bounds__(256)
rescore_pairs_kernel(const float* __restrict__ q, const float* __restrict__ db, int dim, long long q_row0,
                     const int* __restrict__ cand_i, const int* __restrict__ cand_j,
                     const unsigned long long* __restrict__ cand_count, long long cand_capacity, float thr,
                     int* __restrict__ out_i, int* __restrict__ out_j, float* __restrict__ out_s,
                     unsigned long long* __restrict__ out_count, long long capacity, int mirror) {
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  long long n = (long long)*cand_count;
  if (n > cand_capacity) n = cand_capacity;
  for (long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < n; c += warps) {
    const int gi = cand_i[c], gj = cand_j[c];
    const float* a = q + (size_t)(gi - q_row0) * dim;
    const float* b = db + (size_t)gj * dim;
    float s = 0.f;
    for (int k = lane * 4; k < dim; k += 128) {
      const float4 x = *reinterpret_cast<const float4*>(a + k);
      const float4 y = *reinterpret_cast<const float4*>(b + k);
      s += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && s >= thr) {
      // mirror (symmetric self join, candidates have j >= i): the pair below the diagonal has the same fp32 score
      // (the products commute term by term and are summed in the same order)
      const bool twin = mirror && (long long)gj != (long long)gi - q_row0;
      const unsigned long long slot = atomicAdd(out_count, twin ? 2ull : 1ull);
      if ((long long)slot < capacity) {
        out_i[slot] = gi;
        out_j[slot] = gj;
        out_s[slot] = s;
      }
      if (twin && (long long)slot + 1 < capacity) {
        out_i[slot + 1] = (int)(gj + q_row0);
        out_j[slot + 1] = (int)(gi - q_row0);
        out_s[slot + 1] = s;
      }
    }
  }
}

}  // namespace vfp
