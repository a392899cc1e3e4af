// Evaluation metrics over an embedding set (the trainer's second use of the N x N similarity matrix,
// /root/reference/train.py:285-358 compute_discrimination_metrics and :439-481 _compute_retrieval_metrics): the reference
// materialises S = E E^T on the host and then walks it row by row in Python (argpartition / argsort per row, boolean masks over
// all N^2 entries, sklearn's roc_auc_score over the N^2 scores). Here S is never stored: one pass of an exact fp32 tiled
// product visits every ordered pair (i, j), i != j, once and folds it into
//   * the rank of every positive (same video id) of row i: how many s_ij beat it, and how many tie with a smaller index
//     (-> R@k and mAP),
//   * intra / inter sums, sums of squares and counts (-> means, standard deviations, separation gap),
//   * intra / inter counts above each threshold (-> precision / recall / F1 / FPR),
//   * for every inter-video score its position among the SORTED intra-video scores (two binary searches)
//     (-> the Mann-Whitney form of AUC-ROC, ties counted one half like sklearn does).
// fp32 on the CUDA cores, not the bf16 tensor-core screen of the join: ranks and ties need every score exact, and a
// validation set is small (N ~ 10^4: 2.6e10 FMA). Every score, here and in pair_scores_kernel, is the same sequential
// fmaf chain over k = 0 .. dim-1, so a positive's score compares bit-exactly against the row it came from.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vfp {

constexpr int kMsTile = 64;     // 64 x 64 scores per CTA, 4 x 4 per thread
constexpr int kMsK = 16;        // K chunk
constexpr int kMsMaxThr = 8;

struct PairStatsParams {
  const float* e;            // [n][dim]
  const int* ids;            // [n] video id per row
  int n, dim;
  const int* row_ptr;        // [n+1] CSR: positives of row i are entries row_ptr[i] .. row_ptr[i+1]-1
  const int* pos_idx;        // [m] column index p of the positive
  const float* pos_score;    // [m] s_ip (pair_scores_kernel)
  const float* sorted_intra; // [m] all positive scores, ascending
  int m;
  float thr[kMsMaxThr];
  int n_thr;
  unsigned int* rank_greater;     // [m] += #{j != i : s_ij > s_ip}
  unsigned int* rank_tie_before;  // [m] += #{j != i, j < p : s_ij == s_ip}
  double* sums;                   // [4] intra sum, intra sum of squares, inter sum, inter sum of squares
  unsigned long long* counts;     // [4 + 2*kMsMaxThr]: n_intra, n_inter, sum_upper, sum_lower, intra_ge[8], inter_ge[8]
};

// s = <e_i, e_j>, the reference fmaf chain
__global__ void pair_scores_kernel(const float* __restrict__ e, int dim, const int* __restrict__ pi, const int* __restrict__ pj,
                                   int m, float* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  const float* a = e + (size_t)pi[t] * dim;
  const float* b = e + (size_t)pj[t] * dim;
  float acc = 0.0f;
  for (int k = 0; k < dim; ++k) acc = fmaf(a[k], b[k], acc);
  out[t] = acc;
}

__device__ __forceinline__ int lower_bound_f(const float* __restrict__ a, int n, float x) {  // #{a_m < x}
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int upper_bound_f(const float* __restrict__ a, int n, float x) {  // #{a_m <= x}
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) pair_stats_kernel(const PairStatsParams p) {
  __shared__ float As[kMsK][kMsTile + 4];
  __shared__ float Bs[kMsK][kMsTile + 4];
  __shared__ double s_sums[4];
  __shared__ unsigned long long s_counts[4 + 2 * kMsMaxThr];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;   // thread owns rows i0 + ty*4 + r, columns j0 + tx*4 + c
  const int i0 = blockIdx.y * kMsTile, j0 = blockIdx.x * kMsTile;
  if (tid < 4) s_sums[tid] = 0.0;
  if (tid < 4 + 2 * kMsMaxThr) s_counts[tid] = 0ull;

  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;

  for (int k0 = 0; k0 < p.dim; k0 += kMsK) {
    __syncthreads();
    // 64 rows x 16 k of each operand: thread loads 4 consecutive k of one row
    {
      const int row = tid >> 2, kq = (tid & 3) * 4;
      const int gi = i0 + row, gj = j0 + row;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (gi < p.n) a = *reinterpret_cast<const float4*>(p.e + (size_t)gi * p.dim + k0 + kq);
      if (gj < p.n) b = *reinterpret_cast<const float4*>(p.e + (size_t)gj * p.dim + k0 + kq);
      As[kq][row] = a.x; As[kq + 1][row] = a.y; As[kq + 2][row] = a.z; As[kq + 3][row] = a.w;
      Bs[kq][row] = b.x; Bs[kq + 1][row] = b.y; Bs[kq + 2][row] = b.z; Bs[kq + 3][row] = b.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kMsK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
    }
  }

  // ---- fold the 16 scores of this thread ----
  double sum_intra = 0.0, sq_intra = 0.0, sum_inter = 0.0, sq_inter = 0.0;
  unsigned int n_intra = 0, n_inter = 0, ge_intra[kMsMaxThr], ge_inter[kMsMaxThr];
  unsigned long long sum_upper = 0, sum_lower = 0;
#pragma unroll
  for (int t = 0; t < kMsMaxThr; ++t) ge_intra[t] = ge_inter[t] = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r;
    if (i >= p.n) continue;
    const int id_i = p.ids[i];
    const int pb = p.row_ptr[i], pe = p.row_ptr[i + 1];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + tx * 4 + c;
      if (j >= p.n || j == i) continue;
      const float s = acc[r][c];
      const bool same = p.ids[j] == id_i;
      if (same) { sum_intra += s; sq_intra += (double)s * s; ++n_intra; }
      else {
        sum_inter += s; sq_inter += (double)s * s; ++n_inter;
        sum_upper += (unsigned long long)upper_bound_f(p.sorted_intra, p.m, s);
        sum_lower += (unsigned long long)lower_bound_f(p.sorted_intra, p.m, s);
      }
#pragma unroll
      for (int t = 0; t < kMsMaxThr; ++t)
        if (t < p.n_thr && s >= p.thr[t]) { if (same) ++ge_intra[t]; else ++ge_inter[t]; }
      for (int q = pb; q < pe; ++q) {   // ranks of row i's positives
        const float sp = p.pos_score[q];
        if (s > sp) atomicAdd(&p.rank_greater[q], 1u);
        else if (s == sp && j < p.pos_idx[q]) atomicAdd(&p.rank_tie_before[q], 1u);
      }
    }
  }
  // CTA-level reduction through shared memory, one global atomic per counter and CTA
  atomicAdd(&s_sums[0], sum_intra); atomicAdd(&s_sums[1], sq_intra);
  atomicAdd(&s_sums[2], sum_inter); atomicAdd(&s_sums[3], sq_inter);
  atomicAdd(&s_counts[0], (unsigned long long)n_intra); atomicAdd(&s_counts[1], (unsigned long long)n_inter);
  atomicAdd(&s_counts[2], sum_upper); atomicAdd(&s_counts[3], sum_lower);
#pragma unroll
  for (int t = 0; t < kMsMaxThr; ++t) {
    if (ge_intra[t]) atomicAdd(&s_counts[4 + t], (unsigned long long)ge_intra[t]);
    if (ge_inter[t]) atomicAdd(&s_counts[4 + kMsMaxThr + t], (unsigned long long)ge_inter[t]);
  }
  __syncthreads();
  if (tid < 4) atomicAdd(&p.sums[tid], s_sums[tid]);
  if (tid < 4 + 2 * kMsMaxThr && s_counts[tid]) atomicAdd(&p.counts[tid], s_counts[tid]);
}

}  // namespace vfp
